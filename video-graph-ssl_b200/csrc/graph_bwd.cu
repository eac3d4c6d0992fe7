// K5: temporal clip-graph head, backward adjacency kernel (one CTA per video): ds = <dy_i, support_j>, the
// T x T chain rule through sigmoid / logit / clamp / hop weights / softmax, and -- on the
// fused path -- the three input gradients, all without leaving the CTA.
#include "graph_dev.cuh"

namespace gca {

template <int TMAX, int VS, int VH, bool kAgg>
__global__ void __launch_bounds__(G_THREADS)
graph_bwd_kernel(const GraphArgs a)
{
    extern __shared__ __align__(16) float gsm[];
    float* tiles = gsm;
    float* red = tiles + 2 * G_CHUNK_FLOATS;
    float* m0 = red + 4 * G_THREADS;                  // ds -> d_logit
    float* m1 = m0 + G_TMAXMAX * G_TMAXMAX;           // s
    const int b = blockIdx.x, T = a.T;
    const size_t tt = (size_t)b * T * T;
    const size_t offH = (size_t)b * a.C * T * a.HW, offS = (size_t)b * a.Cq * T * a.S;
    pair_dots(a.dy + offH, a.support + offH, a.C, T, a.HW, tiles, red, m0);      // ds[i][j] = <dy_i, support_j>
    adj_backward(m0, a.sim + tt, a.adj + tt, a.s + tt, a.th, T, a.max_hop, a.inv_temp, a.opt, a.u ? a.u + tt : nullptr);
    if constexpr (kAgg) {
        for (int p = threadIdx.x; p < T * T; p += G_THREADS) m1[p] = __ldg(a.s + tt + p);
        __syncthreads();
        aggregate_items<TMAX, VH>(a.dy + offH, a.d_support + offH, m1, true, true, a.C, T, a.HW, threadIdx.x, G_THREADS);
        aggregate_items<TMAX, VS>(a.gk + offS, a.d_gq + offS, m0, false, false, a.Cq, T, a.S, threadIdx.x, G_THREADS);
        aggregate_items<TMAX, VS>(a.gq + offS, a.d_gk + offS, m0, true, false, a.Cq, T, a.S, threadIdx.x, G_THREADS);
    } else {
        for (int p = threadIdx.x; p < T * T; p += G_THREADS) a.dl[tt + p] = m0[p];
    }
}

template <int TM, int VS, int VH, bool kAgg>
static int launch_bwd_one(const GraphArgs& a, cudaStream_t st)
{
    const size_t smem = G_SMEM_FLOATS * sizeof(float);
    auto kern = graph_bwd_kernel<TM, VS, VH, kAgg>;
    GCA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<a.B, G_THREADS, smem, st>>>(a);
    GCA_LAUNCH_CHECK("graph_bwd_kernel");
    count_launch(1);
    return GCA_OK;
}

template <int TM>
static int launch_bwd_vec(const GraphArgs& a, cudaStream_t st)
{
    const bool vs = (a.S % 4 == 0), vh = (a.HW % 4 == 0);
    if (vs && vh) return launch_bwd_one<TM, 4, 4, true>(a, st);
    if (vs)       return launch_bwd_one<TM, 4, 1, true>(a, st);
    if (vh)       return launch_bwd_one<TM, 1, 4, true>(a, st);
    return launch_bwd_one<TM, 1, 1, true>(a, st);
}

int graph_bwd_adj_launch(const GraphArgs& a, bool fused, cudaStream_t st)
{
    if (!fused) return launch_bwd_one<4, 1, 1, false>(a, st);
    switch (pick_tmax(a.T)) {
        case 4:  return launch_bwd_vec<4>(a, st);
        case 8:  return launch_bwd_vec<8>(a, st);
        case 16: return launch_bwd_vec<16>(a, st);
        default: return launch_bwd_one<32, 1, 1, true>(a, st);
    }
}

}  // namespace gca
