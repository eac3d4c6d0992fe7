// Embedding-level graph head: ONE CTA per video with the whole video staged in shared memory.
// All of a video's operands (projections, GCN support, uniforms; plus the output gradient in backward) are contiguous blocks
// in global memory, so the CTA issues one bulk async copy (cp.async.bulk -> UBLKCP, completion on an mbarrier) per operand
// at entry -- a single HBM round trip -- and then runs similarity, softmax, hop weights, relaxed-Bernoulli re-sampling and
// the aggregation (forward) or the whole chain rule (backward) out of shared memory with warp shuffles.
// Used when a video fits (<= ~190 KB) and every block is a multiple of 16 bytes; larger videos take graph_split.cu.
#include "graph_dev.cuh"
#include "tc_ptx.cuh"

namespace gca {

constexpr size_t GS_MAX_SMEM = 200 * 1024;

// M[i][j] = sum_{c,s} A[(c*T + i)*S + s] * B[(c*T + j)*S + s] with A, B in shared memory (T <= 16: one pair per lane);
// result in out_tt (shared).  Thread = (pair, channel group); four independent partial sums per thread keep the FMA pipe
// fed (a single dependent chain was latency-bound), and everything is added in a fixed order.
__device__ void pair_dots_smem(const float* A, const float* Bm, int Cn, int T, int S, float* red, float* out_tt)
{
    const int tid = threadIdx.x;
    const int npairs = T * T;                                   // <= 256
    const int G = G_THREADS / npairs;                           // channel groups
    const int plane = tid % npairs, group = tid / npairs;
    const int pi = plane / T, pj = plane % T;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (group < G) {
        const float* ra = A + (size_t)pi * S;
        const float* rb = Bm + (size_t)pj * S;
        const int cs = T * S;                                    // floats per channel
        if (S == 1) {
            int c = group;
            for (; c + 3 * G < Cn; c += 4 * G) {
                a0 = fmaf(ra[(size_t)c * cs], rb[(size_t)c * cs], a0);
                a1 = fmaf(ra[(size_t)(c + G) * cs], rb[(size_t)(c + G) * cs], a1);
                a2 = fmaf(ra[(size_t)(c + 2 * G) * cs], rb[(size_t)(c + 2 * G) * cs], a2);
                a3 = fmaf(ra[(size_t)(c + 3 * G) * cs], rb[(size_t)(c + 3 * G) * cs], a3);
            }
            for (; c < Cn; c += G) a0 = fmaf(ra[(size_t)c * cs], rb[(size_t)c * cs], a0);
        } else {
            for (int c = group; c < Cn; c += G) {
                const float* pa = ra + (size_t)c * cs;
                const float* pb = rb + (size_t)c * cs;
                int x = 0;
                for (; x + 3 < S; x += 4) {
                    a0 = fmaf(pa[x], pb[x], a0);         a1 = fmaf(pa[x + 1], pb[x + 1], a1);
                    a2 = fmaf(pa[x + 2], pb[x + 2], a2); a3 = fmaf(pa[x + 3], pb[x + 3], a3);
                }
                for (; x < S; ++x) a0 = fmaf(pa[x], pb[x], a0);
            }
        }
    }
    red[tid] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    for (int p = tid; p < npairs; p += G_THREADS) {
        float v = 0.f;
        for (int g = 0; g < G; ++g) v += red[g * npairs + p];
        out_tt[p] = v;
    }
    __syncthreads();
}

struct GsLayout { size_t q, k, sup, dy, u, mats, red, bar, total; };

static __host__ __device__ GsLayout gs_layout(const GraphArgs& a, bool bwd)
{
    GsLayout L;
    const size_t nq = (size_t)a.Cq * a.T * a.S, ns = (size_t)a.C * a.T * a.HW, tt = (size_t)a.T * a.T;
    size_t off = 0;
    L.q = off;   off += nq;
    L.k = off;   off += nq;
    L.sup = off; off += ns;
    L.dy = off;  off += bwd ? ns : 0;
    L.u = off;   off += (tt + 3) / 4 * 4;
    L.mats = off; off += 3 * ((tt + 3) / 4 * 4);
    L.red = off; off += 4 * G_THREADS;
    L.bar = off; off += 4;
    L.total = off * sizeof(float);
    return L;
}

template <int TMAX, int VS, int VH, bool kBwd>
__global__ void __launch_bounds__(G_THREADS)
graph_smem_kernel(const GraphArgs a)
{
    extern __shared__ __align__(128) float gsm[];
    const GsLayout L = gs_layout(a, kBwd);
    float* q_s = gsm + L.q;  float* k_s = gsm + L.k;  float* sup_s = gsm + L.sup;  float* dy_s = gsm + L.dy;
    float* u_s = gsm + L.u;  float* m0 = gsm + L.mats;
    const int T = a.T, tt_n = T * T, tt_pad = (tt_n + 3) / 4 * 4;
    float* m1 = m0 + tt_pad;  float* m2 = m1 + tt_pad;
    float* red = gsm + L.red;
    uint64_t* bar = reinterpret_cast<uint64_t*>(gsm + L.bar);
    const int b = blockIdx.x;
    const size_t tt = (size_t)b * tt_n;
    const size_t nq = (size_t)a.Cq * T * a.S, ns = (size_t)a.C * T * a.HW;

    if (threadIdx.x == 0) {
        ptx::mbar_init(bar, 1);
        ptx::fence_barrier_init();
        const uint32_t bq = (uint32_t)(nq * 4), bs = (uint32_t)(ns * 4), bu = (uint32_t)(tt_n * 4);
        ptx::mbar_arrive_expect_tx(bar, 2 * bq + bs * (kBwd ? 2u : 1u) + (kBwd ? 0u : bu));
        ptx::bulk_load(q_s, a.gq + (size_t)b * nq, bq, bar);
        ptx::bulk_load(k_s, a.gk + (size_t)b * nq, bq, bar);
        ptx::bulk_load(sup_s, a.support + (size_t)b * ns, bs, bar);
        if (kBwd) ptx::bulk_load(dy_s, a.dy + (size_t)b * ns, bs, bar);
        else      ptx::bulk_load(u_s, a.u + tt, bu, bar);
    }
    __syncthreads();                 // the barrier is initialised before anyone polls it
    ptx::mbar_wait(bar, 0);

    if (!kBwd) {
        pair_dots_smem(q_s, k_s, a.Cq, T, a.S, red, m0);
        // adj_forward reads the uniforms through a pointer: give it the staged copy
        adj_forward(m0, m1, m2, u_s, a.th, T, a.max_hop, a.inv_temp, a.sim + tt, a.adj + tt, a.s + tt, a.opt);
        aggregate_items<TMAX, VH, true>(sup_s, a.y + (size_t)b * ns, m0, false, true, a.C, T, a.HW, threadIdx.x, G_THREADS);
    } else {
        pair_dots_smem(dy_s, sup_s, a.C, T, a.HW, red, m0);               // ds[i][j] = <dy_i, support_j>
        adj_backward(m0, a.sim + tt, a.adj + tt, a.s + tt, a.th, T, a.max_hop, a.inv_temp, a.opt, a.u ? a.u + tt : nullptr);
        for (int p = threadIdx.x; p < tt_n; p += G_THREADS) m1[p] = __ldg(a.s + tt + p);
        __syncthreads();
        aggregate_items<TMAX, VH, true>(dy_s, a.d_support + (size_t)b * ns, m1, true, true, a.C, T, a.HW, threadIdx.x, G_THREADS);
        aggregate_items<TMAX, VS, true>(k_s, a.d_gq + (size_t)b * nq, m0, false, false, a.Cq, T, a.S, threadIdx.x, G_THREADS);
        aggregate_items<TMAX, VS, true>(q_s, a.d_gk + (size_t)b * nq, m0, true, false, a.Cq, T, a.S, threadIdx.x, G_THREADS);
    }
}

bool graph_smem_fits(const GraphArgs& a, bool bwd)
{
    if (a.T > 16) return false;                                       // register blocks of the aggregation
    const size_t nq = (size_t)a.Cq * a.T * a.S, ns = (size_t)a.C * a.T * a.HW, tt = (size_t)a.T * a.T;
    if ((nq * 4) % 16 || (ns * 4) % 16 || (!bwd && (tt * 4) % 16)) return false;   // bulk copies move multiples of 16 bytes
    return gs_layout(a, bwd).total <= GS_MAX_SMEM;
}

template <int TM, int VS, int VH, bool kBwd>
static int launch_smem_one(const GraphArgs& a, cudaStream_t st)
{
    const size_t smem = gs_layout(a, kBwd).total;
    auto kern = graph_smem_kernel<TM, VS, VH, kBwd>;
    GCA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<a.B, G_THREADS, smem, st>>>(a);
    GCA_LAUNCH_CHECK("graph_smem_kernel");
    count_launch(1);
    return GCA_OK;
}

template <int TM, bool kBwd>
static int launch_smem_vec(const GraphArgs& a, cudaStream_t st)
{
    const bool vs = (a.S % 4 == 0), vh = (a.HW % 4 == 0);
    if (!kBwd) return vh ? launch_smem_one<TM, 1, 4, false>(a, st) : launch_smem_one<TM, 1, 1, false>(a, st);
    if (vs && vh) return launch_smem_one<TM, 4, 4, true>(a, st);
    if (vs)       return launch_smem_one<TM, 4, 1, true>(a, st);
    if (vh)       return launch_smem_one<TM, 1, 4, true>(a, st);
    return launch_smem_one<TM, 1, 1, true>(a, st);
}

int graph_smem_launch(const GraphArgs& a, bool bwd, cudaStream_t st)
{
    switch (pick_tmax(a.T)) {
        case 4:  return bwd ? launch_smem_vec<4, true>(a, st) : launch_smem_vec<4, false>(a, st);
        case 8:  return bwd ? launch_smem_vec<8, true>(a, st) : launch_smem_vec<8, false>(a, st);
        default: return bwd ? launch_smem_vec<16, true>(a, st) : launch_smem_vec<16, false>(a, st);
    }
}

}  // namespace gca
