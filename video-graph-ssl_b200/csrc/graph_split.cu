// Large feature maps: the T x T pair dots are spread over the whole grid.
//   pairdots_partial_kernel : grid (chunks, videos); each CTA streams a contiguous range of (channel, position) elements of
//                             both operands straight from global memory (coalesced along the position axis, 2*T independent
//                             loads in flight per thread), keeps an 8 x 8 block of dot products in registers and writes one
//                             [T, T] partial per chunk
//   adj_from_partials_kernel: one small CTA per video sums the chunk partials in a fixed order (deterministic) and runs the
//                             T x T element chain (forward: softmax / hop weights / relaxed-Bernoulli; backward: its chain rule)
// followed by graph_agg_kernel (graph.cu).  Replaces temporal_graph.py:161-176 (forward) and the autograd of :56-64, 150-210.
#include "graph_dev.cuh"

namespace gca {

constexpr int PD_THREADS = 256;
constexpr int PD_MAX_CHUNKS = 64;

__global__ void __launch_bounds__(PD_THREADS)
pairdots_partial_kernel(const float* __restrict__ A, const float* __restrict__ Bm, int Cn, int T, int S, int nchunk,
                        float* __restrict__ partials /* [videos, nchunk, T*T] */)
{
    __shared__ float red[PD_THREADS / 32][64];
    const int b = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long E = (long long)Cn * S;
    const long long e0 = E * chunk / nchunk, e1 = E * (chunk + 1) / nchunk;
    const float* Ab = A + (size_t)b * Cn * T * S;
    const float* Bb = Bm + (size_t)b * Cn * T * S;
    float* out = partials + ((size_t)b * nchunk + chunk) * T * T;
    for (int ib = 0; ib < T; ib += 8) {
        for (int jb = 0; jb < T; jb += 8) {
            float acc[8][8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
            long long e = e0 + tid;
            int c = (int)(e / S), s = (int)(e - (long long)c * S);
            for (; e < e1; e += PD_THREADS) {
                const size_t base = (size_t)c * T * S + s;
                float a[8], bb[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = (ib + i < T) ? __ldg(Ab + base + (size_t)(ib + i) * S) : 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) bb[j] = (jb + j < T) ? __ldg(Bb + base + (size_t)(jb + j) * S) : 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
                s += PD_THREADS;
                while (s >= S) { s -= S; ++c; }
            }
            // fixed-order reduction: lanes (xor shuffles), then warps in index order
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float v = warp_sum(acc[i][j]);
                    if (lane == 0) red[warp][i * 8 + j] = v;
                }
            __syncthreads();
            if (tid < 64) {
                const int i = ib + tid / 8, j = jb + tid % 8;
                if (i < T && j < T) {
                    float v = 0.f;
#pragma unroll
                    for (int w = 0; w < PD_THREADS / 32; ++w) v += red[w][tid];
                    out[i * T + j] = v;
                }
            }
            __syncthreads();
        }
    }
}

template <bool kBwd>
__global__ void __launch_bounds__(G_THREADS)
adj_from_partials_kernel(const GraphArgs a, const float* __restrict__ partials, int nchunk)
{
    __shared__ float m0[G_TMAXMAX * G_TMAXMAX], m1[G_TMAXMAX * G_TMAXMAX], m2[G_TMAXMAX * G_TMAXMAX];
    const int b = blockIdx.x, T = a.T;
    const size_t tt = (size_t)b * T * T;
    for (int p = threadIdx.x; p < T * T; p += G_THREADS) {
        float v = 0.f;
        const float* src = partials + (size_t)b * nchunk * T * T + p;
        for (int ch = 0; ch < nchunk; ++ch) v += __ldg(src + (size_t)ch * T * T);
        m0[p] = v;
    }
    __syncthreads();
    if (kBwd) {
        adj_backward(m0, a.sim + tt, a.adj + tt, a.s + tt, a.th, T, a.max_hop, a.inv_temp);
        for (int p = threadIdx.x; p < T * T; p += G_THREADS) a.dl[tt + p] = m0[p];
    } else {
        adj_forward(m0, m1, m2, a.u + tt, a.th, T, a.max_hop, a.inv_temp, a.sim + tt, a.adj + tt, a.s + tt);
    }
}

size_t graph_split_scratch_floats(int B, int T) { return (size_t)B * PD_MAX_CHUNKS * T * T; }

// pair dots of (A, Bm) [B, Cn, T, S] -> element chain; `scratch` holds graph_split_scratch_floats floats
int graph_split_adj_launch(const GraphArgs& a, bool bwd, float* scratch, cudaStream_t st)
{
    const float* A = bwd ? a.dy : a.gq;
    const float* Bm = bwd ? a.support : a.gk;
    const int Cn = bwd ? a.C : a.Cq, S = bwd ? a.HW : a.S;
    const long long E = (long long)Cn * S;
    int nchunk = (int)((E + 4095) / 4096);                    // ~4096 (channel, position) elements per CTA
    const int want = (2 * sm_count_cached() + a.B - 1) / a.B;  // but at least ~2 CTAs per SM over the whole grid
    if (nchunk < want) nchunk = want;
    if (nchunk > PD_MAX_CHUNKS) nchunk = PD_MAX_CHUNKS;
    if (nchunk > E) nchunk = (int)E;
    if (nchunk < 1) nchunk = 1;
    pairdots_partial_kernel<<<dim3(nchunk, a.B), PD_THREADS, 0, st>>>(A, Bm, Cn, a.T, S, nchunk, scratch);
    GCA_LAUNCH_CHECK("pairdots_partial_kernel");
    if (bwd) adj_from_partials_kernel<true><<<a.B, G_THREADS, 0, st>>>(a, scratch, nchunk);
    else     adj_from_partials_kernel<false><<<a.B, G_THREADS, 0, st>>>(a, scratch, nchunk);
    GCA_LAUNCH_CHECK("adj_from_partials_kernel");
    count_launch(2);
    return GCA_OK;
}

}  // namespace gca
