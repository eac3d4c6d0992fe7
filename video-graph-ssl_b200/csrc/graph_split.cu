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

// Warp reduction of 64 per-lane values in 62 shuffles instead of 64 x 5: at every butterfly step a lane keeps half of its
// values and hands the other half to its partner, so after 5 steps lane L holds the two complete sums with indices
// 2 * bitrev-free code(L) + {0, 1}:  idx = 32*b4 + 16*b3 + 8*b2 + 4*b1 + 2*b0 + r  (b_k = bit k of L).  Fixed order.
__device__ __forceinline__ void warp_reduce64(float (&v)[64], int lane)
{
#pragma unroll
    for (int k = 0; k < 32; ++k) { const bool up = lane & 16; const float send = up ? v[k] : v[k + 32], keep = up ? v[k + 32] : v[k];
                                   v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16); }
#pragma unroll
    for (int k = 0; k < 16; ++k) { const bool up = lane & 8; const float send = up ? v[k] : v[k + 16], keep = up ? v[k + 16] : v[k];
                                   v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8); }
#pragma unroll
    for (int k = 0; k < 8; ++k) { const bool up = lane & 4; const float send = up ? v[k] : v[k + 8], keep = up ? v[k + 8] : v[k];
                                  v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4); }
#pragma unroll
    for (int k = 0; k < 4; ++k) { const bool up = lane & 2; const float send = up ? v[k] : v[k + 4], keep = up ? v[k + 4] : v[k];
                                  v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 2); }
#pragma unroll
    for (int k = 0; k < 2; ++k) { const bool up = lane & 1; const float send = up ? v[k] : v[k + 2], keep = up ? v[k + 2] : v[k];
                                  v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 1); }
}
__device__ __forceinline__ int warp_reduce64_index(int lane)
{
    return 32 * ((lane >> 4) & 1) + 16 * ((lane >> 3) & 1) + 8 * ((lane >> 2) & 1) + 4 * ((lane >> 1) & 1) + 2 * (lane & 1);
}

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

// same partials for T <= 8 and S % 4 == 0 with 128-bit loads: 16 independent float4 loads (256 B) in flight per thread
__global__ void __launch_bounds__(PD_THREADS, 2)
pairdots_partial_vec4_kernel(const float* __restrict__ A, const float* __restrict__ Bm, int Cn, int T, int S, int nchunk,
                             float* __restrict__ partials /* [videos, nchunk, T*T] */)
{
    __shared__ float red[PD_THREADS / 32][64];
    const int b = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int SV = S / 4;
    const long long E = (long long)Cn * SV;
    const long long e0 = E * chunk / nchunk, e1 = E * (chunk + 1) / nchunk;
    const float* Ab = A + (size_t)b * Cn * T * S;
    const float* Bb = Bm + (size_t)b * Cn * T * S;
    float* out = partials + ((size_t)b * nchunk + chunk) * T * T;
    float acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = 0.f;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long e = e0 + tid; e < e1; e += PD_THREADS) {
        const int c = (int)(e / SV), sv = (int)(e - (long long)c * SV);
        const size_t base = (size_t)c * T * S + (size_t)sv * 4;
        float4 a[8], bb[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = (i < T) ? __ldg(reinterpret_cast<const float4*>(Ab + base + (size_t)i * S)) : zero;
#pragma unroll
        for (int j = 0; j < 8; ++j) bb[j] = (j < T) ? __ldg(reinterpret_cast<const float4*>(Bb + base + (size_t)j * S)) : zero;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float v = acc[i * 8 + j];
                v = fmaf(a[i].x, bb[j].x, v); v = fmaf(a[i].y, bb[j].y, v);
                v = fmaf(a[i].z, bb[j].z, v); v = fmaf(a[i].w, bb[j].w, v);
                acc[i * 8 + j] = v;
            }
    }
    warp_reduce64(acc, lane);
    {
        const int idx = warp_reduce64_index(lane);
        red[warp][idx] = acc[0];
        red[warp][idx + 1] = acc[1];
    }
    __syncthreads();
    if (tid < 64) {
        const int i = tid / 8, j = tid % 8;
        if (i < T && j < T) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < PD_THREADS / 32; ++w) v += red[w][tid];
            out[i * T + j] = v;
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
    // chunk partials: four interleaved chains per element keep the loads independent; combined in a fixed order
    for (int p = threadIdx.x; p < T * T; p += G_THREADS) {
        float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
        const float* src = partials + (size_t)b * nchunk * T * T + p;
        int ch = 0;
        for (; ch + 3 < nchunk; ch += 4) {
            const float a0 = __ldg(src + (size_t)ch * T * T), a1 = __ldg(src + (size_t)(ch + 1) * T * T);
            const float a2 = __ldg(src + (size_t)(ch + 2) * T * T), a3 = __ldg(src + (size_t)(ch + 3) * T * T);
            v0 += a0; v1 += a1; v2 += a2; v3 += a3;
        }
        for (; ch < nchunk; ++ch) v0 += __ldg(src + (size_t)ch * T * T);
        m0[p] = (v0 + v1) + (v2 + v3);
    }
    __syncthreads();
    if (kBwd) {
        adj_backward(m0, a.sim + tt, a.adj + tt, a.s + tt, a.th, T, a.max_hop, a.inv_temp, a.opt, a.u ? a.u + tt : nullptr);
        for (int p = threadIdx.x; p < T * T; p += G_THREADS) a.dl[tt + p] = m0[p];
    } else {
        adj_forward(m0, m1, m2, a.u + tt, a.th, T, a.max_hop, a.inv_temp, a.sim + tt, a.adj + tt, a.s + tt, a.opt);
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
    {   // two CTAs are resident per SM: among the next few chunk counts take the one whose last wave is fullest
        const double slots = 2.0 * sm_count_cached();
        int best = nchunk; double best_fill = 0.0;
        const long long per_thread_min = (a.T <= 8 && S % 4 == 0) ? 8 : 8;      // elements a thread should still own
        for (int n = nchunk; n <= nchunk + 8 && n <= PD_MAX_CHUNKS && n <= E; ++n) {
            if (n > nchunk && E / n < per_thread_min * PD_THREADS) break;
            const double waves = n * (double)a.B / slots;
            const double fill = waves / (double)((long long)(waves + 0.999999));
            if (fill > best_fill + 1e-9) { best_fill = fill; best = n; }
        }
        nchunk = best;
    }
    if (a.T <= 8 && S % 4 == 0 && ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(Bm)) & 15) == 0)
        pairdots_partial_vec4_kernel<<<dim3(nchunk, a.B), PD_THREADS, 0, st>>>(A, Bm, Cn, a.T, S, nchunk, scratch);
    else
        pairdots_partial_kernel<<<dim3(nchunk, a.B), PD_THREADS, 0, st>>>(A, Bm, Cn, a.T, S, nchunk, scratch);
    GCA_LAUNCH_CHECK("pairdots_partial_kernel");
    if (bwd) adj_from_partials_kernel<true><<<a.B, G_THREADS, 0, st>>>(a, scratch, nchunk);
    else     adj_from_partials_kernel<false><<<a.B, G_THREADS, 0, st>>>(a, scratch, nchunk);
    GCA_LAUNCH_CHECK("adj_from_partials_kernel");
    count_launch(2);
    return GCA_OK;
}

}  // namespace gca
