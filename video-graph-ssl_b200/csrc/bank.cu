// K13: NPID instance bank (reference: lib/memory/mem_bank.py:15-90, RGBMem / CMCMem).
//
// The reference gathers bsz * (K + 1) bank rows into a [bsz, K + 1, d] tensor (2.1 GB at bsz = 256, K = 16384, d = 128),
// multiplies it with the features (bmm) and keeps it alive for autograd.  Here the gathered rows are never materialised:
//   bank_logits_kernel   logits[b, j] = <bank[idx[b, j]], x[b]> / T      one warp per sampled row, 128-bit coalesced row
//                                                                        reads, four rows in flight per warp
//   bank_dx_kernel       dx[b] = sum_j g[b, j] bank[idx[b, j]] / T       the same gather again for the gradient (the bank
//                                                                        is a buffer: nothing flows into it), split over a
//                                                                        few CTAs per feature row, fixed-order partial sums
//   bank_update_kernel   rows y <- normalize(m * row + (1 - m) * x)      mem_bank.py:15-27, computed from the OLD rows into a
//                                                                        scratch block, then scattered: of a duplicated index
//                                                                        the last occurrence wins (index_copy_ on the CPU)
// HBM / L2-bound gathers: (K + 1) * d * 4 bytes per feature row and direction, no reuse between rows of a batch except
// through L2.  Deterministic: no floating-point atomics, fixed association orders.
#include "gca_common.cuh"

namespace gca {

constexpr int BK_THREADS = 256;
constexpr int BK_MAX_CH = 8;                          // float4 chunks a lane holds per row: d <= 1024

__device__ __forceinline__ float dot4(const float4 a, const float4 b, float acc)
{
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc))));
}

// grid (ceil((K+1) / rows per CTA), B); the warps of a CTA take its slice eight rows at a time
template <int NCH>                                  // float4 chunks per lane and row: d <= 128 * NCH
__global__ void __launch_bounds__(BK_THREADS)
bank_logits_kernel(const float* __restrict__ x, const float* __restrict__ bank, const long long* __restrict__ idx, int K1, int d,
                   long long n_data, float inv_T, int rows_per_cta, float* __restrict__ logits)
{
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int d4 = d >> 2;
    float4 xr[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int i = lane + 32 * c;
        xr[c] = (i < d4) ? __ldg(reinterpret_cast<const float4*>(x + (size_t)b * d) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int j0 = blockIdx.x * rows_per_cta, j1 = min(K1, j0 + rows_per_cta);
    const long long* irow = idx + (size_t)b * K1;
    float* lrow = logits + (size_t)b * K1;
    constexpr int NW = BK_THREADS / 32, U = 8;        // rows in flight per warp
    long long rn[U];                                  // indices of the NEXT iteration: their load overlaps this one's row reads
#pragma unroll
    for (int u = 0; u < U; ++u) { const int jj = j0 + warp * U + u; rn[u] = (jj < j1) ? __ldg(irow + jj) : -1; }
    for (int j = j0 + warp * U; j < j1; j += NW * U) {
        long long r[U];
        float acc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { r[u] = rn[u]; acc[u] = 0.f; }
#pragma unroll
        for (int u = 0; u < U; ++u) { const int jj = j + NW * U + u; rn[u] = (jj < j1) ? __ldg(irow + jj) : -1; }
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int i = lane + 32 * c;
            if (i < d4) {
                float4 v[U];
#pragma unroll
                for (int u = 0; u < U; ++u)
                    v[u] = (r[u] >= 0 && r[u] < n_data) ? __ldg(reinterpret_cast<const float4*>(bank + (size_t)r[u] * d) + i)
                                                        : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < U; ++u) acc[u] = dot4(v[u], xr[c], acc[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float s = warp_sum(acc[u]);
            if (lane == 0 && j + u < j1)
                lrow[j + u] = (r[u] >= 0 && r[u] < n_data) ? s * inv_T : __int_as_float(0x7fc00000);   // index out of range: NaN
        }
    }
}

// grid (S, B): CTA (s, b) sums the rows j = s * 8 + warp, + S * 8, ... of feature row b (fixed order per warp), the eight
// warps are added in order, the S partials of a row by bank_dx_reduce_kernel in order
template <int NCH>
__global__ void __launch_bounds__(BK_THREADS)
bank_dx_kernel(const float* __restrict__ g, const float* __restrict__ bank, const long long* __restrict__ idx, int K1, int d,
               long long n_data, float* __restrict__ part /* [S, B, d] */)
{
    extern __shared__ float4 sm[];                    // [8 warps][d4]
    const int b = blockIdx.y, S = gridDim.x, B = gridDim.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int d4 = d >> 2;
    constexpr int NW = BK_THREADS / 32;
    float4 acc[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    const long long* irow = idx + (size_t)b * K1;
    const float* grow = g + (size_t)b * K1;
    const int stride = S * NW;
    constexpr int U = 8;                              // rows in flight per warp
    long long rn[U];
    float wn[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int jj = blockIdx.x * NW + warp + u * stride;
        rn[u] = (jj < K1) ? __ldg(irow + jj) : -1;
        wn[u] = (jj < K1) ? __ldg(grow + jj) : 0.f;
    }
    for (int j = blockIdx.x * NW + warp; j < K1; j += U * stride) {
        long long r[U];
        float w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { r[u] = rn[u]; w[u] = wn[u]; }
#pragma unroll
        for (int u = 0; u < U; ++u) {                  // indices / weights of the next iteration: overlap this one's row reads
            const int jj = j + (U + u) * stride;
            rn[u] = (jj < K1) ? __ldg(irow + jj) : -1;
            wn[u] = (jj < K1) ? __ldg(grow + jj) : 0.f;
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int i = lane + 32 * c;
            if (i < d4) {
                float4 v[U];
#pragma unroll
                for (int u = 0; u < U; ++u)
                    v[u] = (r[u] >= 0 && r[u] < n_data) ? __ldg(reinterpret_cast<const float4*>(bank + (size_t)r[u] * d) + i)
                                                        : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    acc[c].x = fmaf(w[u], v[u].x, acc[c].x); acc[c].y = fmaf(w[u], v[u].y, acc[c].y);
                    acc[c].z = fmaf(w[u], v[u].z, acc[c].z); acc[c].w = fmaf(w[u], v[u].w, acc[c].w);
                }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) { const int i = lane + 32 * c; if (i < d4) sm[warp * d4 + i] = acc[c]; }
    __syncthreads();
    for (int i = threadIdx.x; i < d4; i += BK_THREADS) {
        float4 t = sm[i];
#pragma unroll
        for (int w2 = 1; w2 < NW; ++w2) { const float4 o = sm[w2 * d4 + i]; t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w; }
        reinterpret_cast<float4*>(part + ((size_t)blockIdx.x * B + b) * d)[i] = t;
    }
}

__global__ void __launch_bounds__(BK_THREADS)
bank_dx_reduce_kernel(const float* __restrict__ part, int S, long long n /* B * d */, float inv_T, float* __restrict__ dx)
{
    const long long i = (long long)blockIdx.x * BK_THREADS + threadIdx.x;
    if (i >= n) return;
    float t = part[i];
    for (int s = 1; s < S; ++s) t += part[(size_t)s * n + i];
    dx[i] = t * inv_T;
}

// stage 1: one warp per updated row n: tmp[n] = normalize(m * bank[y[n]] + (1 - m) * x[n]) -- mul, mul, add as separate fp32
// roundings (mul_ / mul / add_ of the reference), x / max(||x||, 1e-12) like F.normalize
__global__ void __launch_bounds__(BK_THREADS)
bank_update_rows_kernel(const float* __restrict__ bank, const float* __restrict__ x, const long long* __restrict__ y, int N, int d,
                        long long n_data, float m, float one_minus_m, float* __restrict__ tmp)
{
    const int n = blockIdx.x * (BK_THREADS / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= N) return;
    const long long r = __ldg(y + n);
    if (r < 0 || r >= n_data) return;                 // out of range: the scatter below skips the row as well
    const int d4 = d >> 2;
    float4 v[BK_MAX_CH];
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < BK_MAX_CH; ++c) {
        const int i = lane + 32 * c;
        if (i < d4) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(bank + (size_t)r * d) + i);
            const float4 f = __ldg(reinterpret_cast<const float4*>(x + (size_t)n * d) + i);
            v[c].x = __fadd_rn(__fmul_rn(w.x, m), __fmul_rn(f.x, one_minus_m));
            v[c].y = __fadd_rn(__fmul_rn(w.y, m), __fmul_rn(f.y, one_minus_m));
            v[c].z = __fadd_rn(__fmul_rn(w.z, m), __fmul_rn(f.z, one_minus_m));
            v[c].w = __fadd_rn(__fmul_rn(w.w, m), __fmul_rn(f.w, one_minus_m));
            ss = fmaf(v[c].x, v[c].x, fmaf(v[c].y, v[c].y, fmaf(v[c].z, v[c].z, fmaf(v[c].w, v[c].w, ss))));
        }
    }
    ss = warp_sum(ss);
    const float nrm = fmaxf(sqrtf(ss), 1e-12f);       // (this file is built without --use_fast_math: IEEE sqrt and division)
#pragma unroll
    for (int c = 0; c < BK_MAX_CH; ++c) {
        const int i = lane + 32 * c;
        if (i < d4) reinterpret_cast<float4*>(tmp + (size_t)n * d)[i] = make_float4(v[c].x / nrm, v[c].y / nrm, v[c].z / nrm, v[c].w / nrm);
    }
}

// stage 2: scatter; a row whose index occurs again later in y is skipped (the last occurrence wins)
__global__ void __launch_bounds__(BK_THREADS)
bank_update_scatter_kernel(float* __restrict__ bank, const float* __restrict__ tmp, const long long* __restrict__ y, int N, int d,
                           long long n_data)
{
    const int n = blockIdx.x * (BK_THREADS / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= N) return;
    const long long r = __ldg(y + n);
    if (r < 0 || r >= n_data) return;
    bool later = false;
    for (int o = n + 1 + lane; o < N; o += 32) later |= (__ldg(y + o) == r);
    if (__any_sync(0xffffffffu, later)) return;
    const int d4 = d >> 2;
    for (int i = lane; i < d4; i += 32)
        reinterpret_cast<float4*>(bank + (size_t)r * d)[i] = __ldg(reinterpret_cast<const float4*>(tmp + (size_t)n * d) + i);
}

static int bank_dx_splits(int B)
{
    int sms = sm_count_cached();
    if (sms < 1) sms = 148;
    int S = (4 * sms + B - 1) / B;                    // about four CTAs per SM in all
    if (S < 1) S = 1;
    if (S > 16) S = 16;
    return S;
}

}  // namespace gca

static int bank_check(const char* fn, int B, int K1, int d, long long n_data)
{
    using namespace gca;
    GCA_CHECK_ARG(B >= 1 && K1 >= 1 && n_data >= 1, "%s: need B >= 1, K + 1 >= 1, n_data >= 1", fn);
    if (d < 4 || d % 4 != 0 || d > 128 * BK_MAX_CH)
        return set_err(GCA_ERR_UNSUPPORTED, "%s: d must be a multiple of 4 in [4, %d] (d=%d)", fn, 128 * BK_MAX_CH, d);
    if (sm_count_cached() < 1) return set_err(GCA_ERR_CUDA, "no CUDA device available (this library has no CPU path)");
    return GCA_OK;
}

extern "C" int gca_bank_logits(const float* x, const float* bank, const long long* idx, int B, int K1, int d, long long n_data,
                               float inv_T, float* logits, void* stream)
{
    using namespace gca;
    GCA_CHECK_ARG(x && bank && idx && logits, "gca_bank_logits: null pointer");
    int rc = bank_check("gca_bank_logits", B, K1, d, n_data);
    if (rc != GCA_OK) return rc;
    GCA_CHECK_ARG(B <= 65535, "gca_bank_logits: B <= 65535");
    // rows per CTA: a multiple of 64 (8 warps x 8 rows), enough CTAs to fill the GPU a few times over
    int sms = sm_count_cached();
    long long want = ((long long)B * K1 + 8ll * sms - 1) / (8ll * sms);
    int rows = (int)((want + 63) / 64 * 64);
    if (rows < 64) rows = 64;
    if (rows > 1024) rows = 1024;
    dim3 grid((K1 + rows - 1) / rows, B);
    const int nch = (d / 4 + 31) / 32;
#define GCA_BANK_LOGITS(N) bank_logits_kernel<N><<<grid, BK_THREADS, 0, (cudaStream_t)stream>>>(x, bank, idx, K1, d, n_data, inv_T, rows, logits)
    if (nch <= 1) GCA_BANK_LOGITS(1); else if (nch <= 2) GCA_BANK_LOGITS(2); else if (nch <= 4) GCA_BANK_LOGITS(4); else GCA_BANK_LOGITS(8);
#undef GCA_BANK_LOGITS
    GCA_LAUNCH_CHECK("bank_logits_kernel");
    count_launch(1);
    return GCA_OK;
}

extern "C" size_t gca_bank_dx_workspace_bytes(int B, int d)
{
    if (B < 1 || d < 1) return 0;
    return gca::align_up((size_t)gca::bank_dx_splits(B) * B * d * sizeof(float), 256);
}

extern "C" int gca_bank_dx(const float* g_logits, const float* bank, const long long* idx, int B, int K1, int d, long long n_data,
                           float inv_T, float* dx, void* workspace, size_t workspace_bytes, void* stream)
{
    using namespace gca;
    GCA_CHECK_ARG(g_logits && bank && idx && dx, "gca_bank_dx: null pointer");
    int rc = bank_check("gca_bank_dx", B, K1, d, n_data);
    if (rc != GCA_OK) return rc;
    GCA_CHECK_ARG(B <= 65535, "gca_bank_dx: B <= 65535");
    const size_t need = gca_bank_dx_workspace_bytes(B, d);
    if (!workspace || workspace_bytes < need)
        return set_err(GCA_ERR_WORKSPACE, "gca_bank_dx: workspace too small: %zu bytes given, %zu needed", workspace_bytes, need);
    const int S = bank_dx_splits(B);
    float* part = (float*)workspace;
    const size_t smem = (size_t)(BK_THREADS / 32) * (d / 4) * sizeof(float4);
    const int nch = (d / 4 + 31) / 32;
#define GCA_BANK_DX(N) bank_dx_kernel<N><<<dim3(S, B), BK_THREADS, smem, (cudaStream_t)stream>>>(g_logits, bank, idx, K1, d, n_data, part)
    if (nch <= 1) GCA_BANK_DX(1); else if (nch <= 2) GCA_BANK_DX(2); else if (nch <= 4) GCA_BANK_DX(4); else GCA_BANK_DX(8);
#undef GCA_BANK_DX
    GCA_LAUNCH_CHECK("bank_dx_kernel");
    const long long n = (long long)B * d;
    bank_dx_reduce_kernel<<<(unsigned)((n + BK_THREADS - 1) / BK_THREADS), BK_THREADS, 0, (cudaStream_t)stream>>>(part, S, n, inv_T, dx);
    GCA_LAUNCH_CHECK("bank_dx_reduce_kernel");
    count_launch(2);
    return GCA_OK;
}

extern "C" size_t gca_bank_update_workspace_bytes(int N, int d)
{
    if (N < 1 || d < 1) return 0;
    return gca::align_up((size_t)N * d * sizeof(float), 256);
}

extern "C" int gca_bank_update(float* bank, const float* x, const long long* y, int N, int d, long long n_data, float momentum,
                               float one_minus_momentum, void* workspace, size_t workspace_bytes, void* stream)
{
    using namespace gca;
    GCA_CHECK_ARG(bank && x && y, "gca_bank_update: null pointer");
    int rc = bank_check("gca_bank_update", N, 1, d, n_data);
    if (rc != GCA_OK) return rc;
    const size_t need = gca_bank_update_workspace_bytes(N, d);
    if (!workspace || workspace_bytes < need)
        return set_err(GCA_ERR_WORKSPACE, "gca_bank_update: workspace too small: %zu bytes given, %zu needed", workspace_bytes, need);
    float* tmp = (float*)workspace;
    const int blocks = (N + BK_THREADS / 32 - 1) / (BK_THREADS / 32);
    bank_update_rows_kernel<<<blocks, BK_THREADS, 0, (cudaStream_t)stream>>>(bank, x, y, N, d, n_data, momentum, one_minus_momentum, tmp);
    GCA_LAUNCH_CHECK("bank_update_rows_kernel");
    bank_update_scatter_kernel<<<blocks, BK_THREADS, 0, (cudaStream_t)stream>>>(bank, tmp, y, N, d, n_data);
    GCA_LAUNCH_CHECK("bank_update_scatter_kernel");
    count_launch(2);
    return GCA_OK;
}
