// K7: retrieval -- cosine similarity matrix + per-query top-k (k <= 64), replacing sklearn cosine_distances +
// full-row np.argsort in tools/video_retrieval.py:174-186.  The [Nq, Ng] similarity panel goes to the workspace, then a
// per-row top-k with ties going to the lower gallery index (== stable argsort of the distances).  Panel: split-bf16
// tcgen05 kernel when d % 64 == 0 (sim_tc.cu); this file keeps the API and the CUDA-core fp32 GEMMs for other widths.
#include "gca_common.cuh"

namespace gca {

__global__ void __launch_bounds__(128)
row_inv_norm_kernel(const float* __restrict__ x, int n, int d, int normalize, float* __restrict__ inv)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = blockIdx.x * 4 + warp;
    if (row >= n) return;
    float a = 0.f;
    if (normalize) {
        const float* r = x + (size_t)row * d;
        for (int c = lane; c < d; c += 32) { const float v = __ldg(r + c); a = fmaf(v, v, a); }
        a = warp_sum(a);
        a = 1.f / fmaxf(sqrtf(a), 1e-12f);
    } else {
        a = 1.f;
    }
    if (lane == 0) inv[row] = a;
}

// C[i][j] = invq[i] * invg[j] * sum_c Q[i][c] G[j][c];  128x128 tile, 256 threads, 8x8 register micro-tiles (two 4-wide
// halves 64 apart so shared-memory reads are conflict-free 128-bit loads), 16-wide k-chunks, register-prefetched double buffer
constexpr int ST_BM = 128, ST_BN = 128, ST_BK = 16;
__global__ void __launch_bounds__(256)
sim_gemm_kernel(const float* __restrict__ Q, const float* __restrict__ G, int Nq, int Ng, int d,
                const float* __restrict__ invq, const float* __restrict__ invg, float* __restrict__ C, long long ldc)
{
    __shared__ __align__(16) float As[2][ST_BK][ST_BM];
    __shared__ __align__(16) float Bs[2][ST_BK][ST_BN];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int i0 = blockIdx.y * ST_BM, j0 = blockIdx.x * ST_BN;
    // global -> register staging: each thread moves 2 float4 of A and 2 of B per chunk (row = tid / 4 (+64), 4 k-values)
    const int lr = tid >> 2, lk = (tid & 3) * 4;
    float4 ra[2], rb[2];
    auto gload = [&](int c0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = lr + 64 * h;
            ra[h] = make_float4(0.f, 0.f, 0.f, 0.f); rb[h] = ra[h];
            if (c0 + lk < d) {                                        // d % 4 == 0 on this path
                if (i0 + r < Nq) ra[h] = __ldg(reinterpret_cast<const float4*>(Q + (size_t)(i0 + r) * d + c0 + lk));
                if (j0 + r < Ng) rb[h] = __ldg(reinterpret_cast<const float4*>(G + (size_t)(j0 + r) * d + c0 + lk));
            }
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = lr + 64 * h;
            As[buf][lk][r] = ra[h].x; As[buf][lk + 1][r] = ra[h].y; As[buf][lk + 2][r] = ra[h].z; As[buf][lk + 3][r] = ra[h].w;
            Bs[buf][lk][r] = rb[h].x; Bs[buf][lk + 1][r] = rb[h].y; Bs[buf][lk + 2][r] = rb[h].z; Bs[buf][lk + 3][r] = rb[h].w;
        }
    };
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    gload(0);
    sstore(0);
    __syncthreads();
    int buf = 0;
    for (int c0 = 0; c0 < d; c0 += ST_BK) {
        if (c0 + ST_BK < d) gload(c0 + ST_BK);                       // next chunk in flight while this one is multiplied
#pragma unroll
        for (int c = 0; c < ST_BK; ++c) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][c][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][c][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][c][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][c][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (c0 + ST_BK < d) {
            sstore(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = i0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
        if (r >= Nq) continue;
        const float sq = invq[r];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int cidx = j0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
            if (cidx < Ng) C[(size_t)r * ldc + cidx] = acc[i][j] * sq * invg[cidx];
        }
    }
}

// fallback for d % 4 != 0: 64x64 tile, scalar loads
constexpr int SS_BM = 64, SS_BN = 64, SS_BK = 16;
__global__ void __launch_bounds__(256)
sim_gemm_small_kernel(const float* __restrict__ Q, const float* __restrict__ G, int Nq, int Ng, int d,
                      const float* __restrict__ invq, const float* __restrict__ invg, float* __restrict__ C, long long ldc)
{
    __shared__ float As[SS_BK][SS_BM + 1];
    __shared__ float Bs[SS_BK][SS_BN + 1];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int i0 = blockIdx.y * SS_BM, j0 = blockIdx.x * SS_BN;
    float acc[4][4] = {};
    for (int c0 = 0; c0 < d; c0 += SS_BK) {
        for (int e = tid; e < SS_BM * SS_BK; e += 256) {
            const int r = e / SS_BK, c = e % SS_BK;
            As[c][r] = (i0 + r < Nq && c0 + c < d) ? __ldg(Q + (size_t)(i0 + r) * d + c0 + c) : 0.f;
            Bs[c][r] = (j0 + r < Ng && c0 + c < d) ? __ldg(G + (size_t)(j0 + r) * d + c0 + c) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < SS_BK; ++c) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[c][ty * 4 + i]; b[i] = Bs[c][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = i0 + ty * 4 + i;
        if (r >= Nq) continue;
        const float sq = invq[r];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int cidx = j0 + tx * 4 + j;
            if (cidx < Ng) C[(size_t)r * ldc + cidx] = acc[i][j] * sq * invg[cidx];
        }
    }
}

}  // namespace gca

namespace gca {
// sim_tc.cu: split-bf16 tcgen05 panel + two-pass row top-k
bool sim_tc_supported(int d);
size_t sim_tc_pieces_bytes(int Nq, int Ng, int d);
int sim_tc_panel(const float* queries, const float* gallery, int Nq, int Ng, int d, int normalize, float* S, long long ldS,
                 void* pieces, float* tilemax, cudaStream_t st);
int sim_tc_col_tiles(int Ng);
int sim_topk_rows(float* S, int Nq, int Ng, long long ldS, int k, int* idx_out, float* val_out, const float* tilemax, int ntn,
                  cudaStream_t st);
}

static long long sim_ld(int Ng) { return ((long long)Ng + 3) / 4 * 4; }       // panel rows stay 16-byte aligned

extern "C" size_t gca_sim_topk_workspace_bytes(int Nq, int Ng, int d, int k)
{
    (void)k;
    if (Nq <= 0 || Ng <= 0 || d <= 0) return 0;
    const size_t panel = gca::align_up((size_t)Nq * sim_ld(Ng) * sizeof(float), 1024);
    const size_t norms = gca::align_up((size_t)(Nq + Ng) * sizeof(float), 256);
    const size_t pieces = gca::sim_tc_supported(d) ? gca::sim_tc_pieces_bytes(Nq, Ng, d) : 0;
    const size_t tmax = gca::sim_tc_supported(d) ? gca::align_up((size_t)Nq * gca::sim_tc_col_tiles(Ng) * sizeof(float), 256) : 0;
    return panel + (pieces > norms ? pieces : norms) + tmax;
}

extern "C" int gca_sim_topk(const float* queries, const float* gallery, int Nq, int Ng, int d, int k, int normalize,
                            int* idx_out, float* val_out, void* workspace, size_t workspace_bytes, void* stream)
{
    using namespace gca;
    GCA_CHECK_ARG(queries && gallery && idx_out, "gca_sim_topk: null pointer");
    GCA_CHECK_ARG(Nq >= 1 && Ng >= 1 && d >= 1, "gca_sim_topk: bad sizes");
    GCA_CHECK_ARG(k >= 1 && k <= 64 && k <= Ng, "gca_sim_topk: k=%d must be in [1, min(64, Ng)]", k);
    if (!workspace || workspace_bytes < gca_sim_topk_workspace_bytes(Nq, Ng, d, k))
        return set_err(GCA_ERR_WORKSPACE, "gca_sim_topk: workspace of %zu bytes needed", gca_sim_topk_workspace_bytes(Nq, Ng, d, k));
    if (sm_count_cached() < 1) return set_err(GCA_ERR_CUDA, "no CUDA device available (this library has no CPU path)");
    cudaStream_t st = (cudaStream_t)stream;
    float* sim = (float*)workspace;
    const long long ld = sim_ld(Ng);
    char* tail = (char*)workspace + align_up((size_t)Nq * ld * sizeof(float), 1024);
    if (sim_tc_supported(d)) {
        // tensor-core path: exact 3-way bf16 split of the normalised rows, 6 MMAs per product term (sim_tc.cu)
        float* tilemax = (float*)(tail + sim_tc_pieces_bytes(Nq, Ng, d));
        int rc = sim_tc_panel(queries, gallery, Nq, Ng, d, normalize, sim, ld, tail, tilemax, st);
        if (rc != GCA_OK) return rc;
        return sim_topk_rows(sim, Nq, Ng, ld, k, idx_out, val_out, tilemax, sim_tc_col_tiles(Ng), st);
    }
    float* invq = (float*)tail;
    float* invg = invq + Nq;
    row_inv_norm_kernel<<<(Nq + 3) / 4, 128, 0, st>>>(queries, Nq, d, normalize, invq);
    row_inv_norm_kernel<<<(Ng + 3) / 4, 128, 0, st>>>(gallery, Ng, d, normalize, invg);
    if (d % 4 == 0) {
        dim3 grid((Ng + ST_BN - 1) / ST_BN, (Nq + ST_BM - 1) / ST_BM);
        sim_gemm_kernel<<<grid, 256, 0, st>>>(queries, gallery, Nq, Ng, d, invq, invg, sim, ld);
    } else {
        dim3 grid((Ng + SS_BN - 1) / SS_BN, (Nq + SS_BM - 1) / SS_BM);
        sim_gemm_small_kernel<<<grid, 256, 0, st>>>(queries, gallery, Nq, Ng, d, invq, invg, sim, ld);
    }
    GCA_LAUNCH_CHECK("sim_topk kernels");
    count_launch(3);
    return sim_topk_rows(sim, Nq, Ng, ld, k, idx_out, val_out, nullptr, 0, st);
}
