// K7 (tensor-core path): retrieval similarity panel S = Q' G'^T at fp32-grade accuracy on tcgen05, plus a two-pass
// per-row top-k.  Replaces sklearn cosine_distances + full-row argsort of tools/video_retrieval.py:174-186.
//
// fp32 accuracy from bf16 MMAs: every (normalised) fp32 operand is split EXACTLY into three bf16 pieces
//     x = h + m + l,   h = bf16(x), m = bf16(x - h), l = bf16(x - h - m)        (8 + 8 + 8 mantissa bits)
// and the six piece products whose magnitude is >= 2^-24 of the full product are accumulated in the fp32 TMEM
// accumulator (l.h, h.l, m.m, m.h, h.m, h.h -- smallest first); the three dropped ones (m.l, l.m, l.l) are below fp32
// rounding.  6 x 51.6 GFLOP of bf16 MMA instead of 51.6 GFLOP of CUDA-core FFMA.
//
// Kernel: persistent, one CTA per SM, 128 x 128 output tiles, K chunks of 64 features.  Warp 4 = TMA producer (2 stages
// x 6 swizzled [128 x 64] bf16 boxes = 96 KB per stage), warp 5 = MMA issuer (24 tcgen05.mma per stage into one of two
// 128-column TMEM accumulators), warps 0-3 = epilogue (tcgen05.ld -> fp32 rows of the panel) overlapping the next tile.
#include "gca_common.cuh"
#include "tc_ptx.cuh"

namespace gca {

using namespace ptx;

constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 64;
constexpr int SG_STAGES = 2;
constexpr int SG_PIECE_BYTES = SG_BM * SG_BK * 2;            // 16 KB: one swizzled [128 rows x 64 features] bf16 box
constexpr int SG_STAGE_BYTES = 6 * SG_PIECE_BYTES;           // A: h, m, l   then   B: h, m, l
constexpr int SG_THREADS = 192;
constexpr size_t SG_SMEM_BYTES = 1024 + (size_t)SG_STAGES * SG_STAGE_BYTES + 256;
constexpr uint32_t SG_TM_COLS = 256;                         // two fp32 accumulators of 128 columns

struct SgBarriers {
    uint64_t full[SG_STAGES], empty[SG_STAGES];
    uint64_t acc_full[2], acc_empty[2];
    uint32_t tmem_base;
};

// one warp per row: optional L2 normalisation (x / max(||x||, 1e-12), IEEE division: this file is built without fast-math),
// then the exact three-way bf16 split.  pieces: [3][n][d] bf16.
__global__ void __launch_bounds__(128)
sim_split_kernel(const float* __restrict__ x, int n, int d, int normalize, __nv_bfloat16* __restrict__ pieces)
{
    const int lane = threadIdx.x & 31, row = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= n) return;
    const float4* r4 = reinterpret_cast<const float4*>(x + (size_t)row * d);
    const int d4 = d / 4;
    float nrm = 1.f;
    if (normalize) {
        float a = 0.f;
        for (int c = lane; c < d4; c += 32) {
            const float4 v = __ldg(r4 + c);
            a = fmaf(v.x, v.x, a); a = fmaf(v.y, v.y, a); a = fmaf(v.z, v.z, a); a = fmaf(v.w, v.w, a);
        }
        a = warp_sum(a);
        nrm = fmaxf(sqrtf(a), 1e-12f);
    }
    const size_t plane = (size_t)n * d;
    for (int c = lane; c < d4; c += 32) {
        const float4 v = __ldg(r4 + c);
        const float xs[4] = {v.x / nrm, v.y / nrm, v.z / nrm, v.w / nrm};
        __nv_bfloat16 h[4], m[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            h[i] = __float2bfloat16_rn(xs[i]);
            const float r1 = xs[i] - __bfloat162float(h[i]);           // exact
            m[i] = __float2bfloat16_rn(r1);
            const float r2 = r1 - __bfloat162float(m[i]);              // exact
            l[i] = __float2bfloat16_rn(r2);
        }
        const size_t o = (size_t)row * d + 4 * c;
        *reinterpret_cast<uint2*>(pieces + o) = *reinterpret_cast<const uint2*>(h);
        *reinterpret_cast<uint2*>(pieces + plane + o) = *reinterpret_cast<const uint2*>(m);
        *reinterpret_cast<uint2*>(pieces + 2 * plane + o) = *reinterpret_cast<const uint2*>(l);
    }
}

__global__ void __launch_bounds__(SG_THREADS, 1)
sim_tc_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap,
              int Nq, int Ng, int d, float* __restrict__ S, long long ldS, float* __restrict__ tilemax /* [Nq, ntn] */)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* stages = smem;
    SgBarriers* bar = reinterpret_cast<SgBarriers*>(smem + (size_t)SG_STAGES * SG_STAGE_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntm = (Nq + SG_BM - 1) / SG_BM, ntn = (Ng + SG_BN - 1) / SG_BN;
    const int ntiles = ntm * ntn;
    const int nk = d / SG_BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < SG_STAGES; ++s) { mbar_init(&bar->full[s], 1); mbar_init(&bar->empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&bar->acc_full[a], 1); mbar_init(&bar->acc_empty[a], 128); }
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc<SG_TM_COLS>(&bar->tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bar->tmem_base;

    if (warp == 4) {
        // ------------------------------------------------------------------ TMA producer
        if (elect_one()) {
            prefetch_tmap(&amap);
            prefetch_tmap(&bmap);
            int it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int tm = tile % ntm, tn = tile / ntm;
                for (int kc = 0; kc < nk; ++kc, ++it) {
                    const int s = it % SG_STAGES;
                    if (it >= SG_STAGES) mbar_wait(&bar->empty[s], ((it / SG_STAGES) - 1) & 1);
                    uint8_t* dst = stages + (size_t)s * SG_STAGE_BYTES;
                    mbar_arrive_expect_tx(&bar->full[s], (uint32_t)SG_STAGE_BYTES);
#pragma unroll
                    for (int p = 0; p < 3; ++p) {        // pieces are stacked along the row axis of each tensor map
                        tma_load_2d(dst + p * SG_PIECE_BYTES, &amap, &bar->full[s], kc * SG_BK, p * Nq + tm * SG_BM);
                        tma_load_2d(dst + (3 + p) * SG_PIECE_BYTES, &bmap, &bar->full[s], kc * SG_BK, p * Ng + tn * SG_BN);
                    }
                }
            }
        }
    } else if (warp == 5) {
        // ------------------------------------------------------------------ MMA issuer (uniform loop, one elected lane issues)
        const bool leader = elect_one();
        constexpr uint32_t idesc = make_idesc_bf16(SG_BM, SG_BN, 0, 0);
        int it = 0, ti = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
            const int ab = ti & 1;
            if (ti >= 2) { mbar_wait(&bar->acc_empty[ab], ((ti >> 1) - 1) & 1); tc_fence_after(); }
            const uint32_t d_tmem = tmem + ab * 128;
            for (int kc = 0; kc < nk; ++kc, ++it) {
                const int s = it % SG_STAGES;
                mbar_wait(&bar->full[s], (it / SG_STAGES) & 1);
                tc_fence_after();
                const uint32_t sbase = smem_u32(stages + (size_t)s * SG_STAGE_BYTES);
                if (leader) {
                    // (A piece, B piece), smallest products first
                    const int pa[6] = {2, 0, 1, 1, 0, 0}, pb[6] = {0, 2, 1, 0, 1, 0};
#pragma unroll
                    for (int c = 0; c < 6; ++c) {
                        const uint32_t abase = sbase + pa[c] * SG_PIECE_BYTES, bbase = sbase + (3 + pb[c]) * SG_PIECE_BYTES;
#pragma unroll
                        for (int kk = 0; kk < SG_BK / 16; ++kk)
                            mma_ss(d_tmem, make_smem_desc_sw128(abase + kk * 32, 16, 1024),
                                   make_smem_desc_sw128(bbase + kk * 32, 16, 1024), idesc, (kc > 0 || c > 0 || kk > 0) ? 1u : 0u);
                    }
                    tc_commit(&bar->empty[s]);
                    if (kc == nk - 1) tc_commit(&bar->acc_full[ab]);
                }
                __syncwarp();
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps 0-3: TMEM lane = tile row
        int ti = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
            const int tm = tile % ntm, tn = tile / ntm;
            const int ab = ti & 1;
            mbar_wait(&bar->acc_full[ab], (ti >> 1) & 1);
            tc_fence_after();
            const int gr = tm * SG_BM + warp * 32 + lane;
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + ab * 128;
            float* out = S + (size_t)gr * ldS + (size_t)tn * SG_BN;
            float rmax = -INFINITY;                      // max of this row over the tile's valid columns (top-k threshold input)
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                uint32_t r[32];
                tmem_ld32(taddr + ch * 32, r);
                tc_wait_ld();
                if (gr < Nq) {
                    const int c0 = tn * SG_BN + ch * 32;
#pragma unroll
                    for (int j = 0; j < 32; ++j) if (c0 + j < Ng) rmax = fmaxf(rmax, __uint_as_float(r[j]));
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        if (c0 + j + 3 < Ng) {
                            *reinterpret_cast<float4*>(out + ch * 32 + j) =
                                make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) if (c0 + j + e < Ng) out[ch * 32 + j + e] = __uint_as_float(r[j + e]);
                        }
                    }
                }
            }
            if (gr < Nq) tilemax[(size_t)gr * ntn + tn] = rmax;
            tc_fence_before();
            mbar_arrive(&bar->acc_empty[ab]);
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc<SG_TM_COLS>(tmem);
}

// ------------------------------------------------------------------------------------------------ top-k
__device__ __forceinline__ bool tk_better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

constexpr int TK_THREADS = 256, TK_CAP = 1024;

// One CTA per query row, two passes over the row:
//   1. maxima of up to 256 disjoint parts of the row (per-tile maxima from the panel kernel, or per-thread strided maxima):
//      the k-th largest of them is a lower bound `thr` of the row's k-th largest value (k distinct elements >= thr)
//   2. elements >= thr (about 1.5 k of them for unstructured data) go to a shared candidate list, which is rank-sorted by
//      (value desc, index asc) -- ties towards the lower gallery index like a stable argsort of the distances.
// More than TK_CAP candidates (massive ties): k rounds of block arg-max over the row instead (rows are scratch).
__global__ void __launch_bounds__(TK_THREADS)
row_topk_fast_kernel(float* __restrict__ sim, int Ng, long long ld, int k, int* __restrict__ idx_out, float* __restrict__ val_out,
                     const float* __restrict__ tilemax, int ntn)
{
    __shared__ float tmax[TK_THREADS];
    __shared__ float cv[TK_CAP];
    __shared__ int ci[TK_CAP];
    __shared__ int ccount;
    __shared__ float thr_s;
    __shared__ float wv[TK_THREADS / 32];
    __shared__ int wi[TK_THREADS / 32];
    __shared__ int win_idx;
    float* row = sim + (size_t)blockIdx.x * ld;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // step 1 input: maxima of disjoint parts of the row -- the panel kernel's per-tile row maxima when it wrote them and
    // there are at least k of them (saves a whole pass over the row), else each thread's strided maximum
    float mx = -INFINITY;
    int nparts = TK_THREADS;
    if (tilemax != nullptr && ntn >= k && ntn <= TK_THREADS) {
        nparts = ntn;
        if (tid < ntn) mx = __ldcg(tilemax + (size_t)blockIdx.x * ntn + tid);
    } else {
        for (int j = tid; j < Ng; j += TK_THREADS) mx = fmaxf(mx, __ldcg(row + j));
    }
    tmax[tid] = mx;
    if (tid == 0) ccount = 0;
    __syncthreads();
    if (tid < nparts) {
        int rank = 0;
        for (int o = 0; o < nparts; ++o) { const float ov = tmax[o]; rank += (ov > mx || (ov == mx && o < tid)) ? 1 : 0; }
        if (rank == k - 1) thr_s = mx;
    }
    __syncthreads();
    const float thr = thr_s;
    if ((ld & 3) == 0 && (reinterpret_cast<uintptr_t>(sim) & 15) == 0) {           // 128-bit loads (rows are 16-byte aligned)
        const int n4 = Ng >> 2;
        const float4* row4 = reinterpret_cast<const float4*>(row);
        for (int j4 = tid; j4 < n4; j4 += TK_THREADS) {
            const float4 v = __ldcg(row4 + j4);
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (vv[e] >= thr) {
                    const int p = atomicAdd(&ccount, 1);
                    if (p < TK_CAP) { cv[p] = vv[e]; ci[p] = 4 * j4 + e; }
                }
            }
        }
        for (int j = 4 * n4 + tid; j < Ng; j += TK_THREADS) {
            const float v = __ldcg(row + j);
            if (v >= thr) {
                const int p = atomicAdd(&ccount, 1);
                if (p < TK_CAP) { cv[p] = v; ci[p] = j; }
            }
        }
    } else {
        for (int j = tid; j < Ng; j += TK_THREADS) {
            const float v = __ldcg(row + j);
            if (v >= thr) {
                const int p = atomicAdd(&ccount, 1);
                if (p < TK_CAP) { cv[p] = v; ci[p] = j; }
            }
        }
    }
    __syncthreads();
    const int C = ccount;
    if (C <= TK_CAP) {
        for (int c = tid; c < C; c += TK_THREADS) {
            const float v = cv[c]; const int i = ci[c];
            int rank = 0;
            for (int o = 0; o < C; ++o) rank += tk_better(cv[o], ci[o], v, i) ? 1 : 0;
            if (rank < k) {
                idx_out[(size_t)blockIdx.x * k + rank] = i;
                if (val_out) val_out[(size_t)blockIdx.x * k + rank] = v;
            }
        }
        return;
    }
    // fallback: k rounds of block arg-max, the winner's owner retires it and rescans its elements
    float bv = -INFINITY; int bi = 0x7fffffff;
    for (int j = tid; j < Ng; j += TK_THREADS) { const float v = row[j]; if (tk_better(v, j, bv, bi)) { bv = v; bi = j; } }
    for (int r = 0; r < k; ++r) {
        float v = bv; int i = bi;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, i, o);
            if (tk_better(ov, oi, v, i)) { v = ov; i = oi; }
        }
        if (lane == 0) { wv[warp] = v; wi[warp] = i; }
        __syncthreads();
        if (tid == 0) {
            float fv = wv[0]; int fi = wi[0];
            for (int w = 1; w < TK_THREADS / 32; ++w) if (tk_better(wv[w], wi[w], fv, fi)) { fv = wv[w]; fi = wi[w]; }
            win_idx = fi;
            idx_out[(size_t)blockIdx.x * k + r] = (fi == 0x7fffffff) ? -1 : fi;
            if (val_out) val_out[(size_t)blockIdx.x * k + r] = fv;
        }
        __syncthreads();
        const int w = win_idx;
        if (w != 0x7fffffff && (w % TK_THREADS) == tid) {
            row[w] = -INFINITY;
            bv = -INFINITY; bi = 0x7fffffff;
            for (int j = tid; j < Ng; j += TK_THREADS) { const float x = row[j]; if (tk_better(x, j, bv, bi)) { bv = x; bi = j; } }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*SgEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int sg_tmap(const void* base, long long rows, int d, CUtensorMap* out)
{
    static SgEncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (SgEncodeFn)p;
    }
    if (!fn) return set_err(GCA_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)d * 2};
    cuuint32_t box[2] = {(cuuint32_t)SG_BK, (cuuint32_t)SG_BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_err(GCA_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return GCA_OK;
}

bool sim_tc_supported(int d) { return d >= SG_BK && d % SG_BK == 0; }

size_t sim_tc_pieces_bytes(int Nq, int Ng, int d) { return align_up((size_t)3 * ((size_t)Nq + Ng) * d * 2, 1024); }

// S[Nq, ldS] = normalised(Q) normalised(G)^T through the split-bf16 tcgen05 kernel; `pieces` is sim_tc_pieces_bytes() of scratch
int sim_tc_panel(const float* queries, const float* gallery, int Nq, int Ng, int d, int normalize, float* S, long long ldS,
                 void* pieces, float* tilemax, cudaStream_t st)
{
    __nv_bfloat16* pq = (__nv_bfloat16*)pieces;
    __nv_bfloat16* pg = pq + (size_t)3 * Nq * d;
    sim_split_kernel<<<(Nq + 3) / 4, 128, 0, st>>>(queries, Nq, d, normalize, pq);
    sim_split_kernel<<<(Ng + 3) / 4, 128, 0, st>>>(gallery, Ng, d, normalize, pg);
    GCA_LAUNCH_CHECK("sim_split_kernel");
    CUtensorMap amap, bmap;
    int rc = sg_tmap(pq, (long long)3 * Nq, d, &amap);
    if (rc != GCA_OK) return rc;
    rc = sg_tmap(pg, (long long)3 * Ng, d, &bmap);
    if (rc != GCA_OK) return rc;
    const int ntiles = ((Nq + SG_BM - 1) / SG_BM) * ((Ng + SG_BN - 1) / SG_BN);
    int grid = sm_count_cached();
    if (grid > ntiles) grid = ntiles;
    GCA_CUDA(cudaFuncSetAttribute(sim_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG_SMEM_BYTES));
    sim_tc_kernel<<<grid, SG_THREADS, SG_SMEM_BYTES, st>>>(amap, bmap, Nq, Ng, d, S, ldS, tilemax);
    GCA_LAUNCH_CHECK("sim_tc_kernel");
    count_launch(3);
    return GCA_OK;
}

int sim_tc_col_tiles(int Ng) { return (Ng + SG_BN - 1) / SG_BN; }

int sim_topk_rows(float* S, int Nq, int Ng, long long ldS, int k, int* idx_out, float* val_out, const float* tilemax, int ntn,
                  cudaStream_t st)
{
    row_topk_fast_kernel<<<Nq, TK_THREADS, 0, st>>>(S, Ng, ldS, k, idx_out, val_out, tilemax, ntn);
    GCA_LAUNCH_CHECK("row_topk_fast_kernel");
    count_launch(1);
    return GCA_OK;
}

}  // namespace gca
