// K1: InfoNCE stream on the 5th-generation tensor cores (GCA_ALGO_TCGEN05): bf16 queue, d = 128.
//
// One CTA = 128 query rows x a contiguous range of 128-key queue tiles; grid = (K-splits, row blocks) ~ one CTA/SM.
// It is a flash-attention forward whose keys and values are the same queue tile:
//     S = q Q^T        tcgen05.mma, A = q (bf16 smem tile), B = queue tile in smem, both K-major
//     p = 2^(S c - m)  one thread per query row, online max with lazy rescale, row sum, rank count
//     O += P Q         tcgen05.mma, A = P (bf16, written back over S in TMEM), B = the SAME smem tile, MN-major
// so a queue tile is fetched once (TMA, 128-byte swizzle, 4-stage mbarrier ring) and the [B, K] logits never
// leave the SM.  q is converted once per CTA into a swizzled smem tile (A operand).  Warp roles: warps 0-3 softmax (TMEM lane quarter = warp), warp 4 TMA producer, warp 5 MMA
// issuer + TMEM allocator.  Per-split partials (max, sum, count, O) go to the workspace; finalize.cu merges them
// in a fixed order.  Replaces mem_moco.py:36-46 + criterion.py:44 + autograd(mm) of the reference.
#include "gca_common.cuh"
#include "infonce_params.cuh"
#include "tc_ptx.cuh"
#include <stdlib.h>

namespace gca {

constexpr int TC_BM = 128, TC_BN = 128, TC_D = 128;
constexpr int TC_STAGES = 4;
constexpr int TC_STAGE_BYTES = TC_BN * TC_D * 2;          // 32 KB: two [128 keys][64 features] swizzled boxes
constexpr int TC_HALF_BYTES = TC_STAGE_BYTES / 2;
constexpr int TC_THREADS = 192;
constexpr uint32_t TM_COLS = 512;
constexpr uint32_t TM_S0 = 0, TM_S1 = 128, TM_O = 256;   // TMEM column map (fp32 S double buffer, fp32 O)
constexpr float TC_RESCALE_LOG2 = 8.f;                    // rescale O only when the row max grows by > 2^8
constexpr int TC_OST_STRIDE = 132;                        // fp32 O staging row (128 + 4 floats): conflict-free row writes
constexpr size_t TC_QTILE_BYTES = (size_t)TC_BM * TC_D * 2;   // bf16 q block, same swizzled layout as a queue tile
constexpr size_t TC_SMEM_BYTES = 1024 + (size_t)TC_STAGES * TC_STAGE_BYTES + TC_QTILE_BYTES + 512;
static_assert((size_t)TC_BM * TC_OST_STRIDE * 4 <= (size_t)TC_STAGES * TC_STAGE_BYTES, "O staging must fit in the tile ring");

struct TcDebug { uint32_t lbo1, sbo1, kstep1, lbo2, sbo2, kstep2; };

struct TcBarriers {
    uint64_t full[TC_STAGES];     // TMA landed a queue tile
    uint64_t empty[TC_STAGES];    // both MMAs that read the tile have completed
    uint64_t s_full[2];           // S = q Q^T of a tile is in TMEM
    uint64_t p_full[2];           // softmax finished with S (and wrote P)
    uint64_t o_done;              // one phase per completed O += P Q
    uint64_t acc_final;           // last O += P Q completed
    uint64_t q_ready;             // the bf16 q block is in shared memory (A operand of S = q Q^T)
    uint32_t tmem_base;
};

template <bool kWantAcc, bool kFixedMax>
__global__ void __launch_bounds__(TC_THREADS, 1)
infonce_tc_kernel(const __grid_constant__ CUtensorMap tmap, const InfoNceStreamParams P, const TcDebug dbg)
{
    using namespace ptx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = blockIdx.x, row0 = blockIdx.y * TC_BM;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* stages = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* qtile = stages + (size_t)TC_STAGES * TC_STAGE_BYTES;         // bf16 q block, two swizzled [128][64] boxes
    TcBarriers* bar = reinterpret_cast<TcBarriers*>(qtile + TC_QTILE_BYTES);
    const int nrows = (P.B - row0 < TC_BM) ? (P.B - row0) : TC_BM;        // valid query rows of this block

    const long long ntiles = (P.K + TC_BN - 1) / TC_BN;
    const long long t_begin = ntiles * split / P.nsplit, t_end = ntiles * (split + 1) / P.nsplit;
    const int n = (int)(t_end - t_begin);

    if (split == 0 && blockIdx.y == 0 && threadIdx.x == 0) *P.counter = 0u;      // re-arm the finalize ticket

    if (warp == 4 && lane == 0) {
        prefetch_tmap(&tmap);
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&bar->full[s], 1); mbar_init(&bar->empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&bar->s_full[b], 1); mbar_init(&bar->p_full[b], 128); }
        mbar_init(&bar->o_done, 1);
        mbar_init(&bar->acc_final, 1);
        mbar_init(&bar->q_ready, 128);
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc<TM_COLS>(&bar->tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bar->tmem_base;

    if (warp < 4) {
        // =============================================================== softmax warps: one thread per query row
        const int row = row0 + warp * 32 + lane;
        const bool valid = row < P.B;
        const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
        const float c2 = P.inv_T * 1.4426950408889634f;                 // logits -> log2 domain

        // Prologue.  Each warp converts ITS 32 query rows with coalesced 128-bit loads (lane = 4 features): the
        // positive logit q.k in fp32 (fixed shuffle order: every split computes the same bits) and q -> bf16 written in
        // the 128-byte-swizzled K-major layout of a queue tile, so S = q Q^T is a plain smem x smem tcgen05.mma.
        float pos_dot = 0.f;
        {
            const float4* qg = reinterpret_cast<const float4*>(P.q + (size_t)row0 * TC_D);
            const float4* kg = reinterpret_cast<const float4*>(P.k + (size_t)row0 * TC_D);
            const int sub = lane >> 4, chunk = (lane & 15) >> 1, half = lane & 1;    // 64-feature box, 16-byte chunk, 8-byte half
#pragma unroll
            for (int batch = 0; batch < 2; ++batch) {
                float4 a[16], b[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int r = warp * 32 + batch * 16 + u;
                    a[u] = make_float4(0.f, 0.f, 0.f, 0.f); b[u] = a[u];
                    if (r < nrows) { a[u] = __ldg(qg + r * 32 + lane); b[u] = __ldg(kg + r * 32 + lane); }
                }
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int r = warp * 32 + batch * 16 + u;
                    float dsum = fmaf(a[u].x, b[u].x, fmaf(a[u].y, b[u].y, fmaf(a[u].z, b[u].z, a[u].w * b[u].w)));
                    dsum = warp_sum(dsum);
                    if (lane == batch * 16 + u) pos_dot = dsum;
                    uint2 pk2;
                    pk2.x = pack_bf16(a[u].x, a[u].y);
                    pk2.y = pack_bf16(a[u].z, a[u].w);
                    *reinterpret_cast<uint2*>(qtile + sub * TC_HALF_BYTES + r * 128 + ((chunk ^ (r & 7)) << 4) + half * 8) = pk2;
                }
            }
            fence_proxy_async();                               // generic-proxy smem writes -> visible to the tensor core
            mbar_arrive(&bar->q_ready);
        }
        const float pos_nat = pos_dot * P.inv_T;
        if (split == 0 && valid) {
            if (P.pos_out) P.pos_out[row] = pos_nat;
            if (P.logits_out) P.logits_out[(size_t)row * P.ld_logits] = pos_nat;
        }

        float m_run = kFixedMax ? (valid ? P.lse_fixed[row] * 1.4426950408889634f : 0.f) : -INFINITY;   // log2 domain
        float s_run = 0.f;
        int cnt = 0;

        for (int i = 0; i < n; ++i) {
            const int buf = i & 1;
            const uint32_t s_addr = lane_addr + (buf ? TM_S1 : TM_S0);
            mbar_wait(&bar->s_full[buf], (i >> 1) & 1);
            tc_fence_after();
            uint32_t sr[128];
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld32(s_addr + 32 * c, sr + 32 * c);
            tc_wait_ld();
            float* sv = reinterpret_cast<float*>(sr);

            const long long key0 = (t_begin + i) * TC_BN;
            const int nvalid = (P.K - key0 < TC_BN) ? (int)(P.K - key0) : TC_BN;
            if (P.logits_out && valid) {                                   // parity / debug path only
                float* dst = P.logits_out + (size_t)row * P.ld_logits + 1 + key0;
#pragma unroll
                for (int j = 0; j < TC_BN; ++j) if (j < nvalid) dst[j] = sv[j] * P.inv_T;
            }
            if (nvalid < TC_BN) {
#pragma unroll
                for (int j = 0; j < TC_BN; ++j) if (j >= nvalid) sv[j] = -INFINITY;   // TMA zero-filled rows past K
            }
            float tmax = -INFINITY;
            int c = 0;
#pragma unroll
            for (int j = 0; j < TC_BN; j += 2) {
                tmax = max3(tmax, sv[j], sv[j + 1]);
                c += (sv[j] > pos_dot) ? 1 : 0;
                c += (sv[j + 1] > pos_dot) ? 1 : 0;
            }
            cnt += c;

            if (!kFixedMax) {
                const float xm = tmax * c2;
                if (i == 0) {
                    m_run = xm;
                } else {
                    const bool need = xm > m_run + TC_RESCALE_LOG2;
                    if (__any_sync(0xffffffffu, need)) {                    // warp-uniform: TMEM ld/st are collective
                        const float m_new = need ? xm : m_run;
                        const float sc = ex2(m_run - m_new);
                        s_run *= sc;
                        m_run = m_new;
                        if (kWantAcc) {
                            mbar_wait(&bar->o_done, (i - 1) & 1);          // O += P Q of tile i-1 has landed
                            tc_fence_after();
#pragma unroll
                            for (int ch = 0; ch < 4; ++ch) {
                                uint32_t t[32];
                                tmem_ld32(lane_addr + TM_O + 32 * ch, t);
                                tc_wait_ld();
#pragma unroll
                                for (int j = 0; j < 32; ++j) t[j] = __float_as_uint(__uint_as_float(t[j]) * sc);
                                tmem_st32(lane_addr + TM_O + 32 * ch, t);
                            }
                            tc_wait_st();
                        }
                    }
                }
            }
            const float neg_m = -m_run;
            float rs = 0.f;
#pragma unroll
            for (int j = 0; j < TC_BN; ++j) { const float p = ex2(fmaf(sv[j], c2, neg_m)); rs += p; sv[j] = p; }
            s_run += rs;
            if (kWantAcc) {
                uint32_t pk[64];
#pragma unroll
                for (int j = 0; j < 64; ++j) pk[j] = pack_bf16(sv[2 * j], sv[2 * j + 1]);
                tmem_st32(s_addr, pk);                                     // P (bf16) overwrites the first half of S
                tmem_st32(s_addr + 32, pk + 32);
                tc_wait_st();
            }
            tc_fence_before();
            mbar_arrive(&bar->p_full[buf]);
        }

        // split partials
        const size_t po = (size_t)split * P.Bpad + row;
        P.part_max[po] = kFixedMax ? 0.f : m_run * 0.6931471805599453f;  // back to natural-log units
        P.part_sum[po] = s_run;
        P.part_cnt[po] = cnt;
        if (kWantAcc) {
            // O row (thread-owned in TMEM) -> padded smem staging (the tile ring is idle now) -> coalesced 512-byte rows
            mbar_wait(&bar->acc_final, 0);
            tc_fence_after();
            float* ost = reinterpret_cast<float*>(stages);
            float4* mine = reinterpret_cast<float4*>(ost + (size_t)(warp * 32 + lane) * TC_OST_STRIDE);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                uint32_t t[32];
                tmem_ld32(lane_addr + TM_O + 32 * ch, t);
                tc_wait_ld();
#pragma unroll
                for (int v = 0; v < 8; ++v)
                    mine[ch * 8 + v] = make_float4(__uint_as_float(t[4 * v]), __uint_as_float(t[4 * v + 1]),
                                                   __uint_as_float(t[4 * v + 2]), __uint_as_float(t[4 * v + 3]));
            }
            __syncwarp();                                   // each warp drains exactly the 32 rows it staged
            float4* dst = reinterpret_cast<float4*>(P.part_acc + ((size_t)split * P.Bpad + row0 + warp * 32) * TC_D);
#pragma unroll 8
            for (int rr = 0; rr < 32; ++rr) {
                if (warp * 32 + rr < nrows)
                    dst[rr * 32 + lane] = *reinterpret_cast<const float4*>(ost + (size_t)(warp * 32 + rr) * TC_OST_STRIDE + lane * 4);
            }
        }
    } else if (warp == 4) {
        // =============================================================== TMA producer
        if (lane == 0) {
            for (int i = 0; i < n; ++i) {
                const int stage = i % TC_STAGES;
                if (i >= TC_STAGES) mbar_wait(&bar->empty[stage], ((i / TC_STAGES) - 1) & 1);
                uint8_t* dst = stages + (size_t)stage * TC_STAGE_BYTES;
                const int key0 = (int)((t_begin + i) * TC_BN);
                mbar_arrive_expect_tx(&bar->full[stage], TC_STAGE_BYTES);
                tma_load_2d(dst, &tmap, &bar->full[stage], 0, key0);               // features  0..63
                tma_load_2d(dst + TC_HALF_BYTES, &tmap, &bar->full[stage], 64, key0);   // features 64..127
            }
        }
    } else {
        // =============================================================== MMA issuer (one thread)
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc_bf16(TC_BM, TC_BN, 0, 0);   // S: B = tile, K-major  (N = keys)
            constexpr uint32_t idesc_o = make_idesc_bf16(TC_BM, TC_D, 0, 1);    // O: B = tile, MN-major (N = features)
            mbar_wait(&bar->q_ready, 0);
            tc_fence_after();
            const uint32_t qbase = smem_u32(qtile);
            for (int i = 0; i <= n; ++i) {
                if (i < n) {
                    const int stage = i % TC_STAGES;
                    mbar_wait(&bar->full[stage], (i / TC_STAGES) & 1);
                    if (!kWantAcc && i >= 2) mbar_wait(&bar->p_full[i & 1], ((i - 2) >> 1) & 1);   // S buffer drained
                    tc_fence_after();
                    const uint32_t sbase = smem_u32(stages + (size_t)stage * TC_STAGE_BYTES);
                    const uint32_t d_tmem = tmem + ((i & 1) ? TM_S1 : TM_S0);
#pragma unroll
                    for (int kk = 0; kk < TC_D / 16; ++kk) {
                        // 16 features per MMA: 32 bytes along the swizzled 128-byte row, next box after 4 steps
                        const uint64_t bd = make_smem_desc_sw128(sbase + (kk >> 2) * TC_HALF_BYTES + (kk & 3) * dbg.kstep1,
                                                                 dbg.lbo1, dbg.sbo1);
                        const uint64_t ad = make_smem_desc_sw128(qbase + (kk >> 2) * TC_HALF_BYTES + (kk & 3) * dbg.kstep1,
                                                                 dbg.lbo1, dbg.sbo1);
                        mma_ss(d_tmem, ad, bd, idesc_s, kk > 0);
                    }
                    tc_commit(&bar->s_full[i & 1]);
                    if (!kWantAcc) tc_commit(&bar->empty[stage]);
                }
                if (kWantAcc && i >= 1) {
                    const int j = i - 1, stage = j % TC_STAGES;
                    mbar_wait(&bar->p_full[j & 1], (j >> 1) & 1);
                    tc_fence_after();
                    const uint32_t sbase = smem_u32(stages + (size_t)stage * TC_STAGE_BYTES);
                    const uint32_t p_tmem = tmem + ((j & 1) ? TM_S1 : TM_S0);
#pragma unroll
                    for (int kk = 0; kk < TC_BN / 16; ++kk) {
                        // 16 keys per MMA = two 8-row swizzle atoms (2 KB); LBO = next 64-feature box, SBO = next 8 keys
                        const uint64_t bd = make_smem_desc_sw128(sbase + kk * dbg.kstep2, dbg.lbo2, dbg.sbo2);
                        mma_ts(tmem + TM_O, p_tmem + kk * 8, bd, idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
                    }
                    tc_commit(&bar->empty[stage]);
                    tc_commit(&bar->o_done);
                    if (j == n - 1) tc_commit(&bar->acc_final);
                }
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc<TM_COLS>(tmem);
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

struct TmapCache { const void* ptr; long long K; CUtensorMap map; bool ok; };

static int get_queue_tmap(const void* queue, long long K, CUtensorMap* out)
{
    static thread_local TmapCache cache[4] = {};
    static thread_local int next = 0;
    for (int i = 0; i < 4; ++i)
        if (cache[i].ok && cache[i].ptr == queue && cache[i].K == K) { *out = cache[i].map; return GCA_OK; }
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_err(GCA_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {(cuuint64_t)TC_D, (cuuint64_t)K};
    cuuint64_t strides[1] = {(cuuint64_t)TC_D * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)TC_BN};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMap m;
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(queue), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_err(GCA_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    cache[next] = TmapCache{queue, K, m, true};
    next = (next + 1) & 3;
    *out = m;
    return GCA_OK;
}

int infonce_tc_nsplit(int B, long long K)
{
    const int nblk = infonce_bpad(B) / TC_BM;
    const long long ntiles = (K + TC_BN - 1) / TC_BN;
    int ns = sm_count_cached() / nblk;
    if (ns < 1) ns = 1;
    if (ns > ntiles) ns = (int)ntiles;
    return ns;
}

static TcDebug tc_debug_knobs()
{
    // canonical values; GCA_TC_DESC="lbo1,sbo1,kstep1,lbo2,sbo2,kstep2" overrides them (bring-up aid only)
    TcDebug d{16, 1024, 32, (uint32_t)TC_HALF_BYTES, 1024, 2048};
    const char* e = getenv("GCA_TC_DESC");
    if (e) {
        unsigned v[6];
        if (sscanf(e, "%u,%u,%u,%u,%u,%u", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5]) == 6)
            d = TcDebug{v[0], v[1], v[2], v[3], v[4], v[5]};
    }
    return d;
}

int infonce_tc_launch(const InfoNceStreamParams& P, bool fixed_max, cudaStream_t st)
{
    if (P.d != TC_D) return set_err(GCA_ERR_UNSUPPORTED, "tcgen05 InfoNCE kernel needs d == %d (got %d)", TC_D, P.d);
    if (P.K >= (1ll << 31) - TC_BN) return set_err(GCA_ERR_UNSUPPORTED, "tcgen05 InfoNCE kernel: K too large");
    if ((reinterpret_cast<uintptr_t>(P.queue) & 15) != 0) return set_err(GCA_ERR_BAD_ARG, "queue must be 16-byte aligned");
    CUtensorMap tmap;
    int rc = get_queue_tmap(P.queue, P.K, &tmap);
    if (rc != GCA_OK) return rc;
    const TcDebug dbg = tc_debug_knobs();
    const bool want_acc = P.part_acc != nullptr;
    dim3 grid(P.nsplit, P.Bpad / TC_BM);
#define GCA_TC_LAUNCH(ACC, FIX) do { \
        auto kern = infonce_tc_kernel<ACC, FIX>; \
        GCA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES)); \
        kern<<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(tmap, P, dbg); } while (0)
    if (want_acc) { if (fixed_max) GCA_TC_LAUNCH(true, true); else GCA_TC_LAUNCH(true, false); }
    else          { if (fixed_max) GCA_TC_LAUNCH(false, true); else GCA_TC_LAUNCH(false, false); }
#undef GCA_TC_LAUNCH
    GCA_LAUNCH_CHECK("infonce_tc_kernel");
    return GCA_OK;
}

}  // namespace gca
