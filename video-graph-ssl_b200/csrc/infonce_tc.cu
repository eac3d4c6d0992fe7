// K1: InfoNCE stream on the 5th-generation tensor cores (GCA_ALGO_TCGEN05): bf16 queue, d = 128.
//
// One CTA = 128 query rows x a contiguous range of 128-key queue tiles; grid = (K-splits, row blocks) ~ one CTA/SM.
// It is a flash-attention forward whose keys and values are the same queue tile:
//     S = q Q^T        tcgen05.mma, A = q (bf16 smem tile), B = queue tile in smem, both K-major
//     p = 2^(S c - m)  one thread per query row, online max with lazy rescale, row sum, rank count; two warpgroups
//                      ping-pong over alternate tiles with their own S / O buffers in TMEM
//     O += P Q         tcgen05.mma, A = P (bf16, written back over S in TMEM), B = the SAME smem tile, MN-major
// so a queue tile is fetched once (TMA, 128-byte swizzle, 4-stage mbarrier ring) and the [B, K] logits never
// leave the SM.  Per-split partials (max, sum, count, O) go to the workspace; finalize.cu merges them
// in a fixed order.  Replaces mem_moco.py:36-46 + criterion.py:44 + autograd(mm) of the reference.
#include "gca_common.cuh"
#include "infonce_params.cuh"
#include "tc_ptx.cuh"
#include "launch_plan.cuh"
#include <stdlib.h>

namespace gca {

constexpr int TC_BM = 128, TC_BN = 128, TC_D = 128;
constexpr int TC_STAGES = 4;
constexpr int TC_STAGE_BYTES = TC_BN * TC_D * 2;          // 32 KB: two [128 keys][64 features] swizzled boxes
constexpr int TC_HALF_BYTES = TC_STAGE_BYTES / 2;
constexpr int TC_THREADS = 320;                           // 8 softmax warps (2 groups) + TMA warp + MMA warp
constexpr int TC_WARP_TMA = 8, TC_WARP_MMA = 9;
constexpr uint32_t TM_COLS = 512;
constexpr int TC_SUB = 64;                                // keys per softmax step: each 128-key tile feeds 64 keys to each group
// TMEM columns: two 128-column S buffers (one 128-key tile each; group g reads columns [64g, 64g+64) and writes its P back
// over the first 32 of them) and one 128-column O accumulator per softmax group: 512 columns in all
__host__ __device__ constexpr uint32_t tm_s(int g, int b) { return (uint32_t)(b * 128 + g * 64); }
__host__ __device__ constexpr uint32_t tm_o(int g) { return (uint32_t)(256 + g * 128); }
constexpr float TC_RESCALE_LOG2 = 8.f;                    // rescale O only when the row max grows by > 2^8
constexpr int TC_OST_STRIDE = 132;                        // fp32 O staging row (128 + 4 floats): conflict-free row writes
constexpr size_t TC_QTILE_BYTES = (size_t)TC_BM * TC_D * 2;   // bf16 q block, same swizzled layout as a queue tile
constexpr size_t TC_SMEM_BYTES = 1024 + (size_t)TC_STAGES * TC_STAGE_BYTES + TC_QTILE_BYTES + 512;
static_assert((size_t)TC_BM * TC_OST_STRIDE * 4 <= (size_t)TC_STAGES * TC_STAGE_BYTES, "O staging must fit in the tile ring");

bool infonce_tcx_enabled();                                         // infonce_tcx.cu
int infonce_tcx_launch(const InfoNceStreamParams& P, cudaStream_t st);

struct TcDebug { unsigned long long* timebuf; };   // bring-up only: phase time stamps (tools/tc_timeline.py)

__device__ __forceinline__ void tc_stamp(const TcDebug& dbg, int slot)
{
    if (dbg.timebuf) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        dbg.timebuf[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 32 + slot] = t;
    }
}
__device__ __forceinline__ void tc_cstamp(const TcDebug& dbg, int slot)          // SM cycle counter, for intra-CTA phases
{
    if (dbg.timebuf) dbg.timebuf[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 32 + slot] = (unsigned long long)clock64();
}

struct TcBarriers {
    uint64_t full[TC_STAGES];     // TMA landed a queue tile
    uint64_t empty[TC_STAGES];    // both MMAs that read the tile have completed
    uint64_t s_full[2];           // [buffer] S = q Q^T of a 128-key tile is in TMEM (both groups wait on it)
    uint64_t p_full[4];           // [group * 2 + buffer] softmax finished with that S buffer (and wrote P over it)
    uint64_t o_done[2];           // per softmax group: one phase per completed O += P Q
    uint64_t acc_final;           // last O += P Q completed
    uint64_t q_ready;             // the bf16 q block is in shared memory (A operand of S = q Q^T)
    uint32_t tmem_base;
};

template <bool kWantAcc, bool kFixedMax>
__global__ void __launch_bounds__(TC_THREADS, 1)
infonce_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap qmap,
                  const InfoNceStreamParams P, const TcDebug dbg)
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
    // tiles are dealt round-robin: split s takes tiles s, s + nsplit, s + 2 nsplit, ...  At any moment the CTAs of a launch
    // read one contiguous window of the queue (148 x 32 KB) that moves through it, instead of 148 far-apart streams
    // that can pile onto the same HBM channels; the online softmax does not care about the order.
    const int n = (split < ntiles) ? (int)((ntiles - split + P.nsplit - 1) / P.nsplit) : 0;
#define TC_TILE_OF(i) ((long long)split + (long long)(i) * P.nsplit)

    if (threadIdx.x == 0) tc_stamp(dbg, 0);
    pdl_launch_dependents();                 // the finalize kernel may begin its (independent) prologue

    if (warp == TC_WARP_TMA && lane == 0) {
        prefetch_tmap(&tmap);
        prefetch_tmap(&qmap);
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&bar->full[s], 1); mbar_init(&bar->empty[s], 1); }
        for (int b = 0; b < 4; ++b) mbar_init(&bar->p_full[b], 128);
        for (int b = 0; b < 2; ++b) { mbar_init(&bar->s_full[b], 1); mbar_init(&bar->o_done[b], 1); }
        mbar_init(&bar->acc_final, 1);
        mbar_init(&bar->q_ready, 1);
        fence_barrier_init();
    }
    if (warp == TC_WARP_MMA) tmem_alloc<TM_COLS>(&bar->tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bar->tmem_base;
    if (threadIdx.x == 0) tc_stamp(dbg, 1);

    if (warp < 8) {
        // =============================================================== softmax: two warpgroups, one thread per query row.
        // Group g (warps 4g..4g+3) owns the tiles of parity g, the S buffer g and its OWN accumulator O_g: two independent
        // online-softmax streams that interleave on the MUFU and tensor pipes (while one group is in its MUFU-bound exp
        // sweep the other loads / scans its next tile) and are merged once at the end.
        const int g = warp >> 2, wq = warp & 3;                       // group, TMEM lane quarter
        const int r_loc = wq * 32 + lane;                             // row within the 128-row block
        const int row = row0 + r_loc;
        const bool valid = row < P.B;
        const uint32_t lane_addr = tmem + ((uint32_t)(wq * 32) << 16);
        const uint32_t o_addr = lane_addr + tm_o(g);
        // logits -> log2 domain.  The prep kernel folded P.q_scale (= log2(e)/T for this family) into the bf16 queries, so the
        // S tile is already in log2 units and the remaining factor is exactly 1
        const float c2 = 1.f;
        const float s_to_nat = 0.6931471805599453f;                    // S tile -> natural-log logits (materialised logits)

        // The bf16 q block and the positive logits were prepared once per step by infonce_prep_kernel (below): 148 CTAs
        // re-reading and re-converting the same fp32 rows would cost more L2 traffic than the queue itself.
        pdl_wait();                                                   // prep kernel results (pos_ws) are visible from here
        const float pos_nat0 = P.pos_ws[row];                         // natural-log units (q.k / T); 0 for padding rows
        const float pos_dot = pos_nat0 * 1.4426950408889634f;         // the positive in the units of the S tile
        if (split == 0 && blockIdx.y == 0 && threadIdx.x == 0) { for (int w = 0; w < 6; ++w) P.counter[w] = 0u; }      // re-arm the finalize control block
        if (split == 0 && g == 0 && valid && P.logits_out) P.logits_out[(size_t)row * P.ld_logits] = pos_nat0;
        if (threadIdx.x == 0) tc_stamp(dbg, 3);

        float m_run = kFixedMax ? (valid ? P.lse_fixed[row] * 1.4426950408889634f : 0.f) : -INFINITY;   // log2 domain
        float s_run = 0.f;
        int cnt = 0;

        // Group g takes keys [64g, 64g+64) of every 128-key tile; its S buffers alternate, so the S GEMM of step v+1 is
        // already in TMEM while step v is being exponentiated.
        // a warp whose 32 rows are all padding (B not a multiple of 128) only keeps the pipeline's barriers moving: its S rows
        // are q = 0 and its P / O rows are never read (GEMM rows are independent), so it skips the load / exp / store work
        const bool warp_has_rows = (row0 + wq * 32) < P.B;
        for (int v = 0; v < n; ++v) {
            const int sb = v & 1;
            const uint32_t s_addr = lane_addr + tm_s(g, sb);
            if (!warp_has_rows) {
                if (v == 0) { if (g == 0) named_barrier_arrive(2, 256); else named_barrier_sync(2, 256); }
                mbar_wait(&bar->s_full[sb], (v >> 1) & 1);
                tc_fence_before();
                mbar_arrive(&bar->p_full[g * 2 + sb]);
                continue;
            }
            if (threadIdx.x == 0 && v == 2) tc_cstamp(dbg, 16);
            if (v == 0 && g == 1) named_barrier_sync(2, 256);
            mbar_wait(&bar->s_full[sb], (v >> 1) & 1);
            tc_fence_after();
            if (threadIdx.x == 0 && v == 0) tc_stamp(dbg, 4);
            if (threadIdx.x == 0 && v == 2) tc_cstamp(dbg, 17);
            const long long key0 = TC_TILE_OF(v) * TC_BN + g * TC_SUB;
            const int nvalid = (P.K - key0 < TC_SUB) ? (int)((P.K - key0 > 0) ? (P.K - key0) : 0) : TC_SUB;
            uint32_t sr[64];
            tmem_ld32(s_addr, sr);
            tmem_ld32(s_addr + 32, sr + 32);
            tc_wait_ld();
            if (threadIdx.x == 0 && v == 2) tc_cstamp(dbg, 18);
            float* sv = reinterpret_cast<float*>(sr);
            if (P.logits_out && valid) {                                   // parity / debug path only
                float* dst = P.logits_out + (size_t)row * P.ld_logits + 1 + key0;
#pragma unroll
                for (int j = 0; j < TC_SUB; ++j) if (j < nvalid) dst[j] = sv[j] * s_to_nat;
            }
            if (nvalid < TC_SUB) {
#pragma unroll
                for (int j = 0; j < TC_SUB; ++j) if (j >= nvalid) sv[j] = -INFINITY;   // TMA zero-filled rows past K
            }
            bool count_step = true;
            if (!kFixedMax) {
                // row max with four independent chains
                float tm0 = -INFINITY, tm1 = -INFINITY, tm2 = -INFINITY, tm3 = -INFINITY;
#pragma unroll
                for (int j = 0; j < TC_SUB; j += 8) {
                    tm0 = max3(tm0, sv[j], sv[j + 1]);
                    tm1 = max3(tm1, sv[j + 2], sv[j + 3]);
                    tm2 = max3(tm2, sv[j + 4], sv[j + 5]);
                    tm3 = max3(tm3, sv[j + 6], sv[j + 7]);
                }
                const float raw_max = fmaxf(fmaxf(tm0, tm1), fmaxf(tm2, tm3));
                // no logit of this step beats the positive in any row of the warp: the rank count below is skipped (the
                // usual case once the encoder has learnt something; with random rows it almost never triggers)
                count_step = __any_sync(0xffffffffu, raw_max > pos_dot);
                const float xm = raw_max * c2;                              // -inf when the step is fully masked
                const bool need = xm > m_run + TC_RESCALE_LOG2;             // (m_run = -inf before the first valid step)
                if (__any_sync(0xffffffffu, need)) {                        // warp-uniform: TMEM ld/st are collective
                    const float m_new = need ? xm : m_run;
                    const float sc = (m_run == -INFINITY) ? 0.f : ex2(m_run - m_new);
                    s_run *= sc;
                    m_run = m_new;
                    if (kWantAcc && v > 0) {                                // (first step: O is still unwritten)
                        mbar_wait(&bar->o_done[g], (v - 1) & 1);            // this group's previous O += P Q has landed
                        tc_fence_after();
#pragma unroll
                        for (int ch = 0; ch < 4; ++ch) {
                            uint32_t t[32];
                            tmem_ld32(o_addr + 32 * ch, t);
                            tc_wait_ld();
#pragma unroll
                            for (int j = 0; j < 32; ++j) t[j] = __float_as_uint(__uint_as_float(t[j]) * sc);
                            tmem_st32(o_addr + 32 * ch, t);
                        }
                        tc_wait_st();
                    }
                }
            }
            // Stagger the two groups by about half a step: group 1 starts only when group 0 is through its first scan, so one
            // group's MUFU-bound exp sweep runs against the other's load / max phases instead of both colliding on the MUFU
            if (v == 0) { if (g == 0) named_barrier_arrive(2, 256); }
            if (threadIdx.x == 0 && v == 2) tc_cstamp(dbg, 19);
            // p = 2^(S c2 - m), row sum and rank count in one sweep (MUFU-bound: the compare/count instructions ride in
            // issue slots that would otherwise idle); four independent chains each, counts kept exact in fp32
            const float neg_m = (m_run == -INFINITY) ? 0.f : -m_run;        // fully masked so far: every p is 2^-inf = 0
            float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;
            float cf0 = 0.f, cf1 = 0.f, cf2 = 0.f, cf3 = 0.f;
            uint32_t pk[32];
            if (count_step) {
#pragma unroll
                for (int j = 0; j < TC_SUB; j += 4) {
                    cf0 += (sv[j] > pos_dot) ? 1.f : 0.f;
                    cf1 += (sv[j + 1] > pos_dot) ? 1.f : 0.f;
                    cf2 += (sv[j + 2] > pos_dot) ? 1.f : 0.f;
                    cf3 += (sv[j + 3] > pos_dot) ? 1.f : 0.f;
                    const float p0 = ex2(fmaf(sv[j], c2, neg_m)), p1 = ex2(fmaf(sv[j + 1], c2, neg_m));
                    const float p2 = ex2(fmaf(sv[j + 2], c2, neg_m)), p3 = ex2(fmaf(sv[j + 3], c2, neg_m));
                    rs0 += p0; rs1 += p1; rs2 += p2; rs3 += p3;
                    pk[j >> 1] = pack_bf16(p0, p1);
                    pk[(j >> 1) + 1] = pack_bf16(p2, p3);
                }
            } else {
#pragma unroll
                for (int j = 0; j < TC_SUB; j += 4) {
                    const float p0 = ex2(fmaf(sv[j], c2, neg_m)), p1 = ex2(fmaf(sv[j + 1], c2, neg_m));
                    const float p2 = ex2(fmaf(sv[j + 2], c2, neg_m)), p3 = ex2(fmaf(sv[j + 3], c2, neg_m));
                    rs0 += p0; rs1 += p1; rs2 += p2; rs3 += p3;
                    pk[j >> 1] = pack_bf16(p0, p1);
                    pk[(j >> 1) + 1] = pack_bf16(p2, p3);
                }
            }
            cnt += (int)((cf0 + cf1) + (cf2 + cf3));
            s_run += (rs0 + rs1) + (rs2 + rs3);
            if (threadIdx.x == 0 && v == 2) tc_cstamp(dbg, 20);
            if (kWantAcc) {
                tmem_st32(s_addr, pk);                                      // P (bf16, 64 keys) over the first 32 S columns
                tc_wait_st();
            }
            tc_fence_before();
            mbar_arrive(&bar->p_full[g * 2 + sb]);
            if (threadIdx.x == 0 && v == 2) tc_cstamp(dbg, 21);
            if (threadIdx.x == 0 && v == 3) tc_cstamp(dbg, 22);
        }
        if (threadIdx.x == 0) tc_stamp(dbg, 5);

        // ---- merge the two groups' streams (max, sum, count, O) and write ONE split partial per row
        float* st_m = reinterpret_cast<float*>(qtile);                     // the q tile is dead once the last S GEMM is done;
        float* st_s = st_m + 256;                                          // these are read only after acc_final below
        int*   st_c = reinterpret_cast<int*>(st_s + 256);
        if (kWantAcc) { mbar_wait(&bar->acc_final, 0); tc_fence_after(); } // every tcgen05.mma of this CTA has completed
        if (threadIdx.x == 0) tc_stamp(dbg, 6);
        st_m[g * 128 + r_loc] = m_run; st_s[g * 128 + r_loc] = s_run; st_c[g * 128 + r_loc] = cnt;
        named_barrier_sync(1, 256);
        const float m_o = st_m[(g ^ 1) * 128 + r_loc], s_o = st_s[(g ^ 1) * 128 + r_loc];
        const int c_o = st_c[(g ^ 1) * 128 + r_loc];
        float w_me = 1.f, w_ot = 0.f, m_all = m_run;
        if (!kFixedMax) {
            m_all = fmaxf(m_run, m_o);                                     // a group whose keys were all past K has m = -inf
            w_me = (m_run == -INFINITY) ? 0.f : ex2(m_run - m_all);
            w_ot = (m_o == -INFINITY) ? 0.f : ex2(m_o - m_all);
        } else {
            w_ot = 1.f;
        }
        if (g == 0) {
            const size_t po = part_stat_index(split, row, P.nsplit);
            P.part_max[po] = kFixedMax ? 0.f : m_all * 0.6931471805599453f;  // back to natural-log units
            // logits beyond the unit-row range (un-normalised inputs): the finalize kernel must not use its packed
            // fixed-point loss word (control word 6, cleared by the finalize kernel)
            if (!kFixedMax && valid && m_all * 0.6931471805599453f > 1.0625f * P.inv_T) P.counter[6] = 1u;
            P.part_sum[po] = s_run * w_me + s_o * w_ot;
            P.part_cnt[po] = cnt + c_o;
        }
        if (kWantAcc) {
            // group g merges feature columns [64g, 64g+64) of its row: O = w_me O_mine + w_ot O_other, staged in the (idle)
            // tile ring with a padded stride, then written out as coalesced 512-byte rows
            float* ost = reinterpret_cast<float*>(stages);
            const uint32_t o_other = lane_addr + tm_o(g ^ 1);
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                const int col = g * 64 + ch * 32;
                uint32_t a[32], b[32];
                tmem_ld32(o_addr + col, a);
                tmem_ld32(o_other + col, b);
                tc_wait_ld();
                float4* dst = reinterpret_cast<float4*>(ost + (size_t)r_loc * TC_OST_STRIDE + col);
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    float4 o;
                    o.x = __uint_as_float(a[4 * v]) * w_me;     o.y = __uint_as_float(a[4 * v + 1]) * w_me;
                    o.z = __uint_as_float(a[4 * v + 2]) * w_me; o.w = __uint_as_float(a[4 * v + 3]) * w_me;
                    o.x = fmaf(__uint_as_float(b[4 * v]), w_ot, o.x);     o.y = fmaf(__uint_as_float(b[4 * v + 1]), w_ot, o.y);
                    o.z = fmaf(__uint_as_float(b[4 * v + 2]), w_ot, o.z); o.w = fmaf(__uint_as_float(b[4 * v + 3]), w_ot, o.w);
                    dst[v] = o;
                }
            }
            named_barrier_sync(1, 256);
            // 8 warps x 16 rows each, one 512-byte row per warp instruction
            float4* gdst = reinterpret_cast<float4*>(P.part_acc + ((size_t)split * P.Bpad + row0) * TC_D);
#pragma unroll 4
            for (int rr = 0; rr < 16; ++rr) {
                const int r = warp * 16 + rr;
                if (r < nrows) gdst[r * 32 + lane] = *reinterpret_cast<const float4*>(ost + (size_t)r * TC_OST_STRIDE + lane * 4);
            }
        }
    } else if (warp == TC_WARP_TMA) {
        // =============================================================== TMA producer
        if (lane == 0) {
            // queue tiles do not depend on the prep kernel: get the first one moving before waiting for it
            const int pre = n < TC_STAGES ? n : TC_STAGES;
            auto load_tile = [&](int i) {
                const int stage = i % TC_STAGES;
                uint8_t* dst = stages + (size_t)stage * TC_STAGE_BYTES;
                const int key0 = (int)(TC_TILE_OF(i) * TC_BN);
                mbar_arrive_expect_tx(&bar->full[stage], TC_STAGE_BYTES);
                tma_load_2d(dst, &tmap, &bar->full[stage], 0, key0);                     // features  0..63
                tma_load_2d(dst + TC_HALF_BYTES, &tmap, &bar->full[stage], 64, key0);    // features 64..127
            };
            tc_stamp(dbg, 9);
            load_tile(0);
            pdl_wait();                                                              // q_bf16 comes from the prep kernel
            mbar_arrive_expect_tx(&bar->q_ready, (uint32_t)TC_QTILE_BYTES);
            tma_load_2d(qtile, &qmap, &bar->q_ready, 0, row0);                       // features  0..63 of the 128 query rows
            tma_load_2d(qtile + TC_HALF_BYTES, &qmap, &bar->q_ready, 64, row0);      // features 64..127
            for (int i = 1; i < pre; ++i) load_tile(i);
            for (int i = pre; i < n; ++i) {
                mbar_wait(&bar->empty[i % TC_STAGES], ((i / TC_STAGES) - 1) & 1);
                load_tile(i);
            }
        }
    } else {
        // =============================================================== MMA issuer: the whole warp walks the loop (uniform
        // control flow keeps descriptors in uniform registers); one elected lane issues tcgen05.mma / commit.
        // Program order per iteration: S GEMM of tile i, then O GEMM of tile i-1 -- so the S GEMM of the next tile (other
        // group's buffer) is already queued while a group is still in its softmax sweep.
        const bool leader = elect_one();
        constexpr uint32_t idesc_s = make_idesc_bf16(TC_BM, TC_BN, 0, 0);   // S: A = q tile, B = 128 queue rows, both K-major
        constexpr uint32_t idesc_o = make_idesc_bf16(TC_BM, TC_D, 0, 1);    // O: A = P (TMEM), B = 64 of those rows, MN-major
        mbar_wait(&bar->q_ready, 0);
        tc_fence_after();
        if (leader) tc_stamp(dbg, 10);
        const uint32_t qbase = smem_u32(qtile);

        // S GEMM of tile v: S[v & 1] = q . tile^T, one 128 x 128 x 128 product for both softmax groups (a 64-key product
        // per group would re-read the q tile from shared memory twice as often, and SS-mode MMAs are smem-bandwidth-bound)
        auto issue_s = [&](int v) {
            const uint32_t sbase = smem_u32(stages + (size_t)(v % TC_STAGES) * TC_STAGE_BYTES);
            const uint32_t d_tmem = tmem + tm_s(0, v & 1);
            if (leader) {
#pragma unroll
                for (int kk = 0; kk < TC_D / 16; ++kk) {
                    // 16 features per MMA: 32 bytes along the swizzled 128-byte row, next 64-feature box after 4 steps
                    // (K-major, 128-byte swizzle: LBO unused, SBO = 1024 B between 8-row groups)
                    const uint32_t koff = (kk >> 2) * TC_HALF_BYTES + (kk & 3) * 32;
                    mma_ss(d_tmem, make_smem_desc_sw128(qbase + koff, 16, 1024), make_smem_desc_sw128(sbase + koff, 16, 1024),
                           idesc_s, kk > 0);
                }
                tc_commit(&bar->s_full[v & 1]);
                if (!kWantAcc) tc_commit(&bar->empty[v % TC_STAGES]);       // forward only: the tile is not needed again
            }
            __syncwarp();
        };

        // prologue: the first two tiles
        for (int v = 0; v < n && v < 2; ++v) {
            mbar_wait(&bar->full[v % TC_STAGES], 0);
            tc_fence_after();
            if (v == 0 && leader) tc_stamp(dbg, 11);
            issue_s(v);
        }
        for (int v = 0; v < n; ++v) {
            const int stage = v % TC_STAGES;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                if (v == 3 && g == 0 && leader) tc_cstamp(dbg, 25);
                mbar_wait(&bar->p_full[g * 2 + (v & 1)], (v >> 1) & 1);     // softmax of (g, v) done: P written / S consumed
                tc_fence_after();
                if (v == 3 && g == 0 && leader) tc_cstamp(dbg, 26);
                if (kWantAcc) {
                    const uint32_t sbase = smem_u32(stages + (size_t)stage * TC_STAGE_BYTES) + g * (TC_SUB * 128);
                    const uint32_t p_tmem = tmem + tm_s(g, v & 1);
                    if (leader) {
#pragma unroll
                        for (int kk = 0; kk < TC_SUB / 16; ++kk) {
                            // 16 keys per MMA = two 8-row swizzle atoms (2 KB); LBO = next 64-feature box, SBO = next 8 keys
                            mma_ts(tmem + tm_o(g), p_tmem + kk * 8, make_smem_desc_sw128(sbase + kk * 2048, TC_HALF_BYTES, 1024),
                                   idesc_o, (v > 0 || kk > 0) ? 1u : 0u);
                        }
                        tc_commit(&bar->o_done[g]);
                        if (g == 1) tc_commit(&bar->empty[stage]);          // both groups' GEMMs on this tile are queued
                        if (g == 1 && v == n - 1) tc_commit(&bar->acc_final);
                    }
                    __syncwarp();
                }
                if (v == 3 && g == 0 && leader) tc_cstamp(dbg, 27);
            }
            // refill S buffer (v & 1) with tile v+2: queued behind the two O GEMMs that read P from it
            if (v + 2 < n) {
                mbar_wait(&bar->full[(v + 2) % TC_STAGES], ((v + 2) / TC_STAGES) & 1);
                tc_fence_after();
                issue_s(v + 2);
            }
            if (v == 3 && leader) tc_cstamp(dbg, 28);
        }
    }
    if (threadIdx.x == 0) tc_stamp(dbg, 7);
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == TC_WARP_MMA) tmem_dealloc<TM_COLS>(tmem);
    if (threadIdx.x == 0) tc_stamp(dbg, 8);
}

// Once per step: q -> bf16 [Bpad, 128] (zero rows past B) and the positive logits q.k / T in fp32.  One warp per row,
// 128-bit coalesced loads; the dot product has a fixed shuffle order (deterministic).
__global__ void __launch_bounds__(256)
infonce_prep_kernel(const float* __restrict__ q, const float* __restrict__ k, int B, int Bpad, float inv_T,
                    __nv_bfloat16* __restrict__ q_bf16, float* __restrict__ pos_ws, float* __restrict__ pos_out,
                    unsigned long long* timebuf, const PeerXchg X, int nprep, unsigned int* range_flag,
                    float* __restrict__ k_hat, float* __restrict__ inv_nq, int normalize,
                    const char* __restrict__ pf_base, unsigned long long pf_bytes, float q_scale, const int npush,
                    const int gather_Bl, const int late_trigger)
{
    // launched with a programmatic dependency on whatever kernel precedes it in the stream (normally the finalize launch of the
    // previous step, whose enqueue CTAs write the queue): resident and past its launch latency when that kernel drains.
    // Nothing of this step may start before it has completed -- the sweep's first queue tiles are loaded before ITS wait.
    ptx::pdl_wait();
    // late_trigger (replica step with the key push riding in this launch): the row CTAs release the streaming kernel only
    // AFTER their rows are written and fenced -- its launch is then the "rows are ready" signal and it never waits for the
    // completion of this launch, i.e. for the NVLink round trip of the push CTAs' system-scope releases.  Otherwise the
    // streaming kernel may start its setup and its first queue-tile loads right away (it waits for this launch later).
    const bool row_cta = (int)blockIdx.x >= npush;
    if (!(late_trigger && row_cta)) ptx::pdl_launch_dependents();
    // warm the L2 with the head of the queue (the first tile waves of the stream kernel): this launch starts ~1 us before
    // the stream kernel's TMA producer can, and a cold queue tile costs a full HBM round trip at the head of every CTA's
    // pipeline.  32 KB bulk prefetches, dealt round-robin over the CTAs of this launch (a few per SM).
    // the push CTAs of a peer exchange are the FIRST blocks of the launch: a row CTA that waits for a peer's rows (gather
    // mode) can never keep them from being scheduled
    const int rb = (int)blockIdx.x - npush;                                 // row block of this CTA (< 0: push CTA)
    if (pf_base != nullptr && rb >= 0) {
        const unsigned long long off = ((unsigned long long)rb + (unsigned long long)nprep * threadIdx.x) * 32768ull;
        if (off < pf_bytes) {
            const unsigned long long rest = pf_bytes - off;
            const unsigned int nb = rest < 32768ull ? (unsigned int)rest : 32768u;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(pf_base + off), "r"(nb) : "memory");
        }
    }
    if (rb < 0) {
        // exchange riding in this launch: push slice c of the local rows to rank p (exchange.cu).  Replica step: the keys k
        // [B, d]; gather mode (K-sharded step): q holds this rank's [q_loc; k_loc] block of 2 * gather_Bl rows
        const int e = blockIdx.x;
        const unsigned long long step = *reinterpret_cast<volatile unsigned long long*>(X.xstate);
        xchg_push_slice(X, reinterpret_cast<const float4*>(gather_Bl > 0 ? q : k), step, e / XCHG_SLICES, e % XCHG_SLICES);
        return;
    }
    const int lane = threadIdx.x & 31, row = rb * (int)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (timebuf && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMin(timebuf + 32 * 1000 + 2, t);
    }
    if (row >= Bpad) return;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (row < B && gather_Bl > 0) {
        // gather mode: global row `row` is row lr of rank p; its q and k rows arrive in slot p of this rank's mailbox
        // ([q rows; k rows] of that rank).  Wait for the two slices that hold them, then read around L1.
        const unsigned long long step = *reinterpret_cast<volatile unsigned long long*>(X.xstate);
        const int p = row / gather_Bl, lr = row - p * gather_Bl;
        const int per = (X.n4 + XCHG_SLICES - 1) / XCHG_SLICES;
        const int iq = lr * (TC_D / 4), ik = (gather_Bl + lr) * (TC_D / 4);
        // (four lanes poll the slices of the two rows side by side: one flag round trip instead of four)
        if (lane < 4) {
            const int i0 = (lane & 2) ? ik : iq;
            xchg_wait_slice(X, step, p, ((lane & 1) ? i0 + TC_D / 4 - 1 : i0) / per);
        }
        __syncwarp();
        const float4* slot = xchg_slot(X.mailboxes[X.rank], X, (int)(step & 1ull), p);
        a = ld_cg_f4(slot + iq + lane);
        b = ld_cg_f4(slot + ik + lane);
    } else if (row < B) {
        a = __ldg(reinterpret_cast<const float4*>(q + (size_t)row * TC_D) + lane);
        b = __ldg(reinterpret_cast<const float4*>(k + (size_t)row * TC_D) + lane);
    }
    if (normalize) {
        // projection-head tail (Normalize(2), project_head.py:4-10) fused in: both rows become unit rows here
        float na = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, a.w * a.w)));
        float nb = fmaf(b.x, b.x, fmaf(b.y, b.y, fmaf(b.z, b.z, b.w * b.w)));
        na = warp_sum(na); nb = warp_sum(nb);
        const float ia = 1.f / fmaxf(sqrtf(na), 1e-12f), ib = 1.f / fmaxf(sqrtf(nb), 1e-12f);
        a.x *= ia; a.y *= ia; a.z *= ia; a.w *= ia;
        b.x *= ib; b.y *= ib; b.z *= ib; b.w *= ib;
        if (lane == 0) inv_nq[row] = ia;
    }
    // the positive keys are staged in device memory for the finalize kernel (normalised or as given): k itself is read
    // exactly once per step, so it may live in pinned host memory (zero-copy callers)
    if (row < B) reinterpret_cast<float4*>(k_hat + (size_t)row * TC_D)[lane] = b;         // (may be a caller buffer of B rows)
    float dsum = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
    dsum = warp_sum(dsum) * inv_T;
    uint2 pk;                                                        // (q_scale: see infonce_tcx.cu -- log2(e)/T or 1)
    pk.x = ptx::pack_bf16(a.x * q_scale, a.y * q_scale);
    pk.y = ptx::pack_bf16(a.z * q_scale, a.w * q_scale);
    reinterpret_cast<uint2*>(q_bf16 + (size_t)row * TC_D)[lane] = pk;
    if (lane == 0) {
        pos_ws[row] = dsum;
        if (pos_out && row < B) pos_out[row] = dsum;
        if (fabsf(dsum) > 1.0625f * inv_T) range_flag[0] = 1u;        // see the partial write of the stream kernel
    }
    if (timebuf && threadIdx.x == 0) atomicMax(timebuf + 32 * 1000 + 4, globaltimer_ns());       // bring-up only
    if (late_trigger) {
        __threadfence();                     // this thread's rows are visible device-wide before its CTA counts as triggered
        ptx::pdl_launch_dependents();
    }
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

// row-major bf16 [rows, 128] tensor, box = 128 rows x 64 features, 128-byte swizzle (queue tiles and the q block alike)
int get_queue_tmap(const void* queue, long long K, CUtensorMap* out)
{
    static thread_local TmapCache cache[8] = {};
    static thread_local int next = 0;
    for (int i = 0; i < 8; ++i)
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
    next = (next + 1) & 7;
    *out = m;
    return GCA_OK;
}

// row-major bf16 [rows, 128] tensor, box = box_rows rows x 64 features, 128-byte swizzle (not cached: callers keep the map)
int make_bf16_tmap(const void* base, long long rows, int box_rows, CUtensorMap* out)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_err(GCA_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {(cuuint64_t)TC_D, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)TC_D * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_err(GCA_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
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

bool pdl_enabled()
{
    static int on = -1;                               // GCA_NO_PDL=1 turns programmatic dependent launch off (A/B timing)
    if (on < 0) { const char* e = getenv("GCA_NO_PDL"); on = (e && e[0] == '1') ? 0 : 1; }
    return on != 0;
}

unsigned long long* debug_timebuf()
{
    const char* tb = getenv("GCA_TC_TIMEBUF");        // device address of a 32 * 1024 uint64 buffer (tools/tc_timeline.py)
    return tb ? (unsigned long long*)strtoull(tb, nullptr, 0) : nullptr;
}

static TcDebug tc_debug_knobs() { return TcDebug{debug_timebuf()}; }

int infonce_tc_launch(const InfoNceStreamParams& P_, bool fixed_max, cudaStream_t st)
{
    InfoNceStreamParams P = P_;
    P.q_scale = P.inv_T * 1.4426950408889634f;        // every kernel of this family works on log2-domain S tiles
    if (P.d != TC_D) return set_err(GCA_ERR_UNSUPPORTED, "tcgen05 InfoNCE kernel needs d == %d (got %d)", TC_D, P.d);
    if (P.K >= (1ll << 31) - TC_BN) return set_err(GCA_ERR_UNSUPPORTED, "tcgen05 InfoNCE kernel: K too large");
    if ((reinterpret_cast<uintptr_t>(P.queue) & 15) != 0) return set_err(GCA_ERR_BAD_ARG, "queue must be 16-byte aligned");
    CUtensorMap tmap, qmap;
    int rc = get_queue_tmap(P.queue, P.K, &tmap);
    if (rc != GCA_OK) return rc;
    rc = get_queue_tmap(P.q_bf16_ws, P.Bpad, &qmap);
    if (rc != GCA_OK) return rc;
    if (P.skip_prep && P.xchg.mailboxes) return set_err(GCA_ERR_BAD_ARG, "the peer key exchange rides in the prep kernel");
    if (P.normalize && (P.xchg.mailboxes || P.skip_prep))
        return set_err(GCA_ERR_UNSUPPORTED, "projection-tail fusion cannot be combined with the peer exchange or skip_prep");
    // loss + gradient in one sweep (the product path): second-generation stream kernel (infonce_tcx.cu)
    const bool use_tcx = P.part_acc != nullptr && !fixed_max && P.logits_out == nullptr && infonce_tcx_enabled();
    if (!P.skip_prep) {
        constexpr int prep_rows = 8;                         // rows (warps) per CTA of the prep launch (1 / 2 / 4 measured: no difference)
        const int nprep = (P.Bpad + prep_rows - 1) / prep_rows;
        // peer exchange: the q|k gather of the K-sharded step rides in this launch (its rows are needed by this very launch);
        // the key push of the replica step goes out as its own small launch on a side stream (joined by the caller after the
        // step's last launch, gca_api.cu)
        const bool gather = P.xchg.mailboxes != nullptr && P.gather_Bl > 0;
        // replica step on the single-pass kernel with programmatic launches, issued directly or from a launch plan: the key push
        // rides in this launch and the row CTAs trigger late (see the kernel) -- one linear chain of launches, which also
        // chains consecutive steps (24.6 us per step at 2 GPUs against 29.9 us).  Otherwise the push goes out on a side
        // stream (the other stream kernels wait for the whole prep grid).
        static int push_side = -1;                          // GCA_PUSH_SIDE=1: side-stream push everywhere (A/B timing)
        if (push_side < 0) { const char* e = getenv("GCA_PUSH_SIDE"); push_side = (e && e[0] == '1') ? 1 : 0; }
        // (a captured step is one graph launch: there the side branch is faster -- 29.9 us against 31.6 us at 2 GPUs)
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (plan_recording() == nullptr && cudaStreamIsCapturing(st, &cap) != cudaSuccess) { (void)cudaGetLastError(); cap = cudaStreamCaptureStatusActive; }
        const bool late = P.xchg.mailboxes != nullptr && !gather && use_tcx && pdl_enabled() && !push_side &&
                          cap == cudaStreamCaptureStatusNone;
        const bool push_first = P.xchg.mailboxes != nullptr && !gather && !late;
        const int npush = (gather || late) ? P.xchg.W * XCHG_SLICES : 0;
        P.no_prep_wait = late ? 1 : 0;
        if (push_first) {
            rc = keys_push_fork(P.k, P.xchg, st);
            if (rc != GCA_OK) return rc;
        }
        // queue head to prefetch into L2 (at most 32 MB; GCA_NO_L2PF=1 turns it off for A/B timing)
        static int pf_on = -1;
        if (pf_on < 0) { const char* e = getenv("GCA_NO_L2PF"); pf_on = (e && e[0] == '1') ? 0 : 1; }
        unsigned long long pf_bytes = (unsigned long long)P.K * TC_D * 2;
        const unsigned long long pf_waves = 2ull * (unsigned long long)P.nsplit * TC_STAGE_BYTES;   // tiles 0 and 1 of every split
        if (pf_bytes > pf_waves) pf_bytes = pf_waves;
        if ((unsigned long long)nprep * (32ull * prep_rows) * 32768ull < pf_bytes) pf_bytes = (unsigned long long)nprep * (32ull * prep_rows) * 32768ull;
        cudaLaunchConfig_t pcfg{};
        pcfg.gridDim = dim3(nprep + npush); pcfg.blockDim = dim3(32 * prep_rows); pcfg.dynamicSmemBytes = 0; pcfg.stream = st;
        cudaLaunchAttribute pattr[1];
        pattr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        pattr[0].val.programmaticStreamSerializationAllowed = 1;
        static int prep_pdl = -1;                              // GCA_PREP_PDL=0: plain stream order in front of the prep launch (A/B timing)
        if (prep_pdl < 0) { const char* e = getenv("GCA_PREP_PDL"); prep_pdl = (e && e[0] == '0') ? 0 : 1; }
        pcfg.attrs = pattr; pcfg.numAttrs = (pdl_enabled() && prep_pdl) ? 1 : 0;
        GCA_CUDA(launch_ex(&pcfg, infonce_prep_kernel, P.q, P.k, P.B, P.Bpad, P.inv_T, (__nv_bfloat16*)P.q_bf16_ws,
                                    P.pos_ws, P.pos_out, debug_timebuf(), P.xchg, nprep, P.counter + 6,
                                    P.k_hat, P.inv_nq, P.normalize,
                                    pf_on ? (const char*)P.queue : (const char*)nullptr, pf_bytes,
                                    P.q_scale, npush, P.gather_Bl, late ? 1 : 0));
        GCA_LAUNCH_CHECK("infonce_prep_kernel");
        count_launch(1);
    }
    const TcDebug dbg = tc_debug_knobs();
    const bool want_acc = P.part_acc != nullptr;
    if (use_tcx) return infonce_tcx_launch(P, st);
    dim3 grid(P.nsplit, P.Bpad / TC_BM);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = TC_SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    int na = 0;
    if (pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;    // PDL: overlap this kernel's setup with the prep kernel
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
#define GCA_TC_LAUNCH(ACC, FIX) do { \
        auto kern = infonce_tc_kernel<ACC, FIX>; \
        GCA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES)); \
        GCA_CUDA(launch_ex(&cfg, kern, tmap, qmap, P, dbg)); } while (0)
    if (want_acc) { if (fixed_max) GCA_TC_LAUNCH(true, true); else GCA_TC_LAUNCH(true, false); }
    else          { if (fixed_max) GCA_TC_LAUNCH(false, true); else GCA_TC_LAUNCH(false, false); }
#undef GCA_TC_LAUNCH
    GCA_LAUNCH_CHECK("infonce_tc_kernel");
    return GCA_OK;
}

}  // namespace gca
