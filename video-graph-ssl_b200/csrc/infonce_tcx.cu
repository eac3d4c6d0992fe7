// K1x: single-pass InfoNCE stream, second generation (the product path for loss + dq on a bf16 queue with d = 128).
//
// Same arithmetic as infonce_tc.cu -- S = q Q^T (tcgen05.mma SS), p = 2^(S - m), O += P Q (tcgen05.mma TS with the SAME
// smem tile through an MN-major descriptor), one CTA = 128 query rows x a round-robin set of 128-key queue tiles -- but
// organised around a FIXED per-row reference exponent m instead of an online row max:
//   * unit rows bound every logit by 1/T, so m = 0 (or 1/T - 100 log2 units for tiny temperatures) can neither overflow nor
//     underflow fp32 / bf16; without a running max there is no max pass, no O rescale, no per-group accumulator: all 16
//     softmax warps feed ONE O accumulator, which frees TMEM for a third S buffer (the softmax never waits for a GEMM);
//   * the prep kernel folds log2(e)/T into the bf16 queries, so the S tile is already the exponent: a key costs one
//     MUFU.EX2, half a packed fp32x2 add (row sum) and half a bf16x2 pack -- no scaling instruction at all;
//   * 16 softmax warps (4 per scheduler, 32 rows x 32 keys each per tile) instead of 8: the TMEM load / store / barrier
//     latencies of one warp hide under the exponentials of the other three;
//   * the rank of the positive (accuracy top-1/top-5) is counted only from the first step whose row sums cannot prove that
//     no key beats the positive, then inside the sweep (pos - S has its sign bit set exactly when S > pos), and -- for
//     callers that only want top-k hit counts -- only until every row of the warp is past GCA_TOPK_RANK_CAP.
// Generality is kept by verification instead of by an online max: after the sweep every row checks that its sum stayed
// inside a window that excludes overflow and harmful underflow (2^-60 .. 2^100).  If any row of the CTA fails (inputs
// far from unit norm), the CTA redoes its key range: one max-only pass (S GEMM + row max), then the same sweep with the
// exact row max as the reference.  Split partials keep the layout of gca_common.cuh, so finalize.cu is unchanged.
// Measured (profiles/r02_notes.md): 12.2 us per launch at (256, 65536) against 14.2 for infonce_tc.cu, 112 us against 152
// at K = 2^20 (1.23 PFLOP/s).  What bounds it now is shared-memory bandwidth: an SS-form 128x128x16 MMA reads 8 KB per 64
// cycles, which together with the TMA writes and the O GEMM's operand reads is the whole 128 B/clk of the SM.
// Replaces mem_moco.py:36-46 + criterion.py:44 + autograd(mm) + metric.py:44-67 of the reference.
#include "gca_common.cuh"
#include "infonce_params.cuh"
#include "tc_ptx.cuh"
#include "launch_plan.cuh"
#include <stdlib.h>


namespace gca {

constexpr int X_BM = 128, X_BN = 128, X_D = 128;
constexpr int X_STAGES = 5;
constexpr int X_STAGE_BYTES = X_BN * X_D * 2;               // 32 KB: two [128 keys][64 features] swizzled boxes
constexpr int X_HALF_BYTES = X_STAGE_BYTES / 2;
constexpr int X_SM_WARPS = 16;                              // softmax warps: lane quarter = warp & 3, key group = warp >> 2
constexpr int X_SM_THREADS = X_SM_WARPS * 32;
constexpr int X_WARP_TMA = 16, X_WARP_MMA = 17;
constexpr int X_THREADS = 18 * 32;
#ifndef GCA_X_TMA_LEAD
#define GCA_X_TMA_LEAD 2
#endif
constexpr int X_TMA_LEAD = GCA_X_TMA_LEAD;                               // queue tiles in flight per CTA
constexpr int X_NSBUF = 3;                                  // S buffers of 128 TMEM columns; O takes the last 128
constexpr uint32_t X_TM_COLS = 512;
__host__ __device__ constexpr uint32_t xtm_s(int b) { return (uint32_t)(b * 128); }
constexpr uint32_t X_TM_O = 384;
constexpr size_t X_QTILE_BYTES = (size_t)X_BM * X_D * 2;     // bf16 q block, same swizzled layout as a queue tile
constexpr int X_OST_STRIDE = 132;                           // fp32 O staging row (128 + 4 floats): conflict-free row writes
constexpr float X_WIN_LO = 8.6736174e-19f;                  // 2^-60
constexpr float X_WIN_HI = 1.2676506e30f;                   // 2^100

struct XBarriers {
    uint64_t full[X_STAGES];      // TMA landed a queue tile
    uint64_t empty[X_STAGES];     // every MMA that reads the tile has completed
    uint64_t s_full[X_NSBUF];     // S = q Q^T of a tile is in TMEM
    uint64_t p_full[X_NSBUF];     // the 16 softmax warps are done with that S buffer (P written over it)
    uint64_t acc_final;           // last O += P Q of a sweep completed
    uint64_t q_ready;             // the bf16 q block is in shared memory
    uint32_t tmem_base;
    uint32_t redo;                // some row of this CTA left the validity window: redo with the exact row max
    float    xs[4][X_BM];         // cross-key-group exchange: row sums / row maxima
    int      xc[4][X_BM];         // rank counts
};
constexpr size_t X_SMEM_BYTES = 1024 + (size_t)X_STAGES * X_STAGE_BYTES + X_QTILE_BYTES + sizeof(XBarriers) + 64;
static_assert((size_t)X_BM * X_OST_STRIDE * 4 <= (size_t)X_STAGES * X_STAGE_BYTES, "O staging must fit in the tile ring");
static_assert(X_SMEM_BYTES <= 232448, "shared memory budget");

struct XDebug { unsigned long long* timebuf; };
__device__ __forceinline__ void x_stamp(const XDebug& dbg, int slot)
{
    if (dbg.timebuf) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        dbg.timebuf[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 32 + slot] = t;
    }
}
__device__ __forceinline__ void x_cstamp(const XDebug& dbg, int slot)
{
    if (dbg.timebuf) dbg.timebuf[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 32 + slot] = (unsigned long long)clock64();
}

// ---------------------------------------------------------------- packed fp32x2 helpers (FFMA2 / FADD2 / FMUL2, sm_100+)
namespace x2 {
__device__ __forceinline__ uint64_t pack(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t add(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t sub(uint64_t a, uint64_t b) { uint64_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
}  // namespace x2

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
                 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 :: "r"(taddr),
                    "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(reinterpret_cast<uint64_t>(gdst)), "r"(ptx::smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait_read()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// One softmax step of a thread: 32 keys of its row.  The S tile already holds log2-domain logits (the prep kernel folded
// log2(e)/T into the bf16 queries), so p = 2^(S - m) needs no scaling instruction, and with the reference exponent m = 0 of
// the normal first sweep (kZeroRef) no subtraction either: the MUFU pairs go from TMEM register to MUFU.EX2 directly.
// Row sums are packed fp32x2 adds; with kCount the rank count rides in the same instruction stream (pos - S has its sign
// bit set exactly when S > pos).
template <bool kCount, bool kZeroRef>
__device__ __forceinline__ float sweep32(const float* sv, uint64_t nm2, float pos_l2, uint32_t* pk, int& cnt)
{
    uint64_t acc0 = x2::pack(0.f, 0.f), acc1 = acc0;
    const uint64_t pp = x2::pack(pos_l2, pos_l2), neg1 = x2::pack(-1.f, -1.f);
    int c0 = 0, c1 = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float xa = sv[2 * i], xb = sv[2 * i + 1], pa, pb;
        if (!kZeroRef) x2::unpack(x2::add(x2::pack(xa, xb), nm2), xa, xb);
        pa = ptx::ex2(xa); pb = ptx::ex2(xb);
        if (i & 1) acc1 = x2::add(acc1, x2::pack(pa, pb)); else acc0 = x2::add(acc0, x2::pack(pa, pb));
        pk[i] = ptx::pack_bf16(pa, pb);
        if (kCount) {
            float da, db;
            x2::unpack(x2::fma(x2::pack(sv[2 * i], sv[2 * i + 1]), neg1, pp), da, db);
            c0 += (int)(__float_as_uint(da) >> 31);
            c1 += (int)(__float_as_uint(db) >> 31);
        }
    }
    if (kCount) cnt += c0 + c1;
    float a0, a1, b0, b1;
    x2::unpack(acc0, a0, a1);
    x2::unpack(acc1, b0, b1);
    return (a0 + a1) + (b0 + b1);
}

__device__ __forceinline__ int count32(const float* sv, float pos_l2)
{
    const uint64_t pp = x2::pack(pos_l2, pos_l2), neg1 = x2::pack(-1.f, -1.f);
    int c0 = 0, c1 = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float da, db;
        x2::unpack(x2::fma(x2::pack(sv[2 * i], sv[2 * i + 1]), neg1, pp), da, db);
        c0 += (int)(__float_as_uint(da) >> 31);
        c1 += (int)(__float_as_uint(db) >> 31);
    }
    return c0 + c1;
}

// kMode of one pass over the CTA's tiles
enum { X_PASS_SWEEP = 0, X_PASS_MAX = 1 };

// kDbg: phase time stamps for tools/tc_timeline.py (a separate instantiation: the product kernel carries none of it)
template <bool kDbg>
__global__ void __launch_bounds__(X_THREADS, 1)
infonce_tcx_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap qmap,
                   const InfoNceStreamParams P, const XDebug dbg, const float m_ref0, const int rank_cap)
{
    using namespace ptx;
#define X_STAMP(slot)  do { if (kDbg && threadIdx.x == 0) x_stamp(dbg, slot); } while (0)
#define X_CSTAMP(cond, slot) do { if (kDbg && threadIdx.x == 0 && (cond)) x_cstamp(dbg, slot); } while (0)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = blockIdx.x, row0 = blockIdx.y * X_BM;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* stages = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* qtile = stages + (size_t)X_STAGES * X_STAGE_BYTES;
    XBarriers* bar = reinterpret_cast<XBarriers*>(qtile + X_QTILE_BYTES);
    const int nrows = (P.B - row0 < X_BM) ? (P.B - row0) : X_BM;

    const int Ki = (int)P.K;                                          // K < 2^31 - 128 (checked by the launcher)
    const int ntiles = (Ki + X_BN - 1) / X_BN;
    // tiles are dealt round-robin over the splits (see infonce_tc.cu): at any moment the launch reads one moving window
    const int n = (split < ntiles) ? (ntiles - split + P.nsplit - 1) / P.nsplit : 0;

    X_STAMP(0);
    pdl_launch_dependents();

    if (warp == X_WARP_TMA && lane == 0) {
        prefetch_tmap(&tmap);
        prefetch_tmap(&qmap);
        for (int s = 0; s < X_STAGES; ++s) { mbar_init(&bar->full[s], 1); mbar_init(&bar->empty[s], 1); }
        for (int b = 0; b < X_NSBUF; ++b) { mbar_init(&bar->s_full[b], 1); mbar_init(&bar->p_full[b], X_SM_WARPS); }
        mbar_init(&bar->acc_final, 1);
        mbar_init(&bar->q_ready, 1);
        bar->redo = 0u;
        fence_barrier_init();
    }
    if (warp == X_WARP_MMA) tmem_alloc<X_TM_COLS>(&bar->tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bar->tmem_base;
    X_STAMP(1);

    // softmax thread state (warps 0..15)
    const int wq = warp & 3, cg = (warp >> 2) & 3;                // TMEM lane quarter, key group (32 keys of every tile)
    const int r_loc = wq * 32 + lane;
    const int row = row0 + r_loc;
    const uint32_t lane_addr = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(cg * 32);
    float m_ref = m_ref0;                                          // reference exponent (log2 domain) of this row
    float pos_l2 = 0.f, pos_nat0 = 0.f;                            // positive logit: log2 domain (like the S tile) / natural
    float s_run = 0.f;
    int cnt = 0;

    // Up to three passes over the CTA's tiles: sweep; if a row left the window: max-only pass + sweep with the exact row max.
    // `base` = tiles handled by earlier passes: every mbarrier keeps counting phases across passes.
    int base = 0;
    int npass = 0;
    for (int pass = X_PASS_SWEEP; ; ) {
        const bool sweep = (pass == X_PASS_SWEEP);
        if (warp < X_SM_WARPS) {
            // =========================================================================================== softmax warps
            if (npass == 0) {
                if (!P.no_prep_wait) pdl_wait();                      // prep kernel results are visible from here (late trigger:
                pos_nat0 = __ldcg(P.pos_ws + row);                    // they were before this kernel could launch).  q.k / T
                pos_l2 = pos_nat0 * 1.4426950408889634f;
                if (split == 0 && blockIdx.y == 0 && threadIdx.x == 0) { for (int w = 0; w < 6; ++w) P.counter[w] = 0u; }   // re-arm the finalize control block
                X_STAMP(3);
            }
            const float p_pos = ex2(pos_l2 - m_ref) * 0.98f;         // a step whose row sum stays below this has no key > pos
            const uint64_t nm2 = x2::pack(-m_ref, -m_ref);
            const bool zero_ref = (m_ref0 == 0.f) && (npass == 0);   // uniform: the normal first sweep subtracts nothing
            float tmax = -INFINITY;
            // rank count state (warp-uniform): 0 = no step needed it yet, 1 = counting inside the sweep, 2 = every row of the
            // warp is past rank_cap (callers that only want top-k hits): nothing left to learn
            int count_st = 0;
            s_run = 0.f; cnt = 0;
            int sb = base % X_NSBUF;
            uint32_t sph = (uint32_t)(base / X_NSBUF) & 1u;
            int key0 = split * X_BN + cg * 32;                        // first key of this thread's 32-key slice of tile v
            const int kstep = P.nsplit * X_BN;
            for (int v = 0; v < n; ++v, key0 += kstep) {
                const uint32_t s_addr = lane_addr + xtm_s(sb);
                X_CSTAMP(v == 2, 16);
                mbar_wait(&bar->s_full[sb], sph);
                tc_fence_after();
                if (kDbg && v == 0 && npass == 0) X_STAMP(4);
                X_CSTAMP(v == 2, 17);
                uint32_t sr[32];
                tmem_ld32(s_addr, sr);
                tc_wait_ld();
                X_CSTAMP(v == 2, 18);
                float* sv = reinterpret_cast<float*>(sr);
                const int nvalid = Ki - key0;                         // >= 32: all of this thread's keys exist
                if (nvalid < 32) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) if (j >= nvalid) sv[j] = -INFINITY;    // TMA zero-filled rows past K
                }
                if (!sweep) {
                    // max-only pass: row max of the raw dot products
                    float t0 = -INFINITY, t1 = -INFINITY;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) { t0 = max3(t0, sv[j], sv[j + 1]); t1 = max3(t1, sv[j + 2], sv[j + 3]); }
                    tmax = fmaxf(tmax, fmaxf(t0, t1));
                } else {
                    // exponentials, row sum, bf16 pack (+ rank count once a step could not rule out a key above the positive:
                    // from then on the count rides inside the sweep for the rest of the pass)
                    uint32_t pk[16];
                    float rs;
                    if (count_st == 1) {
                        rs = zero_ref ? sweep32<true, true>(sv, nm2, pos_l2, pk, cnt) : sweep32<true, false>(sv, nm2, pos_l2, pk, cnt);
                        if (rank_cap > 0 && __all_sync(0xffffffffu, cnt >= rank_cap)) count_st = 2;
                    } else {
                        rs = zero_ref ? sweep32<false, true>(sv, nm2, pos_l2, pk, cnt) : sweep32<false, false>(sv, nm2, pos_l2, pk, cnt);
#ifndef GCA_X_NOCOUNT                                          /* (bring-up switch: timing without the rank count) */
                        if (count_st == 0 && __any_sync(0xffffffffu, !(rs < p_pos))) { cnt += count32(sv, pos_l2); count_st = 1; }
#endif
                    }
                    s_run += rs;
                    X_CSTAMP(v == 2, 19);
                    X_CSTAMP(v == 2, 20);
                    tmem_st16(s_addr, pk);                                // P (bf16, 32 keys) over the first 16 S columns
                    tc_wait_st();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar->p_full[sb]);
                if (++sb == X_NSBUF) { sb = 0; sph ^= 1u; }
                X_CSTAMP(v == 2, 21);
                X_CSTAMP(v == 3, 22);
            }
            if (kDbg && npass == 0) X_STAMP(5);
            // ---- cross-group exchange: row sum and rank count (sweep) or row max (max pass) of the four key groups
            bar->xs[cg][r_loc] = sweep ? s_run : tmax;
            bar->xc[cg][r_loc] = cnt;
            named_barrier_sync(1, X_SM_THREADS);
            const float e0 = bar->xs[0][r_loc], e1 = bar->xs[1][r_loc], e2 = bar->xs[2][r_loc], e3 = bar->xs[3][r_loc];
            if (sweep) {
                s_run = (e0 + e1) + (e2 + e3);                            // fixed order: identical in all four threads of a row
                cnt = (bar->xc[0][r_loc] + bar->xc[1][r_loc]) + (bar->xc[2][r_loc] + bar->xc[3][r_loc]);
                if (npass == 0 && n > 0 && !(s_run >= X_WIN_LO && s_run <= X_WIN_HI)) bar->redo = 1u;
            } else {
                const float mx = fmaxf(fmaxf(e0, e1), fmaxf(e2, e3));
                m_ref = (mx == -INFINITY || !(mx == mx)) ? 0.f : mx;         // exact row max of this CTA's keys: every p <= 1
            }
        } else if (warp == X_WARP_TMA) {
            // =========================================================================================== TMA producer
            if (lane == 0) {
                auto load_tile = [&](int v) {
                    const int gt = base + v, stage = gt % X_STAGES;
                    uint8_t* dst = stages + (size_t)stage * X_STAGE_BYTES;
                    const int key0 = (split + v * P.nsplit) * X_BN;
                    // pacing at the head of the stream: the first X_TMA_LEAD tiles go out at once, the next few wait for a
                    // landed tile each (one tile per SM already covers the HBM latency-bandwidth product; a deeper initial
                    // burst only queues every CTA's first tile behind everybody's later ones); then the ring runs free.
                    if (v >= X_TMA_LEAD && v < 2 * X_TMA_LEAD) { const int gp = gt - X_TMA_LEAD; mbar_wait(&bar->full[gp % X_STAGES], (gp / X_STAGES) & 1); }
                    if (gt >= X_STAGES) mbar_wait(&bar->empty[stage], ((gt / X_STAGES) - 1) & 1);
                    mbar_arrive_expect_tx(&bar->full[stage], X_STAGE_BYTES);
                    tma_load_2d(dst, &tmap, &bar->full[stage], 0, key0);
                    tma_load_2d(dst + X_HALF_BYTES, &tmap, &bar->full[stage], 64, key0);
                };
                int v0 = 0;
                if (npass == 0) {
                    // queue tiles do not depend on the prep kernel: get the first ones moving before waiting for it
                    if (kDbg) x_stamp(dbg, 9);
                    const int pre = n < X_TMA_LEAD ? n : X_TMA_LEAD;
                    for (; v0 < pre; ++v0) load_tile(v0);
                    if (!P.no_prep_wait) pdl_wait();                      // q_bf16 comes from the prep kernel
                    mbar_arrive_expect_tx(&bar->q_ready, (uint32_t)X_QTILE_BYTES);
                    tma_load_2d(qtile, &qmap, &bar->q_ready, 0, row0);
                    tma_load_2d(qtile + X_HALF_BYTES, &qmap, &bar->q_ready, 64, row0);
                }
                for (int v = v0; v < n; ++v) load_tile(v);
            }
        } else {
            // =========================================================================================== MMA issuer
            const bool leader = elect_one();
            constexpr uint32_t idesc_s = make_idesc_bf16(X_BM, X_BN, 0, 0);   // S: A = q tile, B = 128 queue rows, both K-major
            constexpr uint32_t idesc_o = make_idesc_bf16(X_BM, X_D, 0, 1);    // O: A = P (TMEM), B = the same rows, MN-major
            if (npass == 0) {
                mbar_wait(&bar->q_ready, 0);
                tc_fence_after();
                if (kDbg && leader) x_stamp(dbg, 10);
            }
            const uint32_t qbase = smem_u32(qtile);
            auto issue_s = [&](int v) {
                const int gt = base + v, stage = gt % X_STAGES;
                mbar_wait(&bar->full[stage], (gt / X_STAGES) & 1);
                tc_fence_after();
                if (kDbg && v == 0 && npass == 0 && leader) x_stamp(dbg, 11);
                const uint32_t sbase = smem_u32(stages + (size_t)stage * X_STAGE_BYTES);
                const uint32_t d_tmem = tmem + xtm_s(gt % X_NSBUF);
                if (leader) {
#pragma unroll
                    for (int kk = 0; kk < X_D / 16; ++kk) {
                        const uint32_t koff = (kk >> 2) * X_HALF_BYTES + (kk & 3) * 32;
                        mma_ss(d_tmem, make_smem_desc_sw128(qbase + koff, 16, 1024), make_smem_desc_sw128(sbase + koff, 16, 1024),
                               idesc_s, kk > 0);
                    }
                    tc_commit(&bar->s_full[gt % X_NSBUF]);
                    if (!sweep) tc_commit(&bar->empty[stage]);                // max pass: the tile is not needed again
                }
                __syncwarp();
            };
            for (int v = 0; v < n && v < X_NSBUF; ++v) issue_s(v);
            for (int v = 0; v < n; ++v) {
                const int gt = base + v, stage = gt % X_STAGES, sb = gt % X_NSBUF;
                if (kDbg && v == 3 && leader) x_cstamp(dbg, 25);
                mbar_wait(&bar->p_full[sb], (gt / X_NSBUF) & 1);              // all 16 softmax warps are done with S buffer sb
                tc_fence_after();
                if (kDbg && v == 3 && leader) x_cstamp(dbg, 26);
                if (sweep) {
                    const uint32_t sbase = smem_u32(stages + (size_t)stage * X_STAGE_BYTES);
                    const uint32_t p_tmem = tmem + xtm_s(sb);
                    if (leader) {
#pragma unroll
                        for (int kk = 0; kk < X_BN / 16; ++kk) {
                            // 16 keys per MMA: P of key group kk/2 sits in the first 16 columns of its 32-column S block;
                            // B = two 8-row swizzle atoms (2 KB), LBO = next 64-feature box, SBO = next 8 keys
                            mma_ts(tmem + X_TM_O, p_tmem + (uint32_t)((kk >> 1) * 32 + (kk & 1) * 8),
                                   make_smem_desc_sw128(sbase + kk * 2048, X_HALF_BYTES, 1024), idesc_o, (v > 0 || kk > 0) ? 1u : 0u);
                        }
                        tc_commit(&bar->empty[stage]);
                        if (v == n - 1) tc_commit(&bar->acc_final);
                    }
                    __syncwarp();
                }
                if (kDbg && v == 3 && leader) x_cstamp(dbg, 27);
                if (v + X_NSBUF < n) issue_s(v + X_NSBUF);
                if (kDbg && v == 3 && leader) x_cstamp(dbg, 28);
            }
        }
        // ---- pass control (all 18 warps)
        base += n;
        ++npass;
        __syncthreads();
        if (pass == X_PASS_SWEEP) {
            if (npass == 1 && bar->redo != 0u) { pass = X_PASS_MAX; continue; }
            break;
        }
        pass = X_PASS_SWEEP;                                               // after the max pass: sweep again
    }

    if (warp < X_SM_WARPS) {
        // =============================================================================================== epilogue
        if (cg == 0) {
            // one thread per row: the split statistics
            const size_t po = part_stat_index(split, row, P.nsplit);
            const float m_nat = m_ref * 0.6931471805599453f;
            P.part_max[po] = (n > 0) ? m_nat : -INFINITY;
            P.part_sum[po] = (n > 0) ? s_run : 0.f;
            P.part_cnt[po] = cnt;
            // rows whose loss could leave the packed fixed-point word of the finalize kernel (control word 6, see finalize.cu)
            if (n > 0 && row < P.B && (m_nat + __logf(s_run)) - pos_nat0 > 2.125f * P.inv_T + 30.f) P.counter[6] = 1u;
        }
        const int nsweep = (npass == 1) ? 1 : 2;
        if (n > 0) { mbar_wait(&bar->acc_final, (nsweep - 1) & 1); tc_fence_after(); }
        X_STAMP(6);
        // O: this warp's 32 rows x 32 feature columns -> padded smem rows (the tile ring is idle now), then one 512-byte bulk
        // store per row (TMA engine; fully coalesced, no LSU work)
        float* ost = reinterpret_cast<float*>(stages);
        {
            uint32_t a[32];
            if (n > 0) { tmem_ld32(lane_addr + X_TM_O, a); tc_wait_ld(); }
            else {
#pragma unroll
                for (int j = 0; j < 32; ++j) a[j] = 0u;
            }
            float4* dst = reinterpret_cast<float4*>(ost + (size_t)r_loc * X_OST_STRIDE + cg * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                dst[j] = make_float4(__uint_as_float(a[4 * j]), __uint_as_float(a[4 * j + 1]), __uint_as_float(a[4 * j + 2]),
                                     __uint_as_float(a[4 * j + 3]));
        }
        fence_proxy_async();                                               // generic-proxy smem writes -> bulk (async proxy) reads
        named_barrier_sync(1, X_SM_THREADS);
        if (cg == 0 && r_loc < nrows) {
            bulk_store(P.part_acc + ((size_t)split * P.Bpad + row) * X_D, ost + (size_t)r_loc * X_OST_STRIDE, X_D * 4);
            bulk_commit_wait_read();                                       // the smem rows must outlive the copy engine's reads
        }
    }
    X_STAMP(7);
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == X_WARP_MMA) tmem_dealloc<X_TM_COLS>(tmem);
    X_STAMP(8);
#undef X_STAMP
#undef X_CSTAMP
}

// ------------------------------------------------------------------------------------------------ host side
int get_queue_tmap(const void* queue, long long K, CUtensorMap* out);      // infonce_tc.cu
unsigned long long* debug_timebuf();

bool infonce_tcx_enabled()
{
    static int on = -1;                               // GCA_TC_LEGACY=1 keeps the first-generation stream kernel (A/B timing)
    if (on < 0) { const char* e = getenv("GCA_TC_LEGACY"); on = (e && e[0] == '1') ? 0 : 1; }
    return on != 0;
}

// the stream kernel alone (the prep kernel has been launched by the caller, infonce_tc_launch)
int infonce_tcx_launch(const InfoNceStreamParams& P, cudaStream_t st)
{
    CUtensorMap tmap, qmap;
    int rc = get_queue_tmap(P.queue, P.K, &tmap);
    if (rc != GCA_OK) return rc;
    rc = get_queue_tmap(P.q_bf16_ws, P.Bpad, &qmap);
    if (rc != GCA_OK) return rc;
    const XDebug dbg{debug_timebuf()};
    // reference exponent of the first sweep (log2 units): 0 covers every unit-row logit while 1/T <= 60 log2 units
    const float c2 = P.inv_T * 1.4426950408889634f;
    const float m_ref0 = c2 > 100.f ? c2 - 100.f : 0.f;
    const int rank_cap = P.rank_cap;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(P.nsplit, P.Bpad / X_BM); cfg.blockDim = dim3(X_THREADS); cfg.dynamicSmemBytes = X_SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    if (dbg.timebuf) {
        GCA_CUDA(cudaFuncSetAttribute(infonce_tcx_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)X_SMEM_BYTES));
        GCA_CUDA(launch_ex(&cfg, infonce_tcx_kernel<true>, tmap, qmap, P, dbg, m_ref0, rank_cap));
    } else {
        GCA_CUDA(cudaFuncSetAttribute(infonce_tcx_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)X_SMEM_BYTES));
        GCA_CUDA(launch_ex(&cfg, infonce_tcx_kernel<false>, tmap, qmap, P, dbg, m_ref0, rank_cap));
    }
    GCA_LAUNCH_CHECK("infonce_tcx_kernel");
    return GCA_OK;
}

}  // namespace gca
