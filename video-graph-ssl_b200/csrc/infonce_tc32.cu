// K1y: the fp32 parity mode of the InfoNCE head on the 5th-generation tensor cores (GCA_ALGO_TC32): fp32 queue, d = 128.
//
// The reference computes its logits in fp32 (mem_moco.py:42-46); north_star asks the fp32 mode for a loss within 1e-5 and
// gradients within 1e-2.  The logits therefore get fp32-grade accuracy out of bf16 MMAs (the exact three-way split of
// sim_tc.cu: x = h + m + l, 8 + 8 + 8 mantissa bits; the six piece products >= 2^-24 of the full product -- l.h, h.l, m.m,
// m.h, h.m, h.h, smallest first -- accumulate in one fp32 TMEM accumulator; the three dropped ones are below fp32 rounding),
// the softmax statistics stay fp32, and the gradient accumulation O += P Q runs on two pieces of each factor (P = Ph + Pm
// split by the softmax warps, the h and m planes of the queue; Ph.Qh + Ph.Qm + Pm.Qh, error ~2^-17: dq within ~1e-5).
//
// Structure = infonce_tcx.cu (fixed per-row reference exponent verified after the sweep, one O accumulator, three S buffers,
// 16 softmax warps, single pass for loss + gradient) on 64-key tiles: a tile is three [64 keys x 128] bf16 planes; the l
// plane lives in a 2-deep ring (free again as soon as S of its tile is done), the h and m planes -- which O += P Q also reads
// -- in a 3-deep ring, so the reload of a slot never sits between two GEMMs; the q block is three [128 x 128] planes (96 KB).
// S = 48 tcgen05.mma (M128 N64 K16) per tile, O = 12 (M128 N128 K16).
// Launch sequence: queue split (fp32 [K,128] -> planes in the workspace) -> prep (q planes, positives, staged k) -> this.
// Replaces mem_moco.py:36-46 + criterion.py:44 + autograd(mm) + metric.py:44-67 of the reference in its own precision.
#include "gca_common.cuh"
#include "infonce_params.cuh"
#include "tc_ptx.cuh"
#include <stdlib.h>

namespace gca {

constexpr int Y_BM = 128, Y_BN = 64, Y_D = 128;
constexpr int Y_ML_STAGES = 2, Y_H_STAGES = 3;            // rings: "ML" = the l plane alone, "H" = the h + m planes
constexpr int Y_KBOX = Y_BN * 64 * 2;                       // 8 KB: [64 keys][64 features] swizzled box
constexpr int Y_KPLANE = 2 * Y_KBOX;                        // 16 KB: one plane of a tile
constexpr int Y_ML_BYTES = Y_KPLANE;                        // 16 KB: l plane
constexpr int Y_HM_BYTES = 2 * Y_KPLANE;                    // 32 KB: h plane, m plane
constexpr int Y_RING_BYTES = Y_ML_STAGES * Y_ML_BYTES + Y_H_STAGES * Y_HM_BYTES;    // 128 KB
constexpr int Y_QBOX = Y_BM * 64 * 2;                       // 16 KB
constexpr int Y_QPLANE = 2 * Y_QBOX;                        // 32 KB
constexpr size_t Y_QTILE_BYTES = 3 * (size_t)Y_QPLANE;      // 96 KB
constexpr int Y_SM_WARPS = 16;                              // softmax warps: lane quarter = warp & 3, key group = warp >> 2 (16 keys)
constexpr int Y_SM_THREADS = Y_SM_WARPS * 32;
constexpr int Y_WARP_TMA = 16, Y_WARP_MMA = 17;
constexpr int Y_THREADS = 18 * 32;
constexpr int Y_NSBUF = 3;                                  // S buffers of 64 TMEM columns
constexpr uint32_t Y_TM_COLS = 512;
__host__ __device__ constexpr uint32_t ytm_s(int b) { return (uint32_t)(b * 64); }
constexpr uint32_t Y_TM_O = 256;
constexpr int Y_OST_STRIDE = 132;
constexpr float Y_WIN_LO = 8.6736174e-19f;                  // 2^-60
constexpr float Y_WIN_HI = 1.2676506e30f;                   // 2^100

struct YBarriers {
    uint64_t full_ml[Y_ML_STAGES], empty_ml[Y_ML_STAGES];     // l plane of a tile landed / S of the tile is done
    uint64_t full_h[Y_H_STAGES], empty_h[Y_H_STAGES];         // h + m planes landed / S and O of the tile are done
    uint64_t s_full[Y_NSBUF], p_full[Y_NSBUF];
    uint64_t acc_final, q_ready;
    uint32_t tmem_base;
    uint32_t redo;
};
// cross-key-group exchange (row sums / maxima, rank counts): aliased onto the l ring, which only the S GEMMs read -- all of
// them have completed when the softmax warps leave their tile loop
struct YExchange { float xs[4][Y_BM]; int xc[4][Y_BM]; };
static_assert(sizeof(YExchange) <= Y_ML_BYTES, "exchange area must fit in one l slot");
constexpr size_t Y_SMEM_BYTES = 1024 + (size_t)Y_RING_BYTES + Y_QTILE_BYTES + sizeof(YBarriers) + 64;
static_assert((size_t)Y_BM * Y_OST_STRIDE * 4 <= (size_t)Y_RING_BYTES, "O staging must fit in the tile rings");
static_assert(Y_SMEM_BYTES <= 232448, "shared memory budget");

namespace y2 {
__device__ __forceinline__ uint64_t pack(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t add(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
}  // namespace y2

__device__ __forceinline__ void y_tmem_ld16(uint32_t taddr, uint32_t* r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void y_tmem_st16(uint32_t taddr, const uint32_t* r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
                 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 :: "r"(taddr),
                    "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
__device__ __forceinline__ void y_bulk_store(void* gdst, const void* smem_src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(reinterpret_cast<uint64_t>(gdst)), "r"(ptx::smem_u32(smem_src)), "r"(bytes) : "memory");
}

// exponentials of 16 logits (log2 domain), fp32 row sum, bf16 pack; count = #logits above the positive
// pk[0..8) = Ph (bf16 pairs), pk[8..16) = Pm = bf16(p - Ph): the two pieces of P that feed O += P Q
template <bool kZeroRef>
__device__ __forceinline__ float y_sweep16(const float* sv, float m_ref, uint32_t* pk)
{
    uint64_t acc0 = y2::pack(0.f, 0.f), acc1 = acc0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float xa = kZeroRef ? sv[2 * i] : sv[2 * i] - m_ref, xb = kZeroRef ? sv[2 * i + 1] : sv[2 * i + 1] - m_ref;
        const float pa = ptx::ex2(xa), pb = ptx::ex2(xb);
        if (i & 1) acc1 = y2::add(acc1, y2::pack(pa, pb)); else acc0 = y2::add(acc0, y2::pack(pa, pb));
        const uint32_t ph = ptx::pack_bf16(pa, pb);
        pk[i] = ph;
        pk[8 + i] = ptx::pack_bf16(pa - __uint_as_float(ph << 16), pb - __uint_as_float(ph & 0xffff0000u));   // exact residuals
    }
    float a0, a1, b0, b1;
    y2::unpack(acc0, a0, a1);
    y2::unpack(acc1, b0, b1);
    return (a0 + a1) + (b0 + b1);
}
__device__ __forceinline__ int y_count16(const float* sv, float pos_l2)
{
    int c = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) c += (sv[i] > pos_l2) ? 1 : 0;
    return c;
}

enum { Y_PASS_SWEEP = 0, Y_PASS_MAX = 1 };

__global__ void __launch_bounds__(Y_THREADS, 1)
infonce_tc32_kernel(const __grid_constant__ CUtensorMap kmap, const __grid_constant__ CUtensorMap qmap,
                    const InfoNceStreamParams P, const float m_ref0)
{
    using namespace ptx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = blockIdx.x, row0 = blockIdx.y * Y_BM;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* stages = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* hring = stages + (size_t)Y_ML_STAGES * Y_ML_BYTES;         // `stages` = the m/l ring, then the h ring
    uint8_t* qtile = stages + (size_t)Y_RING_BYTES;
    YBarriers* bar = reinterpret_cast<YBarriers*>(qtile + Y_QTILE_BYTES);
    YExchange* xch = reinterpret_cast<YExchange*>(stages);
    const int nrows = (P.B - row0 < Y_BM) ? (P.B - row0) : Y_BM;
    const int Ki = (int)P.K;
    const int ntiles = (Ki + Y_BN - 1) / Y_BN;
    const int n = (split < ntiles) ? (ntiles - split + P.nsplit - 1) / P.nsplit : 0;     // round-robin tile dealing

    pdl_launch_dependents();
    if (warp == Y_WARP_TMA && lane == 0) {
        prefetch_tmap(&kmap);
        prefetch_tmap(&qmap);
        for (int s = 0; s < Y_ML_STAGES; ++s) { mbar_init(&bar->full_ml[s], 1); mbar_init(&bar->empty_ml[s], 1); }
        for (int s = 0; s < Y_H_STAGES; ++s) { mbar_init(&bar->full_h[s], 1); mbar_init(&bar->empty_h[s], 1); }
        for (int b = 0; b < Y_NSBUF; ++b) { mbar_init(&bar->s_full[b], 1); mbar_init(&bar->p_full[b], Y_SM_WARPS); }
        mbar_init(&bar->acc_final, 1);
        mbar_init(&bar->q_ready, 1);
        bar->redo = 0u;
        fence_barrier_init();
    }
    if (warp == Y_WARP_MMA) tmem_alloc<Y_TM_COLS>(&bar->tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bar->tmem_base;

    const int wq = warp & 3, cg = (warp >> 2) & 3;
    const int r_loc = wq * 32 + lane;
    const int row = row0 + r_loc;
    const uint32_t lane_addr = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(cg * 16);
    float m_ref = m_ref0;
    float pos_l2 = 0.f, pos_nat0 = 0.f;
    float s_run = 0.f;
    int cnt = 0;

    // the queue planes (split kernel) and the q planes / positives (prep kernel) come from the two launches in front
    pdl_wait();

    int base = 0, npass = 0;
    for (int pass = Y_PASS_SWEEP; ; ) {
        const bool sweep = (pass == Y_PASS_SWEEP);
        if (warp < Y_SM_WARPS) {
            // =========================================================================================== softmax warps
            if (npass == 0) {
                pos_nat0 = P.pos_ws[row];
                pos_l2 = pos_nat0 * 1.4426950408889634f;
                if (split == 0 && blockIdx.y == 0 && threadIdx.x == 0) { for (int w = 0; w < 6; ++w) P.counter[w] = 0u; }
            }
            const bool zero_ref = (m_ref0 == 0.f) && (npass == 0);
            float tmax = -INFINITY;
            s_run = 0.f; cnt = 0;
            int sb = base % Y_NSBUF;
            uint32_t sph = (uint32_t)(base / Y_NSBUF) & 1u;
            int key0 = split * Y_BN + cg * 16;
            const int kstep = P.nsplit * Y_BN;
            for (int v = 0; v < n; ++v, key0 += kstep) {
                const uint32_t s_addr = lane_addr + ytm_s(sb);
                mbar_wait(&bar->s_full[sb], sph);
                tc_fence_after();
                uint32_t sr[16];
                y_tmem_ld16(s_addr, sr);
                tc_wait_ld();
                float* sv = reinterpret_cast<float*>(sr);
                const int nvalid = Ki - key0;
                if (nvalid < 16) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (j >= nvalid) sv[j] = -INFINITY;      // rows past K (next plane / zero fill)
                }
                if (!sweep) {
                    float t0 = -INFINITY, t1 = -INFINITY;
#pragma unroll
                    for (int j = 0; j < 16; j += 4) { t0 = max3(t0, sv[j], sv[j + 1]); t1 = max3(t1, sv[j + 2], sv[j + 3]); }
                    tmax = fmaxf(tmax, fmaxf(t0, t1));
                } else {
                    uint32_t pk[16];
                    cnt += y_count16(sv, pos_l2);                  // exact fp32 comparison against the fp32 positive
                    s_run += zero_ref ? y_sweep16<true>(sv, 0.f, pk) : y_sweep16<false>(sv, m_ref, pk);
                    y_tmem_st16(s_addr, pk);                       // Ph | Pm (bf16, 16 keys each) over the group's 16 S columns
                    tc_wait_st();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar->p_full[sb]);
                if (++sb == Y_NSBUF) { sb = 0; sph ^= 1u; }
            }
            xch->xs[cg][r_loc] = sweep ? s_run : tmax;
            xch->xc[cg][r_loc] = cnt;
            named_barrier_sync(1, Y_SM_THREADS);
            const float e0 = xch->xs[0][r_loc], e1 = xch->xs[1][r_loc], e2 = xch->xs[2][r_loc], e3 = xch->xs[3][r_loc];
            if (sweep) {
                s_run = (e0 + e1) + (e2 + e3);
                cnt = (xch->xc[0][r_loc] + xch->xc[1][r_loc]) + (xch->xc[2][r_loc] + xch->xc[3][r_loc]);
                if (npass == 0 && n > 0 && !(s_run >= Y_WIN_LO && s_run <= Y_WIN_HI)) bar->redo = 1u;
            } else {
                const float mx = fmaxf(fmaxf(e0, e1), fmaxf(e2, e3));
                m_ref = (mx == -INFINITY || !(mx == mx)) ? 0.f : mx;
            }
        } else if (warp == Y_WARP_TMA) {
            // =========================================================================================== TMA producer
            if (lane == 0) {
                if (npass == 0) {
                    mbar_arrive_expect_tx(&bar->q_ready, (uint32_t)Y_QTILE_BYTES);
#pragma unroll
                    for (int p = 0; p < 3; ++p) {          // planes are stacked along the row axis of the tensor maps
                        tma_load_2d(qtile + p * Y_QPLANE, &qmap, &bar->q_ready, 0, p * P.Bpad + row0);
                        tma_load_2d(qtile + p * Y_QPLANE + Y_QBOX, &qmap, &bar->q_ready, 64, p * P.Bpad + row0);
                    }
                }
                for (int v = 0; v < n; ++v) {
                    const int gt = base + v, sm_ = gt % Y_ML_STAGES, sh = gt % Y_H_STAGES;
                    const int key0 = (split + v * P.nsplit) * Y_BN;
                    uint8_t* dl = stages + (size_t)sm_ * Y_ML_BYTES;
                    uint8_t* dh = hring + (size_t)sh * Y_HM_BYTES;
                    // planes (h, m, l) are stacked along the row axis of the tensor map
                    if (gt >= Y_ML_STAGES) mbar_wait(&bar->empty_ml[sm_], ((gt / Y_ML_STAGES) - 1) & 1);
                    mbar_arrive_expect_tx(&bar->full_ml[sm_], Y_ML_BYTES);
                    tma_load_2d(dl, &kmap, &bar->full_ml[sm_], 0, 2 * Ki + key0);
                    tma_load_2d(dl + Y_KBOX, &kmap, &bar->full_ml[sm_], 64, 2 * Ki + key0);
                    if (gt >= Y_H_STAGES) mbar_wait(&bar->empty_h[sh], ((gt / Y_H_STAGES) - 1) & 1);
                    mbar_arrive_expect_tx(&bar->full_h[sh], Y_HM_BYTES);
#pragma unroll
                    for (int p = 0; p < 2; ++p) {
                        tma_load_2d(dh + p * Y_KPLANE, &kmap, &bar->full_h[sh], 0, p * Ki + key0);
                        tma_load_2d(dh + p * Y_KPLANE + Y_KBOX, &kmap, &bar->full_h[sh], 64, p * Ki + key0);
                    }
                }
            }
        } else if (warp == Y_WARP_MMA) {
            // =========================================================================================== MMA issuer
            const bool leader = elect_one();
            constexpr uint32_t idesc_s = make_idesc_bf16(Y_BM, Y_BN, 0, 0);   // S: A = q plane, B = 64 queue rows, both K-major
            constexpr uint32_t idesc_o = make_idesc_bf16(Y_BM, Y_D, 0, 1);    // O: A = P (TMEM), B = the h plane rows, MN-major
            if (npass == 0) { mbar_wait(&bar->q_ready, 0); tc_fence_after(); }
            const uint32_t qbase = smem_u32(qtile);
            auto issue_s = [&](int v) {
                const int gt = base + v, sm_ = gt % Y_ML_STAGES, sh = gt % Y_H_STAGES;
                const uint32_t lbase = smem_u32(stages + (size_t)sm_ * Y_ML_BYTES);
                const uint32_t hbase = smem_u32(hring + (size_t)sh * Y_HM_BYTES);          // h plane, then m plane
                const uint32_t d_tmem = tmem + ytm_s(gt % Y_NSBUF);
                // six (q piece, queue piece) products, the big one last: h.l  |  m.m, h.m, l.h, m.h, h.h
                // (q pieces: 0 = h, 1 = m, 2 = l)
                mbar_wait(&bar->full_ml[sm_], (gt / Y_ML_STAGES) & 1);
                tc_fence_after();
                if (leader) {
#pragma unroll
                    for (int kk = 0; kk < Y_D / 16; ++kk) {
                        const uint32_t qoff = (kk >> 2) * Y_QBOX + (kk & 3) * 32;
                        const uint32_t koff = (kk >> 2) * Y_KBOX + (kk & 3) * 32;
                        mma_ss(d_tmem, make_smem_desc_sw128(qbase + qoff, 16, 1024), make_smem_desc_sw128(lbase + koff, 16, 1024),
                               idesc_s, kk > 0 ? 1u : 0u);
                    }
                }
                __syncwarp();
                mbar_wait(&bar->full_h[sh], (gt / Y_H_STAGES) & 1);
                tc_fence_after();
                if (leader) {
                    constexpr int PA[5] = {1, 0, 2, 1, 0}, PB[5] = {1, 1, 0, 0, 0};      // queue plane in the slot: 0 = h, 1 = m
#pragma unroll
                    for (int t = 0; t < 5; ++t) {
#pragma unroll
                        for (int kk = 0; kk < Y_D / 16; ++kk) {
                            const uint32_t qoff = PA[t] * Y_QPLANE + (kk >> 2) * Y_QBOX + (kk & 3) * 32;
                            const uint32_t koff = PB[t] * Y_KPLANE + (kk >> 2) * Y_KBOX + (kk & 3) * 32;
                            mma_ss(d_tmem, make_smem_desc_sw128(qbase + qoff, 16, 1024), make_smem_desc_sw128(hbase + koff, 16, 1024),
                                   idesc_s, 1u);
                        }
                    }
                    tc_commit(&bar->s_full[gt % Y_NSBUF]);
                    tc_commit(&bar->empty_ml[sm_]);                            // S done: the l slot is free again
                    if (!sweep) tc_commit(&bar->empty_h[sh]);                  // max pass: no O GEMM follows
                }
                __syncwarp();
            };
            for (int v = 0; v < n && v < Y_ML_STAGES; ++v) issue_s(v);
            for (int v = 0; v < n; ++v) {
                const int gt = base + v, sh = gt % Y_H_STAGES, sb = gt % Y_NSBUF;
                mbar_wait(&bar->p_full[sb], (gt / Y_NSBUF) & 1);
                tc_fence_after();
                if (sweep) {
                    const uint32_t hbase = smem_u32(hring + (size_t)sh * Y_HM_BYTES);
                    const uint32_t p_tmem = tmem + ytm_s(sb);
                    if (leader) {
#pragma unroll
                        for (int kk = 0; kk < Y_BN / 16; ++kk) {
                            // 16 keys per MMA: key group kk keeps Ph in the first 8 columns of its 16-column S block and Pm in the
                            // last 8; B = two 8-row swizzle atoms (2 KB) of the h or m plane, LBO = next 64-feature box, SBO = next
                            // 8 keys.  Pm.Qh, Ph.Qm, Ph.Qh: the small products first
                            const uint32_t ph = p_tmem + (uint32_t)(kk * 16), pm = ph + 8u;
                            const uint64_t qh = make_smem_desc_sw128(hbase + kk * 2048, Y_KBOX, 1024);
                            const uint64_t qm = make_smem_desc_sw128(hbase + Y_KPLANE + kk * 2048, Y_KBOX, 1024);
                            mma_ts(tmem + Y_TM_O, pm, qh, idesc_o, (v > 0 || kk > 0) ? 1u : 0u);
                            mma_ts(tmem + Y_TM_O, ph, qm, idesc_o, 1u);
                            mma_ts(tmem + Y_TM_O, ph, qh, idesc_o, 1u);
                        }
                        tc_commit(&bar->empty_h[sh]);
                        if (v == n - 1) tc_commit(&bar->acc_final);
                    }
                    __syncwarp();
                }
                if (v + Y_ML_STAGES < n) issue_s(v + Y_ML_STAGES);
            }
        }
        base += n;
        ++npass;
        __syncthreads();
        if (pass == Y_PASS_SWEEP) {
            if (npass == 1 && bar->redo != 0u) { pass = Y_PASS_MAX; continue; }
            break;
        }
        pass = Y_PASS_SWEEP;
    }

    if (warp < Y_SM_WARPS) {
        // =============================================================================================== epilogue
        if (cg == 0) {
            const size_t po = part_stat_index(split, row, P.nsplit);
            const float m_nat = m_ref * 0.6931471805599453f;
            P.part_max[po] = (n > 0) ? m_nat : -INFINITY;
            P.part_sum[po] = (n > 0) ? s_run : 0.f;
            P.part_cnt[po] = cnt;
            if (n > 0 && row < P.B && (m_nat + __logf(s_run)) - pos_nat0 > 2.125f * P.inv_T + 30.f) P.counter[6] = 1u;
        }
        if (P.part_acc != nullptr) {
            const int nsweep = (npass == 1) ? 1 : 2;
            if (n > 0) { mbar_wait(&bar->acc_final, (nsweep - 1) & 1); tc_fence_after(); }
            float* ost = reinterpret_cast<float*>(stages);
            // O: this warp's 32 rows x 32 feature columns (feature block = key-group index) -> padded smem rows
            uint32_t a[32];
            if (n > 0) { tmem_ld32(tmem + ((uint32_t)(wq * 32) << 16) + Y_TM_O + (uint32_t)(cg * 32), a); tc_wait_ld(); }
            else {
#pragma unroll
                for (int j = 0; j < 32; ++j) a[j] = 0u;
            }
            float4* dst = reinterpret_cast<float4*>(ost + (size_t)r_loc * Y_OST_STRIDE + cg * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                dst[j] = make_float4(__uint_as_float(a[4 * j]), __uint_as_float(a[4 * j + 1]), __uint_as_float(a[4 * j + 2]),
                                     __uint_as_float(a[4 * j + 3]));
            fence_proxy_async();
            named_barrier_sync(1, Y_SM_THREADS);
            if (cg == 0 && r_loc < nrows) {
                y_bulk_store(P.part_acc + ((size_t)split * P.Bpad + row) * Y_D, ost + (size_t)r_loc * Y_OST_STRIDE, Y_D * 4);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == Y_WARP_MMA) tmem_dealloc<Y_TM_COLS>(tmem);
}

// fp32 [n, 128] -> three bf16 planes [3][n][128] (exact split x = h + m + l); scale != 1: the split is of x * scale
__global__ void __launch_bounds__(256)
split3_kernel(const float4* __restrict__ x, long long n4, long long plane4, float scale, uint2* __restrict__ planes)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(x + i);
        const float xs[4] = {v.x * scale, v.y * scale, v.z * scale, v.w * scale};
        __nv_bfloat16 h[4], m[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            h[j] = __float2bfloat16_rn(xs[j]);
            const float r1 = xs[j] - __bfloat162float(h[j]);           // exact
            m[j] = __float2bfloat16_rn(r1);
            const float r2 = r1 - __bfloat162float(m[j]);              // exact
            l[j] = __float2bfloat16_rn(r2);
        }
        planes[i] = *reinterpret_cast<const uint2*>(h);
        planes[plane4 + i] = *reinterpret_cast<const uint2*>(m);
        planes[2 * plane4 + i] = *reinterpret_cast<const uint2*>(l);
    }
}

// Once per step: q * (log2(e)/T) split into three bf16 planes [3][Bpad][128] (zero rows past B), positives q.k / T in fp32,
// k staged for the finalize kernel.  One warp per row.
__global__ void __launch_bounds__(256)
infonce_prep32_kernel(const float* __restrict__ q, const float* __restrict__ k, int B, int Bpad, float inv_T, float q_scale,
                      __nv_bfloat16* __restrict__ q_planes, float* __restrict__ pos_ws, float* __restrict__ pos_out,
                      float* __restrict__ k_hat, unsigned int* range_flag)
{
    ptx::pdl_launch_dependents();
    const int lane = threadIdx.x & 31, row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= Bpad) return;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (row < B) {
        a = __ldg(reinterpret_cast<const float4*>(q + (size_t)row * Y_D) + lane);
        b = __ldg(reinterpret_cast<const float4*>(k + (size_t)row * Y_D) + lane);
        reinterpret_cast<float4*>(k_hat + (size_t)row * Y_D)[lane] = b;
    }
    float dsum = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
    dsum = warp_sum(dsum) * inv_T;
    const float xs[4] = {a.x * q_scale, a.y * q_scale, a.z * q_scale, a.w * q_scale};
    __nv_bfloat16 h[4], m[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        h[j] = __float2bfloat16_rn(xs[j]);
        const float r1 = xs[j] - __bfloat162float(h[j]);
        m[j] = __float2bfloat16_rn(r1);
        l[j] = __float2bfloat16_rn(r1 - __bfloat162float(m[j]));
    }
    const size_t plane = (size_t)Bpad * Y_D, o = (size_t)row * Y_D + 4 * lane;
    *reinterpret_cast<uint2*>(q_planes + o) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(q_planes + plane + o) = *reinterpret_cast<const uint2*>(m);
    *reinterpret_cast<uint2*>(q_planes + 2 * plane + o) = *reinterpret_cast<const uint2*>(l);
    if (lane == 0) {
        pos_ws[row] = dsum;
        if (pos_out && row < B) pos_out[row] = dsum;
        if (fabsf(dsum) > 1.0625f * inv_T) range_flag[0] = 1u;
    }
}

// ------------------------------------------------------------------------------------------------ host side
int make_bf16_tmap(const void* base, long long rows, int box_rows, CUtensorMap* out);      // infonce_tc.cu

int infonce_tc32_nsplit(int B, long long K)
{
    const int nblk = infonce_bpad(B) / Y_BM;
    const long long ntiles = (K + Y_BN - 1) / Y_BN;
    int ns = sm_count_cached() / nblk;
    if (ns < 1) ns = 1;
    if (ns > ntiles) ns = (int)ntiles;
    return ns;
}

size_t infonce_tc32_extra_ws(int B, long long K) { return align_up((size_t)3 * K * Y_D * 2, 1024) + align_up((size_t)3 * infonce_bpad(B) * Y_D * 2, 1024); }

// P.q_bf16_ws is unused; `extra` = the tail of the workspace (infonce_tc32_extra_ws bytes, 1024-aligned): queue planes, q planes
int infonce_tc32_launch(const InfoNceStreamParams& P_, void* extra, cudaStream_t st)
{
    InfoNceStreamParams P = P_;
    if (P.d != Y_D) return set_err(GCA_ERR_UNSUPPORTED, "GCA_ALGO_TC32 needs d == %d (got %d)", Y_D, P.d);
    if (P.K >= (1ll << 29)) return set_err(GCA_ERR_UNSUPPORTED, "GCA_ALGO_TC32: K too large");
    if ((reinterpret_cast<uintptr_t>(P.queue) & 15) != 0) return set_err(GCA_ERR_BAD_ARG, "queue must be 16-byte aligned");
    if (P.lse_fixed || P.logits_out || P.xchg.mailboxes || P.normalize)
        return set_err(GCA_ERR_UNSUPPORTED, "GCA_ALGO_TC32 covers the fused forward (loss, gradient, rank) only");
    char* kplanes = (char*)extra;
    char* qplanes = kplanes + align_up((size_t)3 * P.K * Y_D * 2, 1024);
    P.q_scale = P.inv_T * 1.4426950408889634f;
    CUtensorMap kmap, qmap;
    int rc = make_bf16_tmap(kplanes, 3 * P.K, Y_BN, &kmap);
    if (rc != GCA_OK) return rc;
    rc = make_bf16_tmap(qplanes, 3ll * P.Bpad, Y_BM, &qmap);
    if (rc != GCA_OK) return rc;
    // 1. queue planes
    const long long n4 = P.K * Y_D / 4;
    long long blocks = (n4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    split3_kernel<<<(int)blocks, 256, 0, st>>>((const float4*)P.queue, n4, n4, 1.f, (uint2*)kplanes);
    GCA_LAUNCH_CHECK("split3_kernel");
    // 2. q planes, positives, staged keys
    infonce_prep32_kernel<<<(P.Bpad + 7) / 8, 256, 0, st>>>(P.q, P.k, P.B, P.Bpad, P.inv_T, P.q_scale, (__nv_bfloat16*)qplanes,
                                                            P.pos_ws, P.pos_out, P.k_hat, P.counter + 6);
    GCA_LAUNCH_CHECK("infonce_prep32_kernel");
    count_launch(2);
    // 3. the sweep
    const float c2 = P.q_scale;
    const float m_ref0 = c2 > 100.f ? c2 - 100.f : 0.f;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(P.nsplit, P.Bpad / Y_BM); cfg.blockDim = dim3(Y_THREADS); cfg.dynamicSmemBytes = Y_SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    GCA_CUDA(cudaFuncSetAttribute(infonce_tc32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Y_SMEM_BYTES));
    GCA_CUDA(cudaLaunchKernelEx(&cfg, infonce_tc32_kernel, kmap, qmap, P, m_ref0));
    GCA_LAUNCH_CHECK("infonce_tc32_kernel");
    return GCA_OK;
}

}  // namespace gca
