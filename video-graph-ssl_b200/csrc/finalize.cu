// Fixed-order merge of the per-split online-softmax partials written by the InfoNCE stream kernels, plus the
// K-shard (multi-GPU) combine / finish steps.  Everything here is deterministic: no floating-point atomics,
// every sum has a fixed association order.  Math: SURVEY.md Appendix A.1-A.4.
#include "gca_common.cuh"
#include "infonce_params.cuh"
#include "tc_ptx.cuh"
#include "launch_plan.cuh"
#include <stdlib.h>

namespace gca {

constexpr int FIN_THREADS = 256;
constexpr int FIN_COLS = 128;                 // feature columns per slab
constexpr int FIN_GROUPS = FIN_THREADS / FIN_COLS;
constexpr int FIN_MAX_SPLITS = 1024;
constexpr int FIN_STAT = 5;                   // split statistics a lane of warp 0 keeps in registers (<= 160 splits)
constexpr int FIN_CHUNK = 40;                 // gradient partials a thread keeps in flight at once
constexpr int FIN_VCH = 10;                   // vector path: 128-bit partial loads a thread keeps in flight (8 warps x 10 = 80 splits)

// last-block ticket: returns true in exactly one block, after every other block's global writes are visible
__device__ __forceinline__ bool last_block_ticket(unsigned int* counter, unsigned int nblocks, int* flag_smem)
{
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(counter, 1u);
        *flag_smem = (t == nblocks - 1);
    }
    __syncthreads();
    const bool last = (*flag_smem != 0);
    if (last) __threadfence();
    return last;
}

// Optional completion word in (pinned) host memory, armed per workspace by gca_workspace_set_done_flag: every CTA of a
// FIN_FULL launch -- row CTAs and enqueue CTAs -- orders its stores before a device-scope count; the CTA that takes the last
// count fences at system scope and publishes the step number.  A host thread polling that word sees a step complete (results stored, host inputs
// consumed) about a microsecond after the kernel's last store, without the driver's stream-synchronisation round trip.
__device__ __forceinline__ void publish_done(unsigned int* counter)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long p = *reinterpret_cast<volatile unsigned long long*>(counter + 12);
        if (p != 0ull) {
            __threadfence();                                   // this CTA's stores (device and host) are ordered before its count
            if (atomicAdd(counter + 14, 1u) == gridDim.x - 1u) {
                counter[14] = 0u;
                const unsigned int seq = counter[15] + 1u;
                counter[15] = seq;
                // ONE system-scope fence, by the CTA that has observed every other CTA's count: cumulativity orders all their
                // stores before the completion word for an observer on the host (320 system fences cost ~4 us per step)
                __threadfence_system();
                *reinterpret_cast<volatile unsigned int*>(p) = seq;
            }
        }
    }
}

// mean of rows[0..n) in a fixed order, by one block
__device__ __forceinline__ float block_mean_fixed(const float* rows, int n, float* red)
{
    float a = 0.f;
    for (int i = threadIdx.x; i < n; i += FIN_THREADS) a += __ldcg(rows + i);
    return block_sum<FIN_THREADS>(a, red) / (float)n;
}

// Optional fused tail of the step: the ring-buffer enqueue (mem_moco.py:17-27) rides in extra CTAs of the finalize
// launch -- legal because the launch is stream-ordered after the queue-streaming kernel, the only reader of the queue.
template <typename QT>
__device__ __forceinline__ void enqueue_rows(const FinalizeParams& F, int blk, int nblk)
{
    QT* queue = reinterpret_cast<QT*>(F.enq_queue);
    long long index = F.enq_state ? *reinterpret_cast<volatile long long*>(F.enq_state) : F.enq_index;
    const int d4 = F.d / 4;
    const long long total = (long long)F.enq_N * d4;
    const float4* keys = reinterpret_cast<const float4*>(F.enq_keys);
    const bool peer = F.xchg.mailboxes != nullptr;
    unsigned long long xstep = 0;
    bool timed_out = false;
    if (peer) {
        // the gathered keys are this rank's mailbox slots of the current parity: [W][B*d] == all_k, pushed by the peers'
        // prep kernels while the queue was being swept.  Wait for every (rank, slice) flag, then read around L1.
        xstep = *reinterpret_cast<volatile unsigned long long*>(F.xchg.xstate);
        int ok = 1;
        if ((int)threadIdx.x < F.xchg.W * XCHG_SLICES)
            ok = xchg_wait_slice(F.xchg, xstep, threadIdx.x / XCHG_SLICES, threadIdx.x % XCHG_SLICES) ? 1 : 0;
        if (F.timebuf && threadIdx.x == 0) atomicMax(F.timebuf + 32 * 1000 + 6, globaltimer_ns());   // bring-up only: flags seen
        // a peer that did not deliver within the timeout (xstate[2] is raised, sticky): this CTA writes nothing -- the mailbox
        // holds stale rows -- and the last CTA below leaves the ring pointer where it was.  The step's loss and gradient do not
        // depend on the gathered keys; the caller polls the flag (PeerKeyExchange.check(), GraphedReplicaStep.step()).
        timed_out = (__syncthreads_and(ok) == 0);
        keys = xchg_slot(F.xchg.mailboxes[F.xchg.rank], F.xchg, (int)(xstep & 1ull), 0);
    }
    // four rows' worth of loads in flight per thread before the first store (W ranks' keys: W * B * d / 4 items over <= 64 CTAs)
    const long long step_i = (long long)nblk * FIN_THREADS;
    for (long long i0 = (long long)blk * FIN_THREADS + threadIdx.x; !timed_out && i0 < total; i0 += 4 * step_i) {
        float4 v[4];
        long long off[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long i = i0 + u * step_i;
            off[u] = -1;
            if (i < total) {
                const int row = (int)(i / d4), c4 = (int)(i - (long long)row * d4);
                long long slot = index + row;
                if (slot >= F.enq_K) slot -= F.enq_K;
                bool mine = true;
                if (F.enq_kend > 0) {                         // K-sharded queue: this rank stores only the slots it owns
                    mine = (slot >= F.enq_kbegin && slot < F.enq_kend);
                    slot -= F.enq_kbegin;
                }
                if (mine) {
                    v[u] = peer ? ld_cg_f4(keys + i) : __ldg(keys + i);
                    off[u] = slot * d4 + c4;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (off[u] < 0) continue;
            if constexpr (sizeof(QT) == 4) {
                reinterpret_cast<float4*>(queue)[off[u]] = v[u];
            } else {
                __nv_bfloat162 lo = __floats2bfloat162_rn(v[u].x, v[u].y), hi = __floats2bfloat162_rn(v[u].z, v[u].w);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&lo);
                pk.y = *reinterpret_cast<uint32_t*>(&hi);
                reinterpret_cast<uint2*>(queue)[off[u]] = pk;
            }
        }
    }
    if (F.enq_state) {                                        // device-resident pointer: the last enqueue CTA advances it
        __shared__ int last;
        __syncthreads();
        if (threadIdx.x == 0)
            last = (atomicAdd(reinterpret_cast<unsigned long long*>(F.enq_state + 1), 1ull) == (unsigned long long)(nblk - 1));
        __syncthreads();
        if (last && threadIdx.x == 0) {
            long long nx = index + F.enq_N;
            if (nx >= F.enq_K) nx -= F.enq_K;
            // (every enqueue CTA raised the flag before taking its ticket if its wait expired)
            const bool failed = peer && *reinterpret_cast<volatile unsigned long long*>(F.xchg.xstate + 2) != 0ull;
            if (!failed) F.enq_state[0] = nx;
            F.enq_state[1] = 0;
            if (peer) F.xchg.xstate[0] = xstep + 1;       // every enqueue CTA has read the step before taking its ticket
        }
    }
    if (F.timebuf && threadIdx.x == 0) atomicMax(F.timebuf + 32 * 1000 + 5, globaltimer_ns());       // bring-up only
}

// kVec (d == 128, <= 8 * FIN_VCH splits: the tcgen05 family at its usual sizes): every warp reads whole 512-byte partial rows
// with one 128-bit load per lane (warp w takes splits w, w + 8, ...), a quarter of the load instructions and L1 requests of
// the column-per-thread path; the eight per-warp sums of a column are then added in warp order (fixed: deterministic).
template <int kMode, bool kVec>
__global__ void __launch_bounds__(FIN_THREADS)
infonce_finalize_kernel(const FinalizeParams F)
{
    __shared__ float4 colsum4[kVec ? (FIN_THREADS / 32) * 32 : 1];
    __shared__ float w_s[kVec ? 8 * FIN_VCH : FIN_MAX_SPLITS];
    __shared__ float red[FIN_THREADS / 32];
    __shared__ float colsum[FIN_GROUPS][FIN_COLS];
    __shared__ float row_stat[2];                              // lse, pos of this row
    __shared__ int   flag;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // launched with a programmatic dependency on the stream kernel: everything below reads its partials (or, for the
    // enqueue CTAs, overwrites queue rows it may still be reading)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // the next launch in the stream (the prep kernel of the next step waits for this whole grid before it touches anything)
    // may become resident now
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (F.timebuf && tid == 0) {                              // bring-up only
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMin(F.timebuf + 32 * 1000 + 0, t);
        F.timebuf[32 * 400 + 4 * (blockIdx.x & 255)] = t;
    }
    if (kMode == FIN_FULL && (int)blockIdx.x >= F.B) {        // fused enqueue CTAs
        if (F.enq_dtype == GCA_F32) enqueue_rows<float>(F, blockIdx.x - F.B, gridDim.x - F.B);
        else                        enqueue_rows<__nv_bfloat16>(F, blockIdx.x - F.B, gridDim.x - F.B);
        publish_done(F.counter);
        return;
    }
    const int b = blockIdx.x;
    int ns = F.nsplit;
    const int grp = tid / FIN_COLS, col = tid % FIN_COLS;
    size_t stride = (size_t)F.Bpad * F.d;                      // floats between two splits of the accumulator partials
    const float* p_max = F.part_max; const float* p_sum = F.part_sum; const int* p_cnt = F.part_cnt; const float* p_acc = F.part_acc;
    float* out_row = (kMode == FIN_SHARD) ? (F.out_acc ? F.out_acc + (size_t)b * F.d : nullptr) : (F.dq ? F.dq + (size_t)b * F.d : nullptr);
    float* o_max = F.out_max ? F.out_max + b : nullptr; float* o_sum = F.out_sum ? F.out_sum + b : nullptr;
    int* o_cnt = F.out_cnt ? F.out_cnt + b : nullptr;
    // K-sharded step over peer memory (PeerMerge, gca_common.cuh)
    const bool peer_push = (kMode == FIN_SHARD) && F.merge.mailboxes != nullptr;
    const bool peer_merge = (kMode == FIN_FULL) && F.merge.mailboxes != nullptr && F.merge.wait != 0;
    unsigned long long mstep = 0ull;
    int mpar = 0;
    int brow = b;                                              // row index inside a split's partial arrays
    if (peer_push || peer_merge) {
        const PeerMerge& M = F.merge;
        mstep = *reinterpret_cast<volatile unsigned long long*>(M.mstate);
        mpar = (int)(mstep & 1ull);
        const size_t Bg = (size_t)M.W * M.Bl;
        if (peer_push) {
            // the merged partial of global row b (relative to its max) goes into this rank's own rows block
            float* base = pm_rows(M.mailboxes[M.rank], M, mpar);
            out_row = F.part_acc ? base + (size_t)b * M.d : nullptr;
            o_max = base + Bg * M.d + b; o_sum = o_max + Bg; o_cnt = reinterpret_cast<int*>(o_sum + Bg);
        } else {
            // wait until every rank has published its rows block of this step, then pull my row from each of them
            int ok = 1;
            if (tid < M.W) ok = pm_wait(M, mpar, tid, mstep) ? 1 : 0;
            (void)__syncthreads_and(ok);                       // on a timeout mstate[2] is raised (sticky) and the row is garbage
            brow = M.rank * M.Bl + b;
            ns = M.W;
        }
    }
    // base pointers of split sp: accumulator rows, max / sum / count arrays
    auto acc_of = [&](int sp) -> const float* {
        return peer_merge ? pm_rows(F.merge.mailboxes[sp], F.merge, mpar) : p_acc + (size_t)sp * stride;
    };
    // statistics of split sp for this row: part_* [Bpad, nsplit] (gca_common.cuh), or -- cross-rank merge -- the max / sum / count
    // arrays behind rank sp's rows block
    auto max_of = [&](int sp) -> const float* {
        return peer_merge ? pm_rows(F.merge.mailboxes[sp], F.merge, mpar) + (size_t)F.merge.W * F.merge.Bl * F.merge.d + brow
                          : p_max + part_stat_index(sp, brow, F.nsplit);
    };
    const size_t stat_gap = peer_merge ? (size_t)F.merge.W * F.merge.Bl : 0;     // merge: max -> sum -> count arrays of a block
    if (peer_merge) p_acc = F.dq ? reinterpret_cast<const float*>(F.merge.mailboxes) : nullptr;      // (non-null marker only)
    const bool want_acc = (out_row != nullptr && p_acc != nullptr);

    // Loads first, arithmetic later: warp 0 puts the split statistics of its row in flight (they head the memory queue),
    // then every thread puts its share of the gradient partials in flight (raw values: the split weights are not needed to
    // LOAD them) -- one L2 round trip for the whole block instead of a chain of dependent ones.
    // control word 6: a stream / prep kernel of the tcgen05 family saw a logit outside the unit-row range (uniform for
    // the whole launch: written before this kernel started, cleared by the last ticket holder below)
    const unsigned int range_flag = (kMode == FIN_FULL && F.pk_nb > 0) ? __ldcg(F.counter + 6) : 0u;
    float st_m[FIN_STAT], st_s[FIN_STAT];
    int st_c[FIN_STAT];
    const bool stat_fast = (ns <= 32 * FIN_STAT);
    if (warp == 0 && kMode != FIN_BWD && stat_fast) {
#pragma unroll
        for (int i = 0; i < FIN_STAT; ++i) {
            const int sp = lane + 32 * i;
            const int spc = sp < ns ? sp : 0;
            const float* pm_ = max_of(spc);
            const float* ps_ = peer_merge ? pm_ + stat_gap : p_sum + part_stat_index(spc, brow, F.nsplit);
            const int* pc_ = peer_merge ? reinterpret_cast<const int*>(ps_ + stat_gap) : p_cnt + part_stat_index(spc, brow, F.nsplit);
            st_m[i] = (sp < ns) ? ld_part(pm_, peer_merge) : -INFINITY;
            st_s[i] = (sp < ns) ? ld_part(ps_, peer_merge) : 0.f;
            st_c[i] = (sp < ns) ? ld_part(pc_, peer_merge) : 0;
        }
    }
    float v[kVec ? 1 : FIN_CHUNK];
    float4 v4[kVec ? FIN_VCH : 1];
    if constexpr (kVec) {
        if (want_acc) {
#pragma unroll
            for (int i = 0; i < FIN_VCH; ++i) {
                const int sp = warp + i * (FIN_THREADS / 32);
                v4[i] = (sp < ns) ? ld_part(reinterpret_cast<const float4*>(acc_of(sp) + (size_t)brow * FIN_COLS) + lane, peer_merge)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    } else if (want_acc && col < F.d) {
#pragma unroll
        for (int i = 0; i < FIN_CHUNK; ++i) {
            const int sp = grp + i * FIN_GROUPS;
            v[i] = (sp < ns) ? ld_part(acc_of(sp) + (size_t)brow * F.d + col, peer_merge) : 0.f;
        }
    }

    int row_rank = 0;                                          // thread 0: #negatives above the positive (FIN_FULL)
    if (warp == 0) {
        float lse = 0.f, pos = 0.f;
        if (kMode == FIN_BWD) {
            lse = F.lse_in[b];
            pos = F.pos ? F.pos[b] : 0.f;
        } else {
            // M = max over splits (and the positive), S = sum of rescaled split sums: fixed order for a given split count
            if (kMode == FIN_FULL) pos = F.pos[b];
            float m = -INFINITY, part = 0.f;
            int cnt = 0;
            if (stat_fast) {
#pragma unroll
                for (int i = 0; i < FIN_STAT; ++i) m = fmaxf(m, st_m[i]);
                m = warp_max(m);
                if (kMode == FIN_FULL) m = fmaxf(m, pos);
#pragma unroll
                for (int i = 0; i < FIN_STAT; ++i) {
                    const int sp = lane + 32 * i;
                    const float e = (st_m[i] == -INFINITY) ? 0.f : __expf(st_m[i] - m);
                    if (sp < ns) w_s[sp] = e;                  // exp(part_max - M); turned into exp(. - lse) below
                    part += st_s[i] * e;
                    cnt += st_c[i];
                }
            } else {                                           // many splits: two dependent passes over the statistics
                for (int sp = lane; sp < ns; sp += 32) m = fmaxf(m, ld_part(max_of(sp), peer_merge));
                m = warp_max(m);
                if (kMode == FIN_FULL) m = fmaxf(m, pos);
                for (int sp = lane; sp < ns; sp += 32) {
                    const float* pm_ = max_of(sp);
                    const float* ps_ = peer_merge ? pm_ + stat_gap : p_sum + part_stat_index(sp, brow, F.nsplit);
                    const int* pc_ = peer_merge ? reinterpret_cast<const int*>(ps_ + stat_gap) : p_cnt + part_stat_index(sp, brow, F.nsplit);
                    const float pm = ld_part(pm_, peer_merge);
                    const float e = (pm == -INFINITY) ? 0.f : __expf(pm - m);
                    w_s[sp] = e;
                    part += ld_part(ps_, peer_merge) * e;
                    cnt += ld_part(pc_, peer_merge);
                }
            }
            float S = warp_sum(part);
            cnt = warp_sum_i(cnt);
            if (kMode == FIN_FULL) {
                S += __expf(pos - m);
                lse = m + logf(S);
                if (lane == 0) {
                    F.lse[b] = lse;
                    F.loss_rows[b] = lse - pos;
                    if (F.rank_gt) F.rank_gt[b] = cnt;
                    row_rank = cnt;
                }
                const float corr = __expf(m - lse);            // = 1 / S
                for (int sp = lane; sp < ns; sp += 32) w_s[sp] *= corr;
            } else if (lane == 0) {                            // FIN_SHARD: keep the merged partial relative to m
                *o_max = m; *o_sum = S; *o_cnt = cnt;
            }
        }
        if (lane == 0) { row_stat[0] = lse; row_stat[1] = pos; }
    }
    __syncthreads();
    // Mean loss and top-k hit counts without a second pass over the rows.  Packed form (the usual case): ONE 64-bit
    // atomicAdd per row carries [ticket | top-1 bit | top-5 bit | loss as a fixed-point integer]; integer addition is
    // associative, so the total is exact and order-independent (deterministic), and the value the atomic RETURNS to the last
    // ticket holder plus its own contribution is the complete sum -- no fence, no read-back.  It is issued here, before the
    // gradient accumulation, and only looked at afterwards, so its L2 round trip is off the critical path.
    unsigned long long pk_mine = 0ull, pk_old = 0ull;
    const bool pk_on = (kMode == FIN_FULL) && F.pk_nb > 0 && (F.loss_mean != nullptr || F.top_hits != nullptr) && tid == 0 &&
                       range_flag == 0u;
    if (pk_on) {
        const float lrow = fmaxf(row_stat[0] - row_stat[1], 0.f);          // lse - pos of this row: >= 0 up to rounding
        const int nb = F.pk_nb;
        pk_mine = 1ull | ((row_rank < 1 ? 1ull : 0ull) << nb) | ((row_rank < 5 ? 1ull : 0ull) << (2 * nb)) |
                  ((unsigned long long)__double2ll_rn((double)lrow * (double)(1ull << F.pk_frac)) << (3 * nb));
        pk_old = atomicAdd(reinterpret_cast<unsigned long long*>(F.counter + 2), pk_mine);
    }
    if (F.timebuf && tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); F.timebuf[32 * 400 + 4 * (blockIdx.x & 255) + 1] = t; }

    // Gradient accumulator: each of the FIN_GROUPS thread groups sums its interleaved splits in order, then the groups are
    // added in order: deterministic for a given split count.
    if (want_acc) {
        const float lse = row_stat[0], pos = row_stat[1];
        const float p0m1 = (kMode == FIN_SHARD) ? 0.f : (__expf(pos - lse) - 1.f);
        const float scale = (kMode == FIN_FULL) ? F.inv_T / (float)F.B
                          : (kMode == FIN_BWD)  ? F.inv_T * F.grad_scale : 1.f;
        for (int c0 = 0; c0 < F.d; c0 += FIN_COLS) {
            const int c = c0 + col;
            float t = 0.f;
            if constexpr (kVec) {
                float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int i = 0; i < FIN_VCH; ++i) {
                    const int sp = warp + i * (FIN_THREADS / 32);
                    if (sp < ns) {
                        const float w = (kMode == FIN_BWD) ? 1.f : w_s[sp];
                        a4.x = fmaf(w, v4[i].x, a4.x); a4.y = fmaf(w, v4[i].y, a4.y);
                        a4.z = fmaf(w, v4[i].z, a4.z); a4.w = fmaf(w, v4[i].w, a4.w);
                    }
                }
                colsum4[warp * 32 + lane] = a4;
                __syncthreads();
                if (grp == 0) {
                    const float* cs = reinterpret_cast<const float*>(colsum4);
                    t = cs[col];
#pragma unroll
                    for (int w = 1; w < FIN_THREADS / 32; ++w) t += cs[w * FIN_COLS + col];
                }
            } else {
            float a = 0.f;
            if (c < F.d) {
                for (int base = 0; base < ns; base += FIN_CHUNK * FIN_GROUPS) {
                    if (c0 > 0 || base > 0) {                  // beyond what was preloaded above
#pragma unroll
                        for (int i = 0; i < FIN_CHUNK; ++i) {
                            const int sp = base + grp + i * FIN_GROUPS;
                            v[i] = (sp < ns) ? ld_part(acc_of(sp) + (size_t)brow * F.d + c, peer_merge) : 0.f;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < FIN_CHUNK; ++i) {
                        const int sp = base + grp + i * FIN_GROUPS;
                        if (sp < ns) a = (kMode == FIN_BWD) ? a + v[i] : fmaf(w_s[sp], v[i], a);
                    }
                }
            }
            colsum[grp][col] = a;
            __syncthreads();
            if (grp == 0 && c < F.d) {
                t = colsum[0][col];
#pragma unroll
                for (int g = 1; g < FIN_GROUPS; ++g) t += colsum[g][col];
            }
            }
            if (grp == 0 && c < F.d) {
                if (kMode != FIN_SHARD) t = scale * fmaf(p0m1, F.k[(size_t)b * F.d + c], t);
            }
            if (kMode == FIN_FULL && F.zq != nullptr) {
                // q = zq / ||zq||  =>  d loss / d zq = (g - (g . q) q) / ||zq||   (g = d loss / d q, this row; d == FIN_COLS)
                const float inv = F.inv_nq[b];
                const float qh = (grp == 0 && c < F.d) ? F.zq[(size_t)b * F.d + c] * inv : 0.f;
                __syncthreads();                                       // colsum is re-used by nobody below; red is free
                const float dot = block_sum<FIN_THREADS>(t * qh, red);
                t = (t - dot * qh) * inv;
            }
            if (grp == 0 && c < F.d) out_row[c] = t;
            __syncthreads();
        }
    }
    if (peer_push) {
        // last CTA of the launch (ticket): every row of this rank's block is written -> publish the step to every peer with one
        // system-scope release store each (the fence + ticket orders the other CTAs' rows before it: cumulativity)
        __shared__ int last_cta;
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            last_cta = (atomicAdd(F.merge.mstate + 1, 1ull) == (unsigned long long)gridDim.x - 1ull);
        }
        __syncthreads();
        if (last_cta) {
            if (tid == 0) { F.merge.mstate[1] = 0ull; __threadfence(); }
            __syncthreads();
            if (tid < F.merge.W) st_release_sys_u64(pm_flag(F.merge.mailboxes[tid], F.merge, mpar, F.merge.rank), mstep + 1ull);
            if (F.timebuf && tid == 0) atomicMax(F.timebuf + 32 * 1000 + 3, globaltimer_ns());      // bring-up only
        }
        return;
    }

    if (F.timebuf && tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); F.timebuf[32 * 400 + 4 * (blockIdx.x & 255) + 2] = t; atomicMax(F.timebuf + 32 * 1000 + 3, t); }
    // counter layout (gca_common.cuh): [0] ticket, [2..3] loss accumulator (or the packed word), [4] top-1, [5] top-5.
    if (pk_on) {
        const int nb = F.pk_nb;
        const unsigned long long mask = (1ull << nb) - 1ull;
        if ((pk_old & mask) == (unsigned long long)(F.B - 1)) {              // every other row is already in pk_old
            const unsigned long long tot = pk_old + pk_mine;
            if (F.loss_mean) *F.loss_mean = (float)((double)(tot >> (3 * nb)) / (double)(1ull << F.pk_frac) / (double)F.B);
            if (F.top_hits) { F.top_hits[0] = (int)((tot >> nb) & mask); F.top_hits[1] = (int)((tot >> (2 * nb)) & mask); }
            *reinterpret_cast<unsigned long long*>(F.counter + 2) = 0ull;   // re-arm for the next launch on this workspace
            if (peer_merge) {                                               // every row CTA read the step before its ticket
                F.merge.mstate[0] = mstep + 1ull;
                if (F.merge.gather_state) F.merge.gather_state[0] = mstep + 1ull;
            }
            if (F.timebuf) {
                unsigned long long tt;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
                F.timebuf[32 * 1000 + 1] = tt;
            }
        }
    } else if (kMode == FIN_FULL && (F.loss_mean != nullptr || F.top_hits != nullptr) && tid == 0) {
        // unpacked form (very large B or temperatures so small that the fixed-point fields do not fit 64 bits)
        const float lrow = row_stat[0] - row_stat[1];                      // lse - pos of this row
        unsigned long long* acc64 = reinterpret_cast<unsigned long long*>(F.counter + 2);
        atomicAdd(acc64, (unsigned long long)__double2ll_rn((double)lrow * 68719476736.0));   // exact: ulp(lrow) >= 2^-36
        const int r = row_rank;
        if (r < 1) atomicAdd(F.counter + 4, 1u);
        if (r < 5) atomicAdd(F.counter + 5, 1u);
        __threadfence();
        const unsigned int t = atomicAdd(F.counter, 1u);
        if (t == (unsigned int)F.B - 1) {
            __threadfence();
            const long long tot = (long long)atomicAdd(acc64, 0ull);
            if (F.loss_mean) *F.loss_mean = (float)((double)tot / 68719476736.0 / (double)F.B);
            if (F.top_hits) { F.top_hits[0] = (int)atomicAdd(F.counter + 4, 0u); F.top_hits[1] = (int)atomicAdd(F.counter + 5, 0u); }
            F.counter[0] = 0u; F.counter[2] = 0u; F.counter[3] = 0u; F.counter[4] = 0u; F.counter[5] = 0u;
            F.counter[6] = 0u;
            if (peer_merge) {
                F.merge.mstate[0] = mstep + 1ull;
                if (F.merge.gather_state) F.merge.gather_state[0] = mstep + 1ull;
            }
            if (F.timebuf) {
                unsigned long long tt;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
                F.timebuf[32 * 1000 + 1] = tt;
            }
        }
    }
    if (kMode == FIN_FULL) publish_done(F.counter);
}

int infonce_finalize_launch(const FinalizeParams& F_, int mode, cudaStream_t st)
{
    FinalizeParams F = F_;
    F.timebuf = debug_timebuf();
    if (F.timebuf && mode == FIN_FULL && F.merge.mailboxes && F.merge.wait) F.timebuf += 8;   // the merge launch stamps its own words
    if (F.nsplit > FIN_MAX_SPLITS) return set_err(GCA_ERR_UNSUPPORTED, "finalize: %d splits > %d", F.nsplit, FIN_MAX_SPLITS);
    F.pk_nb = 0; F.pk_frac = 0;
    if (mode == FIN_FULL && F.range_checked) {
        // packed loss/hit/ticket word: nb bits each for the ticket and the two hit counts, the rest for the loss sum as a
        // fixed-point integer.  A row loss is at most 2 L + ln(K+1) when every |logit| <= L; the stream kernels of the
        // tcgen05 family raise control word 6 when a logit leaves L = 1.0625 / T (unit rows stay within 1 / T), and the
        // kernel then takes the unpacked path.
        int nb = 1;
        while ((1ll << nb) <= (long long)F.B) ++nb;
        const double lmax = (2.125 * (double)F.inv_T + 45.0) * (double)F.B;
        int ib = 1;
        while ((double)(1ull << ib) <= lmax && ib < 40) ++ib;
        const int frac = 64 - 3 * nb - ib - 1;
        if (frac >= 20) { F.pk_nb = nb; F.pk_frac = frac > 36 ? 36 : frac; }
    }
    int enq_blocks = 0;
    if (F.xchg.mailboxes && F.xchg.W * XCHG_SLICES > FIN_THREADS)
        return set_err(GCA_ERR_UNSUPPORTED, "finalize: peer exchange over %d ranks", F.xchg.W);
    if (mode == FIN_FULL && F.enq_queue != nullptr && F.enq_N > 0) {
        enq_blocks = (int)(((long long)F.enq_N * (F.d / 4) + FIN_THREADS - 1) / FIN_THREADS);
        if (enq_blocks > 64) enq_blocks = 64;
    }
    cudaLaunchConfig_t cfg{};
    static int vec_on = -1;                                   // GCA_FIN_NOVEC=1: column-per-thread loads everywhere (A/B timing)
    if (vec_on < 0) { const char* e = getenv("GCA_FIN_NOVEC"); vec_on = (e && e[0] == '1') ? 0 : 1; }
    const bool peer_merge = mode == FIN_FULL && F.merge.mailboxes != nullptr && F.merge.wait != 0;
    const bool peer_push = mode == FIN_SHARD && F.merge.mailboxes != nullptr;
    if ((peer_merge || peer_push) && (F.merge.W > FIN_THREADS || F.merge.d != F.d))
        return set_err(GCA_ERR_UNSUPPORTED, "finalize: peer merge over %d ranks / d = %d", F.merge.W, F.merge.d);
    if (peer_merge && F.loss_mean == nullptr && F.top_hits == nullptr)
        return set_err(GCA_ERR_BAD_ARG, "finalize: the peer merge needs loss_mean (its ticket advances the step counter)");
    const int ns_eff = peer_merge ? F.merge.W : F.nsplit;
    const bool vec = vec_on && F.d == FIN_COLS && ns_eff <= 8 * FIN_VCH &&
                     (peer_merge ? F.dq != nullptr
                                 : F.part_acc != nullptr && ((mode == FIN_SHARD) ? (F.out_acc != nullptr || peer_push) : F.dq != nullptr));
    cfg.gridDim = dim3(F.B + enq_blocks); cfg.blockDim = dim3(FIN_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;    // PDL: be resident when the stream kernel drains
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
#define GCA_FIN_LAUNCH(MODE) do { \
        if (vec) GCA_CUDA(launch_ex(&cfg, infonce_finalize_kernel<MODE, true>, F)); \
        else     GCA_CUDA(launch_ex(&cfg, infonce_finalize_kernel<MODE, false>, F)); } while (0)
    if (mode == FIN_FULL)       GCA_FIN_LAUNCH(FIN_FULL);
    else if (mode == FIN_SHARD) GCA_FIN_LAUNCH(FIN_SHARD);
    else                        GCA_FIN_LAUNCH(FIN_BWD);
#undef GCA_FIN_LAUNCH
    GCA_LAUNCH_CHECK("infonce_finalize_kernel");
    count_launch(1);
    return GCA_OK;
}

// ---------------------------------------------------------------------------------------------
// K-shard combine: merge the W per-rank partials (+ the positive) and rescale this rank's accumulator
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FIN_THREADS)
shard_combine_kernel(const float* __restrict__ all_max, const float* __restrict__ all_sum,
                     const int* __restrict__ all_cnt, int W, int rank_id, int B, int d,
                     const float* __restrict__ pos_logit, float* lse_out, float* loss_rows, int* rank_gt,
                     float* part_acc)
{
    const int b = blockIdx.x, tid = threadIdx.x;
    const float pos = pos_logit[b];
    float m = pos;
    for (int r = 0; r < W; ++r) m = fmaxf(m, all_max[(size_t)r * B + b]);
    float S = __expf(pos - m);
    int cnt = 0;
    for (int r = 0; r < W; ++r) {                            // fixed rank order on every rank -> identical lse everywhere
        const float pm = all_max[(size_t)r * B + b];
        S += (pm == -INFINITY) ? 0.f : all_sum[(size_t)r * B + b] * __expf(pm - m);
        cnt += all_cnt[(size_t)r * B + b];
    }
    const float lse = m + logf(S);
    if (tid == 0) {
        lse_out[b] = lse;
        if (loss_rows) loss_rows[b] = lse - pos;
        if (rank_gt) rank_gt[b] = cnt;
    }
    if (part_acc) {
        const float pm = all_max[(size_t)rank_id * B + b];
        const float w = (pm == -INFINITY) ? 0.f : __expf(pm - lse);
        for (int c = tid; c < d; c += FIN_THREADS) part_acc[(size_t)b * d + c] *= w;
    }
}

__global__ void __launch_bounds__(FIN_THREADS)
shard_finish_kernel(const float* __restrict__ acc, const float* __restrict__ k, const float* __restrict__ pos_logit,
                    const float* __restrict__ lse, const float* loss_rows, int B_loc, int d, float inv_T,
                    float* dq_unit, float* loss_mean, unsigned int* counter_unused)
{
    __shared__ float red[FIN_THREADS / 32];
    const int b = blockIdx.x, tid = threadIdx.x;
    if (b < B_loc && dq_unit) {
        const float p0m1 = __expf(pos_logit[b] - lse[b]) - 1.f;
        const float scale = inv_T / (float)B_loc;
        for (int c = tid; c < d; c += FIN_THREADS)
            dq_unit[(size_t)b * d + c] = scale * fmaf(p0m1, k[(size_t)b * d + c], acc[(size_t)b * d + c]);
    }
    if (b == B_loc && loss_mean) {                            // one extra block: the mean over the local rows
        const float mean = block_mean_fixed(loss_rows, B_loc, red);
        if (tid == 0) *loss_mean = mean;
    }
}

}  // namespace gca

extern "C" int gca_infonce_shard_combine(const float* all_max, const float* all_sum, const int* all_cnt, int W,
                                         int rank_id, int B, int d, const float* pos_logit, float* lse,
                                         float* loss_rows, int* rank_gt, float* part_acc, void* stream)
{
    using namespace gca;
    GCA_CHECK_ARG(all_max && all_sum && all_cnt && pos_logit && lse, "gca_infonce_shard_combine: null pointer");
    GCA_CHECK_ARG(W >= 1 && rank_id >= 0 && rank_id < W && B >= 1 && d >= 1, "gca_infonce_shard_combine: bad sizes");
    shard_combine_kernel<<<B, FIN_THREADS, 0, (cudaStream_t)stream>>>(all_max, all_sum, all_cnt, W, rank_id, B, d,
                                                                     pos_logit, lse, loss_rows, rank_gt, part_acc);
    GCA_LAUNCH_CHECK("shard_combine_kernel");
    count_launch(1);
    return GCA_OK;
}

extern "C" int gca_infonce_shard_finish(const float* acc, const float* k, const float* pos_logit, const float* lse,
                                        const float* loss_rows, int B_loc, int d, float inv_T, float* dq_unit,
                                        float* loss_mean, void* stream)
{
    using namespace gca;
    GCA_CHECK_ARG(pos_logit && lse, "gca_infonce_shard_finish: null pointer");
    GCA_CHECK_ARG(!dq_unit || (acc && k), "gca_infonce_shard_finish: dq_unit needs acc and k");
    GCA_CHECK_ARG(!loss_mean || loss_rows, "gca_infonce_shard_finish: loss_mean needs loss_rows");
    GCA_CHECK_ARG(B_loc >= 1 && d >= 1, "gca_infonce_shard_finish: bad sizes");
    shard_finish_kernel<<<B_loc + 1, FIN_THREADS, 0, (cudaStream_t)stream>>>(acc, k, pos_logit, lse, loss_rows, B_loc,
                                                                            d, inv_T, dq_unit, loss_mean, nullptr);
    GCA_LAUNCH_CHECK("shard_finish_kernel");
    count_launch(1);
    return GCA_OK;
}
