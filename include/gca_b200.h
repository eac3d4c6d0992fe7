/* gca_b200.h -- C ABI of libgca_b200.so, the B200 (sm_100a) implementation of GCA's contrastive hot path.
 *
 * The reference (ACMMM2021-Anonymous/video-graph-ssl) is pure Python/PyTorch and has no FFI of its own;
 * the boundary it offers is a set of nn.Module call conventions.  Each entry point
 * below states the reference code it replaces (paths relative to the reference repo).  The Python mirror of
 * those module interfaces lives in video-graph-ssl_b200/gca_b200/ and binds this header with ctypes
 * (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller, contiguous, 16-byte aligned; row-major.
 *   - the library allocates no persistent device memory, frees nothing, never synchronises the device;
 *     scratch space is caller-provided (`gca_*_workspace_bytes`).  Every call only enqueues work on
 *     `stream` (a cudaStream_t passed as void*) and returns.
 *   - return value: GCA_OK or a negative GCA_ERR_*; `gca_last_error()` gives a thread-local message.
 *     No exceptions, no abort, and NO CPU FALLBACK: without a CUDA device every compute call fails.
 *   - thread-safety: re-entrant; the only global state is a per-thread cache of TMA descriptors.
 *   - `dtype_queue`: GCA_F32 = queue rows stored fp32 (parity mode), GCA_BF16 = stored bf16 (fast mode).
 *   - temperature is passed as inv_T = 1/T.
 */
#ifndef GCA_B200_H
#define GCA_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCA_ABI_VERSION 1

#define GCA_OK               0
#define GCA_ERR_BAD_ARG     (-1)
#define GCA_ERR_UNSUPPORTED (-2)
#define GCA_ERR_CUDA        (-3)
#define GCA_ERR_WORKSPACE   (-4)

#define GCA_F32  0
#define GCA_BF16 1

/* kernel family for the InfoNCE stream */
#define GCA_ALGO_AUTO    0   /* d == 128: tcgen05 (bf16 queue) or tc32 (fp32 queue); other widths, materialised logits, two-pass backward: ffma */
#define GCA_ALGO_FFMA    1   /* CUDA-core fp32 FMA, exact fp32 arithmetic (parity mode), any d % 32 == 0, d <= 1024 */
#define GCA_ALGO_TCGEN05 2   /* TMA -> smem -> tcgen05.mma (bf16 x bf16 -> fp32 in TMEM); bf16 queue, d == 128 */
#define GCA_ALGO_TC32    3   /* fp32 queue, d == 128, fp32-grade logits on tcgen05: exact 3-way bf16 split of q and the queue, six
                                piece products per logit in one fp32 TMEM accumulator (loss within 1e-5 of the reference's fp32
                                arithmetic); the gradient accumulates bf16 pieces (<= 1e-2).  gca_infonce_fwd / gca_moco_step only;
                                the workspace grows by 6 * K * d bytes (gca_infonce_workspace_bytes with this algo) */

/* ranks are exact below this value when the caller asks for top-k hit counts only (rank_gt == NULL) */
#define GCA_TOPK_RANK_CAP 8

/* graph head flags (0 = the reference's arithmetic).  The variants below have NO counterpart in the reference (SURVEY.md
 * D5-D7: its only sparsifier is the hop mask, its only augmentation the relaxed Bernoulli, its only normalisation the row
 * softmax); they are default-OFF options whose parity is UNPINNED -- the checker is our own restatement (oracle/graph.py,
 * graph_variants).  They apply to the hop-weighted adjacency `adj`, in this order:
 *   THRESHOLD : adj_ij < tau -> 0                                   (removed like a hop-masked edge)
 *   TOPK      : keep the topk largest entries of every row (ties towards the lower column), the rest -> 0
 *   EDGE_DROP : hard seeded edge drop instead of the relaxed Bernoulli: s_ij = adj_ij * [u_ij >= p_drop]
 *   SYMNORM   : s <- D^-1/2 s D^-1/2, D = diag(row sums of s)       (symmetric normalisation before the aggregation)
 * Cosine adjacency and the feature mask are compositions around the kernel (gca_b200.ops.TemporalGraphAug). */
#define GCA_GRAPH_REFERENCE 0u
#define GCA_GRAPH_THRESHOLD 1u
#define GCA_GRAPH_TOPK      2u
#define GCA_GRAPH_EDGE_DROP 4u
#define GCA_GRAPH_SYMNORM   8u
typedef struct GcaGraphOpts { unsigned flags; float tau; int topk; float p_drop; } GcaGraphOpts;

int         gca_version(void);
const char* gca_last_error(void);
/* number of SMs of the current device, or a negative error */
int         gca_sm_count(void);
/* number of kernels this library has launched in this process so far (bench.py's gpu_launches bookkeeping) */
long long   gca_launch_count(void);

/* ---------------------------------------------------------------------------------------------------------
 * MoCo queue: in-place ring-buffer enqueue.
 * Replaces BaseMoCo._update_memory (lib/memory/mem_moco.py:17-27): ids = (arange(N) + index) mod K,
 * queue.index_copy_(0, ids, keys).  The caller keeps the pointer and advances it itself
 * ((index + N) % K, mem_moco.py:14-15).
 *   queue    : [k_end - k_begin, d] rows of the shard holding GLOBAL slots [k_begin, k_end) of a ring of
 *              K_global slots (single GPU: k_begin = 0, k_end = K_global)
 *   keys     : [N, d] fp32; stored as-is (GCA_F32) or rounded to nearest-even bf16 (GCA_BF16)
 *   index    : global pointer BEFORE this enqueue.  Requires N <= K_global (the reference's index_copy_
 *              with duplicate indices is undefined); d % 4 == 0.
 * --------------------------------------------------------------------------------------------------------- */
int gca_enqueue(void* queue, int dtype_queue, long long K_global, long long k_begin, long long k_end, int d,
                const float* keys, int N, long long index, void* stream);

/* Same enqueue with the ring pointer in DEVICE memory, for CUDA-graph replay: state[0] = pointer (read by the
 * kernel, advanced by N modulo K_global when the kernel finishes), state[1] = scratch ticket, must start at 0. */
int gca_enqueue_devptr(void* queue, int dtype_queue, long long K_global, long long k_begin, long long k_end, int d,
                       const float* keys, int N, long long* state, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * InfoNCE head, single device.  One call = RGBMoCo._compute_logit (mem_moco.py:29-49) + NCESoftmaxLoss
 * (lib/memory/criterion.py:34-45) + the rank of the positive that `accuracy` (lib/evaluation/metric.py:44-67)
 * derives from the logits + (optionally) dLoss/dq, without the [B, K+1] logits ever reaching HBM.
 *   q, k        : [B, d] fp32 (k is treated as a constant, mem_moco.py:69)
 *   queue       : [K, d] (dtype_queue) -- read in place; the caller orders it before this step's enqueue
 *   outputs     : loss_mean[1]   = mean_b (lse_b - pos_b)                   (criterion.py:44)
 *                 loss_rows[B], lse[B] (log-sum-exp over the K+1 logits), pos_logit[B] = q_b.k_b / T
 *                 rank_gt[B]     = number of negatives STRICTLY greater than the positive (top-k hit <=> < k).  May be
 *                                  NULL when only top_hits is wanted: the tcgen05 kernel then stops counting in a warp
 *                                  once all of its rows are past GCA_TOPK_RANK_CAP (exact below it, which is all that
 *                                  top-1 / top-5 need)
 *                 top_hits[2]    = #rows with rank_gt < 1, #rows with rank_gt < 5 (the counts behind
 *                                  accuracy(topk=(1,5)), train_video_contrast_dis.py:428); may be NULL
 *                 dq_unit[B, d]  = d loss_mean / d q (may be NULL: forward only)
 *                 logits_out     = [B, K+1] fp32 materialised logits (may be NULL; FFMA algo only)
 *   workspace   : gca_infonce_workspace_bytes(...) bytes of scratch, ZERO-FILLED ONCE by the caller before first use
 *                 (its first 256 bytes hold self-resetting tickets and a grid-barrier generation counter)
 * Deterministic: partial results are merged in a fixed order (no floating-point atomics).
 * --------------------------------------------------------------------------------------------------------- */
size_t gca_infonce_workspace_bytes(int B, long long K, int d, int dtype_queue, int algo);

int gca_infonce_fwd(const float* q, const float* k, const void* queue, int dtype_queue,
                    int B, long long K, int d, float inv_T, int algo,
                    float* loss_mean, float* loss_rows, float* lse, float* pos_logit, int* rank_gt, int* top_hits,
                    float* dq_unit, float* logits_out,
                    void* workspace, size_t workspace_bytes, void* stream);

/* One whole trainer-side head step in two launches: gca_infonce_fwd followed by the enqueue of `enqueue_keys[N, d]`
 * into the SAME queue (RGBMoCo.forward, mem_moco.py:60-88: logits from the queue as it was, then _update_memory, then
 * _update_pointer), the enqueue riding in the finalize launch.  Pointer: `state` == NULL -> `index` (host value, the
 * caller advances its copy); `state` != NULL -> device-resident {pointer, ticket} as in gca_enqueue_devptr. */
int gca_moco_step(const float* q, const float* k, void* queue, int dtype_queue, int B, long long K, int d, float inv_T,
                  int algo, const float* enqueue_keys, int N, long long index, long long* state,
                  void* keys_ready_event /* cudaEvent_t or NULL: the stream waits on it before the enqueue launch, so a
                                            key all-gather on another stream can overlap the queue sweep */,
                  float* loss_mean, float* loss_rows, float* lse, float* pos_logit, int* rank_gt, int* top_hits,
                  float* dq_unit, void* workspace, size_t workspace_bytes, void* stream);

/* gca_moco_step with the tail of the projection head fused in (Normalize(2) after the last Linear of ProjectHead,
 * lib/modeling/project_head.py:4-10, 22-28, applied by both encoders before RGBMoCo.forward): zq and zk are the
 * UN-normalised [B, d] outputs of that Linear.  The first launch L2-normalises both rows (x / max(||x||, 1e-12)) while it
 * prepares the bf16 query block and the positives; the last launch pushes the gradient back through the normalisation,
 *     dz_unit = (g - (g . q) q) / ||zq||,   g = d loss / d q,   q = zq / ||zq||,
 * and enqueues the normalised keys (enqueue_keys == NULL, N == B) or caller-gathered ones.  k_hat_out [B, d] receives the
 * normalised keys (may be NULL).  Same launches as gca_moco_step; bf16 queue with d == 128 only (tcgen05 family). */
int gca_moco_step_proj(const float* zq, const float* zk, void* queue, int dtype_queue, int B, long long K, int d, float inv_T,
                       int algo, const float* enqueue_keys, int N, long long index, long long* state,
                       float* loss_mean, float* loss_rows, float* lse, float* pos_logit, int* rank_gt, int* top_hits,
                       float* dz_unit, float* k_hat_out, void* workspace, size_t workspace_bytes, void* stream);

/* Stage 1 of gca_infonce_fwd on its own: only the queue-streaming kernel, leaving the per-split partials in the
 * workspace (layout: csrc/gca_common.cuh).  For profiling / roofline timing of the dominant kernel.
 * want_acc: bit 0 = also accumulate the gradient partials; bit 1 = skip the per-step q -> bf16 / positive-logit
 * preparation launch and reuse what the previous call left in this workspace (times the streaming kernel alone). */
int gca_infonce_partials(const float* q, const float* k, const void* queue, int dtype_queue,
                         int B, long long K, int d, float inv_T, int algo, int want_acc,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Two-pass backward by recomputation (north_star formulation): dq = grad_scale * d sum_b(loss_b) / dq, given the
 * row log-sum-exp from the forward pass.  The queue must still hold the rows the forward pass saw (i.e. call
 * before the enqueue, or on a snapshot): mem_moco.py:72-73 take the logits from a clone made before :82. */
int gca_infonce_bwd(const float* q, const float* k, const void* queue, int dtype_queue,
                    int B, long long K, int d, float inv_T, int algo,
                    const float* lse, float grad_scale, float* dq,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * InfoNCE head with the queue sharded along K over `W` ranks (replaces the replicated queue of
 * tools/train_video_contrast_dis.py:233-242 + mem_moco.py:81-83).  Per step and rank:
 *   1. gca_infonce_shard_fwd   : all B_glob gathered rows against the local shard -> online-softmax partials
 *   2. (caller) all-gather of part_max / part_sum / part_cnt over ranks           -> [W, B_glob] each
 *   3. gca_infonce_shard_combine: merge the W partials and the positive -> lse, loss rows, rank; scales this
 *                                 rank's part_acc by exp(part_max - lse) in place
 *   4. (caller) reduce-scatter(sum) of part_acc over ranks                         -> acc [B_loc, d]
 *   5. gca_infonce_shard_finish: dq_unit = ((exp(pos - lse) - 1) k + acc) / (T * B_loc), loss_mean over B_loc
 * part_max is in natural-log units (max logit of the shard), part_sum = sum exp(logit - part_max),
 * part_acc[b, :] = sum_j exp(logit_bj - part_max_b) queue_j.
 * --------------------------------------------------------------------------------------------------------- */
int gca_infonce_shard_fwd(const float* q, const float* k, const void* shard, int dtype_queue,
                          int B, long long K_shard, int d, float inv_T, int algo,
                          float* pos_logit, float* part_max, float* part_sum, int* part_cnt, float* part_acc,
                          void* workspace, size_t workspace_bytes, void* stream);

int gca_infonce_shard_combine(const float* all_max, const float* all_sum, const int* all_cnt, int W, int rank_id,
                              int B, int d, const float* pos_logit,
                              float* lse, float* loss_rows, int* rank_gt, float* part_acc, void* stream);

int gca_infonce_shard_finish(const float* acc, const float* k, const float* pos_logit, const float* lse,
                             const float* loss_rows, int B_loc, int d, float inv_T,
                             float* dq_unit, float* loss_mean, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Key all-gather over NVLink peer memory (replaces Trainer._global_gather, tools/train_video_contrast_dis.py:182-187,
 * for the momentum keys every replica enqueues, train...:222).  One launch per step and rank, no NCCL call on the data
 * path; safe to capture in a CUDA graph (the step counter is device-resident).
 *   mailboxes : DEVICE array of W pointers; entry p is rank p's mailbox allocation as mapped into THIS process
 *               (symmetric / IPC memory; entry `rank` is the local one).  Every mailbox is gca_keys_exchange_bytes()
 *               long and zero-filled before the first step; all ranks call once per step, in the same order.
 *   keys_local: [B, d] fp32 of this rank;  all_k: [W*B, d] fp32 out, rank-major rows (== torch.cat(all_gather)).
 *   xstate    : 4 x int64 device words, zero-initialised: [0] step counter, [1] ticket, [2] raised to 1 if a peer did
 *               not show up within timeout_ms (0 = wait for ever); all_k is then incomplete for that step.
 * --------------------------------------------------------------------------------------------------------- */
size_t gca_keys_exchange_bytes(int B, int d, int W);
int gca_keys_exchange(const float* keys_local, int B, int d, int W, int rank, void* const* mailboxes,
                      float* all_k, long long* xstate, int timeout_ms, void* stream);

/* The data-parallel head step with the key gather fused in (RGBMoCo.forward with all_k = _global_gather(k),
 * train...:222 + mem_moco.py:60-88): gca_moco_step whose enqueue rows are the keys of ALL W ranks, moved through the
 * mailboxes above instead of an all_k buffer.  Extra CTAs of the first launch push this rank's k[B, d] to every peer
 * while the queue is being swept; the enqueue CTAs of the last launch wait for the peers' flags and write the W*B
 * gathered rows (rank-major, like torch.cat(all_gather)) straight from the local mailbox into the ring.  Same
 * launches as gca_moco_step, no collective call, no side stream.  bf16 queue with d == 128 only (tcgen05 family);
 * `state` (device ring pointer) is required; mailboxes / xstate / timeout_ms as for gca_keys_exchange (both entry
 * points advance the same step counter, so they can share a mailbox if every rank issues the same call sequence). */
int gca_moco_step_peer(const float* q, const float* k, void* queue, int dtype_queue, int B, long long K, int d, float inv_T,
                       int algo, int W, int rank, void* const* mailboxes, long long* xstate, int timeout_ms,
                       long long* state, float* loss_mean, float* loss_rows, float* lse, float* pos_logit, int* rank_gt,
                       int* top_hits, float* dq_unit, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * NPID instance bank (RGBMem / CMCMem, lib/memory/mem_bank.py:15-90; the MEM_TYPE 'bank' branch of lib/memory/build.py:6-9).
 *   bank   [n_data, d] fp32 row-major (L2-normalised rows), x [B, d] fp32, idx [B, K1] int64 (K1 = K + 1 sampled bank rows
 *   per feature row, column 0 the positive: mem_bank.py:64-66), d a multiple of 4 up to 1024.
 * gca_bank_logits : logits[b, j] = <bank[idx[b, j]], x[b]> * inv_T      (index_select + bmm + div, mem_bank.py:67-73, 29-39)
 *                   without materialising the [B, K1, d] gather.  An index outside [0, n_data) yields NaN.
 * gca_bank_dx     : dx[b] = inv_T * sum_j g_logits[b, j] * bank[idx[b, j]]  -- autograd of the above w.r.t. x (the bank is a
 *                   buffer).  Fixed-order sums (deterministic).
 * gca_bank_update : bank[y[n]] <- normalize(momentum * bank[y[n]] + one_minus_momentum * x[n]), n < N  (mem_bank.py:15-27):
 *                   every row is computed from the bank as it was BEFORE the call; of a duplicated index the last
 *                   occurrence wins (index_copy_ on the CPU).  Pass one_minus_momentum = (float)(1.0 - momentum) computed in
 *                   double, as the reference's Python expression does. */
int gca_bank_logits(const float* x, const float* bank, const long long* idx, int B, int K1, int d, long long n_data,
                    float inv_T, float* logits, void* stream);
size_t gca_bank_dx_workspace_bytes(int B, int d);
int gca_bank_dx(const float* g_logits, const float* bank, const long long* idx, int B, int K1, int d, long long n_data,
                float inv_T, float* dx, void* workspace, size_t workspace_bytes, void* stream);
size_t gca_bank_update_workspace_bytes(int N, int d);
int gca_bank_update(float* bank, const float* x, const long long* y, int N, int d, long long n_data, float momentum,
                    float one_minus_momentum, void* workspace, size_t workspace_bytes, void* stream);

/* Host-visible completion word for callers whose step results land in pinned host memory (zero-copy outputs): once armed
 * for a workspace, the last launch of every gca_infonce_fwd / gca_moco_step* call on that workspace stores the number of
 * steps completed on it (1, 2, ...) into *host_word -- a 32-bit word in pinned, device-addressable host memory -- after all
 * of the step's stores are visible system-wide and its host-resident inputs have been read.  A host thread that polls the
 * word replaces the stream synchronisation after the step (~1 us instead of the driver's wake-up path).  host_word == NULL
 * disarms.  Synchronises `stream` (a set-up call, not a per-step one). */
int gca_workspace_set_done_flag(void* workspace, unsigned int* host_word, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Launch plans.  A head step is 3-4 kernel launches whose parameters do not change from step to step when the caller
 * uses static buffers and the device-resident ring pointer (`state` != NULL) -- what the reference's training loop does
 * every iteration (tools/train_video_contrast_dis.py:405-428).  Between gca_plan_begin() and gca_plan_end() the step
 * entry points of the tcgen05 family (gca_moco_step, gca_moco_step_proj, gca_moco_step_peer on a bf16 queue with
 * d == 128) called by THIS thread are recorded instead of launched; gca_plan_run() re-issues the recorded launches
 * on `stream` with three host-side launch calls, keeping the programmatic dependencies between them and across
 * consecutive steps (the first launch of step s+1 is resident while the last launch of step s drains -- a CUDA-graph
 * launch per step cannot do that).  Every pointer given while recording must stay valid for the life of the plan;
 * `state` must be non-NULL (a host `index` would be frozen into the plan).  gca_plan_end() fails with
 * GCA_ERR_UNSUPPORTED if a call made while recording was not recordable (launches of that call which bypass the
 * recorder were issued on the spot, the others dropped: its results are undefined).
 * Plans are bound to the device that was current in gca_plan_begin(); gca_plan_run is not thread-safe per plan. */
typedef struct gca_plan gca_plan;
int gca_plan_begin(void);
int gca_plan_end(gca_plan** plan);
int gca_plan_run(gca_plan* plan, void* stream);
int gca_plan_launches(const gca_plan* plan);    /* kernel launches per gca_plan_run */
void gca_plan_destroy(gca_plan* plan);

/* The K-sharded head step with every exchange over NVLink peer memory (the partition of SURVEY.md section 8e; replaces the
 * replicated queue + key gather of tools/train_video_contrast_dis.py:182-187, 222, 233-242 and the three NCCL collectives
 * of the gca_infonce_shard_* sequence above).  One call per step and rank launches
 *   1. the q|k gather            qk_loc [2, B_loc, d] of every rank through the mailboxes (the gca_keys_exchange protocol on
 *                                rows 2*B_loc; head of every mailbox).  tcgen05 family: rides in the prep launch (push CTAs
 *                                + row warps that read their q, k rows from the mailbox; qk_all unused, may be NULL);
 *                                otherwise one exchange launch into qk_all [2, W*B_loc, d]
 *   2. the shard sweep           all W*B_loc rows against this rank's [K/W, d] shard (prep + stream kernel)
 *   3. the split merge           one CTA per row; the merged partials of ALL rows (acc, max, sum, count) stay in this rank's
 *                                own mailbox; the last CTA publishes the step to every peer (one release store each)
 *   4. the cross-rank merge      one CTA per LOCAL row waits for the W flags, pulls the row's W partials from the peers'
 *                                mailboxes (remote loads, all in flight at once), merges them in rank order (fixed
 *                                association: deterministic), adds the positive and writes loss rows, lse, rank,
 *                                dq_unit [B_loc, d] (scaled by 1 / (T * B_loc)) and the local mean loss
 *      + the sharded enqueue     extra CTAs of launch 4 write the gathered keys that land in this rank's slots;
 *                                `enq_state` = device ring pointer over the GLOBAL ring of K slots
 * No collective call; CUDA-graph capturable.  Mailboxes: symmetric memory of gca_shard_peer_bytes() bytes per rank, zeroed
 * before the first step; `pstate`: 8 x int64 device words, zero-initialised ([0..2] gather step / ticket / timeout flag,
 * [4..6] merge step / ticket / timeout flag).  pos_logit_all [W*B_loc]; loss_rows, lse, rank_gt [B_loc] (local rows);
 * top_hits [2] may be NULL.  Sharding changes no result: same outputs as gca_moco_step on the concatenated queue up to
 * the association order of the fp32 merges (tests/sharded_graph_worker.py). */
size_t gca_shard_peer_bytes(int B_loc, int d, int W);
int gca_shard_step_peer(const float* qk_loc, void* shard, int dtype_queue, int B_loc, long long K, int d, float inv_T,
                        int algo, int W, int rank, void* const* mailboxes, long long* pstate, int timeout_ms,
                        long long* enq_state, float* qk_all, float* loss_mean, float* loss_rows, float* lse,
                        float* pos_logit_all, int* rank_gt, int* top_hits, float* dq_unit,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * BatchNorm1d (+ ReLU) epilogue of the SimSiam projection / prediction MLPs (replaces the nn.BatchNorm1d -> nn.ReLU pairs
 * of ProjectionMLP / PredictionMLP, lib/modeling/project_head.py:36-76; the nn.Linear in front stays a library GEMM).
 * x, y, dy, dx: [B, C] fp32 row-major; gamma, beta, running_*, save_*, dgamma, dbeta: [C] (gamma / beta NULL = 1 / 0).
 * Training mode (use_running = 0): batch statistics (biased variance normalises, the unbiased one feeds running_var with
 * `momentum`, exactly nn.BatchNorm1d); eval mode (use_running = 1): the running statistics normalise, nothing is updated.
 * save_mean / save_invstd are what the backward needs; relu != 0 applies max(., 0) to y and its mask to dy.
 * One launch each; deterministic (fixed reduction order). */
int gca_bn1d_fwd(const float* x, int B, int C, const float* gamma, const float* beta, float eps, float momentum, int relu,
                 int use_running, float* running_mean, float* running_var, float* y, float* save_mean, float* save_invstd,
                 void* stream);
int gca_bn1d_bwd(const float* x, const float* dy, int B, int C, const float* gamma, const float* beta,
                 const float* save_mean, const float* save_invstd, int relu, int use_running, float* dx, float* dgamma,
                 float* dbeta, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Temporal clip-graph head.  Replaces, for one GCN layer (the shipped default), the part of
 * TemporalGraphAug.forward (lib/ops/module_wrappers/temporal_graph.py:227-239) between the 1x1x1 convolutions:
 * similarity + row softmax (:161-176), hop mask and theta(hop) weights (:25-36, :204-210), relaxed-Bernoulli
 * re-sampling with caller-supplied uniforms (:187-192), aggregation + skip (:59-62).  The convolutions / pooling
 * (:119-129, :58) stay with cuDNN in the caller.
 *   gq, gk   : [B, Cq, T, S]  projections in conv-output layout (S = H'*W'); the dot product runs over (Cq, S)
 *   support  : [B, C, T, HW]  GCN conv output
 *   u        : [B, T, T]      uniforms in [0, 1) drawn by the caller (torch.rand, same generator state as the
 *                             reference's rsample would consume)
 *   outputs  : sim, adj, s [B, T, T] (kept for backward), y [B, C, T, HW]
 *   T <= 32; S % 4 == 0 or S == 1 is handled; any HW.
 * --------------------------------------------------------------------------------------------------------- */
int gca_graph_fwd(const float* gq, const float* gk, int Cq, int S, const float* support, int C, int HW,
                  int T, int B, const float* u, float alpha, int max_hop, float temperature, unsigned flags,
                  float* sim, float* adj, float* s, float* y, void* workspace, size_t workspace_bytes, void* stream);

int gca_graph_bwd(const float* gq, const float* gk, int Cq, int S, const float* support, int C, int HW,
                  int T, int B, const float* sim, const float* adj, const float* s, const float* dy,
                  float alpha, int max_hop, float temperature, unsigned flags,
                  float* d_gq, float* d_gk, float* d_support,
                  void* workspace, size_t workspace_bytes, void* stream);
/* The same two calls with the default-OFF variants (opts == NULL or opts->flags == 0: identical to the calls above).
 * The backward additionally takes the uniforms `u` of the forward (keep mask of EDGE_DROP, pre-normalisation s of SYMNORM). */
int gca_graph_fwd_ex(const float* gq, const float* gk, int Cq, int S, const float* support, int C, int HW,
                     int T, int B, const float* u, float alpha, int max_hop, float temperature, const GcaGraphOpts* opts,
                     float* sim, float* adj, float* s, float* y, void* workspace, size_t workspace_bytes, void* stream);
int gca_graph_bwd_ex(const float* gq, const float* gk, int Cq, int S, const float* support, int C, int HW,
                     int T, int B, const float* sim, const float* adj, const float* s, const float* dy, const float* u,
                     float alpha, int max_hop, float temperature, const GcaGraphOpts* opts,
                     float* d_gq, float* d_gk, float* d_support,
                     void* workspace, size_t workspace_bytes, void* stream);
/* scratch for gca_graph_fwd / gca_graph_bwd (d_logit + per-chunk pair-dot partials; only touched for large feature maps) */
size_t gca_graph_workspace_bytes(int B, int T);

/* ---------------------------------------------------------------------------------------------------------
 * SimSiam negative cosine.  Replaces D.forward 'v2' (lib/memory/criterion.py:53-62 == lib/modeling/
 * graph_wrappers.py:99-108): loss = -mean_b cos(p_b, stopgrad(z_b)), eps = 1e-8 as F.cosine_similarity.
 *   loss[1], cos_rows[B] (may be NULL), dp_unit[B, d] = d loss / d p (may be NULL)
 * --------------------------------------------------------------------------------------------------------- */
int gca_negcos_fwd_bwd(const float* p, const float* z, int B, int d, float* loss, float* cos_rows, float* dp_unit,
                       void* workspace, size_t workspace_bytes, void* stream);
size_t gca_negcos_workspace_bytes(int B, int d);

/* ---------------------------------------------------------------------------------------------------------
 * Retrieval: cosine similarity + per-query top-k.  Replaces the arithmetic of topk_retrieval
 * (tools/video_retrieval.py:174-186): optional L2 normalisation, cosine distance matrix, argsort -- only the k
 * nearest are produced (k <= 64), most similar first, ties towards the lower gallery index.
 *   queries [Nq, d], gallery [Ng, d] fp32; idx_out [Nq, k] int32; val_out [Nq, k] fp32 cosine (may be NULL)
 * --------------------------------------------------------------------------------------------------------- */
size_t gca_sim_topk_workspace_bytes(int Nq, int Ng, int d, int k);
int gca_sim_topk(const float* queries, const float* gallery, int Nq, int Ng, int d, int k, int normalize,
                 int* idx_out, float* val_out, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Momentum (EMA) update of the key encoder in one multi-tensor launch.  Replaces Trainer._momentum_update
 * (tools/train_video_contrast_dis.py:176-180, :440): for every parameter  p_ema = m * p_ema + (1 - m) * p.
 *   chunk_table : DEVICE array of nchunks descriptors {float* ema; const float* src; long long n;}
 *                 (gca_ema_chunk_bytes() each), built once per model by the caller; a descriptor covers at most
 *                 a few thousand contiguous elements of one parameter tensor
 *   one_minus_momentum : the reference's `alpha = 1 - m`, a Python double rounded ONCE to fp32 -- the caller passes
 *                 (float)(1.0 - m); computing 1.f - momentum in fp32 is off by up to 5e-5 relative for m = 0.999
 * --------------------------------------------------------------------------------------------------------- */
size_t gca_ema_chunk_bytes(void);
int gca_ema_update(const void* chunk_table, int nchunks, float momentum, float one_minus_momentum, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GCA_B200_H */
