// Shared helpers for libgca_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/gca_b200.h"

namespace gca {

// thread-local error text behind gca_last_error()
char* err_buf();
int set_err(int code, const char* fmt, ...);

#define GCA_CHECK_ARG(cond, ...) do { if (!(cond)) return gca::set_err(GCA_ERR_BAD_ARG, __VA_ARGS__); } while (0)
#define GCA_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return gca::set_err(GCA_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define GCA_LAUNCH_CHECK(name) do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) \
    return gca::set_err(GCA_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e_)); } while (0)

void count_launch(int n);       // bookkeeping behind gca_launch_count()
int sm_count_cached();          // SMs of the current device (cached per device); <=0 on error

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// element of the split statistics part_max / part_sum / part_cnt (layout below)
__host__ __device__ __forceinline__ size_t part_stat_index(int split, int row, int nsplit) { return (size_t)row * nsplit + split; }

// ---------------------------------------------------------------------------------------------
// Workspace layout shared by the InfoNCE stream kernels and the finalize kernels.
//   [0, 256)                 : control block, 64 words: [0] ticket, [2..3] packed loss / hit word, [4] top-1, [5] top-5 hits,
//                              [6] out-of-range-logit flag (all self-resetting); [12..13] device-addressable pointer to a
//                              32-bit host completion word (gca_workspace_set_done_flag; 0 = off), [14] CTAs of the finalize
//                              launch that are done, [15] number of completed steps (what is written to the completion word)
//   part_max [Bpad, nsplit]  : natural-log max logit of the split        (-inf when the split is empty)
//   part_sum [Bpad, nsplit]  : sum exp(logit - part_max)
//   part_cnt [Bpad, nsplit]  : #negatives > positive in the split
//                              (row-major over the ROW: the finalize CTA of a row reads its nsplit statistics with three
//                              coalesced loads -- split-major they were 32 different cache lines per load, ~480 L1 wavefronts
//                              per CTA queued in front of the gradient partials; part_stat_index() below)
//   part_acc [nsplit, Bpad, d] : sum exp(logit - part_max) * queue_row
//   pos_tmp  [Bpad]          : positive logits (backward-recompute entry point)
// Bpad = B rounded up to 128 so both kernel families index it the same way.
// ---------------------------------------------------------------------------------------------
struct InfoNceWs {
    unsigned int* counter;
    float* part_max;
    float* part_sum;
    int*   part_cnt;
    float* part_acc;
    float* pos_tmp;     // [Bpad] positive logits when the caller gave no output buffer
    float* pos_ws;      // [Bpad] positive logits for the tcgen05 kernel (written by its prep kernel)
    void*  q_bf16;      // [Bpad, d] bf16 copy of q for the tcgen05 kernel
    float* k_hat;       // [Bpad, d] L2-normalised keys (projection-tail fusion: written by the prep kernel)
    float* inv_nq;      // [Bpad] 1 / ||zq|| of the same mode
    int    nsplit;
    int    Bpad;
    size_t bytes;
};

// Key all-gather through peer "mailboxes" (exchange.cu has the protocol): either a stand-alone kernel
// (gca_keys_exchange) or fused into the head step -- pushed by extra CTAs of the prep kernel, consumed directly by the
// enqueue CTAs of the finalize kernel (gca_moco_step_peer).  mailboxes == nullptr: off.
constexpr int XCHG_SLICES = 8;
struct PeerXchg {
    char* const* mailboxes;          // device array [W]: every rank's mailbox as mapped into this process
    int W, rank;
    int n4;                          // float4 per rank and step (B * d / 4)
    unsigned long long* xstate;      // [0] step counter, [1] ticket, [2] timeout flag
    unsigned long long timeout_ns;   // 0 = wait for ever
};

// Cross-rank merge of the K-sharded step over peer memory (finalize.cu).  Region of every rank's mailbox at byte offset `off`:
//     rows [2 parities][Bg * (d + 4) floats]   then   flags [2][W] (u64),          Bg = W * Bl
// A rows block holds THIS rank's merged partials of all Bg global rows against its shard: acc [Bg, d], then max [Bg],
// sum [Bg], count [Bg] (int bits), pad [Bg].  Step s uses parity s & 1 (the q|k gather at the head of every step keeps the
// ranks within one step of each other).  The split-merge launch (FIN_SHARD) writes the block locally; its last CTA (ticket)
// publishes flag[par][rank] = s + 1 into every peer's mailbox -- ONE system-scope release per rank and step.  The merge CTA of
// a local row waits for the W flags in its own mailbox and then PULLS that row's W partials straight from the peers'
// mailboxes (remote loads over NVLink, all in flight at once) and merges them in rank order.
struct PeerMerge {
    char* const* mailboxes;          // device array [W]; nullptr = off
    unsigned long long off;          // byte offset of the merge region inside every mailbox
    int W, rank, Bl, d;
    unsigned long long* mstate;      // device memory: [0] merge step counter, [1] ticket, [2] timeout flag
    unsigned long long* gather_state;   // non-null: the q|k gather rode in the prep launch; its step counter advances with mstate
    unsigned long long timeout_ns;   // 0 = wait for ever
    int wait;                        // FIN_FULL launch: wait for the W ranks' flags, then pull and merge their partials
};

static inline int infonce_bpad(int B) { return (B + 127) / 128 * 128; }
// upper bound on the number of K-splits any kernel family uses (2 CTAs worth per SM, at least 1)
int infonce_max_splits(int B);
int side_stream_fork(cudaStream_t st, cudaStream_t* side);                                     // gca_api.cu
int side_stream_join(cudaStream_t st);
int keys_push_fork(const float* keys_local, const PeerXchg& X, cudaStream_t st);              // exchange.cu: side-stream push ...
int keys_push_join(cudaStream_t st);                                                          // ... joined after the step's last launch
int keys_exchange_launch(const float* keys_local, int B, int d, int W, int rank, void* const* mailboxes, float* all_k,
                         long long* xstate, int timeout_ms, int parts, cudaStream_t st);      // exchange.cu
InfoNceWs infonce_ws_carve(void* base, int B, int d, int nsplit);

}  // namespace gca

// ------------------------------------------------------------------ device helpers
namespace gca {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum in a fixed order (deterministic): warp shuffles, then warp 0 over the per-warp values
template <int kThreads>
__device__ __forceinline__ float block_sum(float v, float* smem_red /* >= kThreads/32 floats */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) smem_red[warp] = v;
    __syncthreads();
    float r = (threadIdx.x < kThreads / 32) ? smem_red[threadIdx.x] : 0.f;
    if (warp == 0) r = warp_sum(r);
    if (threadIdx.x == 0) smem_red[0] = r;
    __syncthreads();
    r = smem_red[0];
    __syncthreads();
    return r;
}

// ---- peer mailbox primitives (protocol: exchange.cu)
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_cg_f4(const float4* p) {       // L2 only: the line was written by a peer GPU
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ size_t xchg_flags_off(const PeerXchg& X) { return (size_t)2 * X.W * X.n4 * sizeof(float4); }
__device__ __forceinline__ unsigned long long* xchg_flag(char* box, const PeerXchg& X, int par, int from, int slice) {
    return reinterpret_cast<unsigned long long*>(box + xchg_flags_off(X)) + ((size_t)par * X.W + from) * XCHG_SLICES + slice;
}
__device__ __forceinline__ float4* xchg_slot(char* box, const PeerXchg& X, int par, int from) {
    return reinterpret_cast<float4*>(box) + ((size_t)par * X.W + from) * (size_t)X.n4;
}
// One CTA: store slice `c` of this rank's keys into rank p's mailbox and publish it.  The bar.sync orders every thread's
// stores before thread 0's release store (cumulativity), so one system-scope release per CTA is enough.
__device__ __forceinline__ void xchg_push_slice(const PeerXchg& X, const float4* keys_local, unsigned long long step, int p, int c)
{
    const int par = (int)(step & 1ull);
    const int per = (X.n4 + XCHG_SLICES - 1) / XCHG_SLICES;
    const int lo = c * per, hi = min(X.n4, lo + per);
    char* peer = X.mailboxes[p];
    float4* dst = xchg_slot(peer, X, par, X.rank);
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) dst[i] = __ldg(keys_local + i);
    __syncthreads();
    // (st.release.sys is itself a system-scope release: no separate fence in front of it)
    if (threadIdx.x == 0) st_release_sys_u64(xchg_flag(peer, X, par, X.rank, c), step + 1);
}
// One thread: wait until slice c of rank p's keys for `step` has landed in my mailbox.  Returns false on timeout.
__device__ __forceinline__ bool xchg_wait_slice(const PeerXchg& X, unsigned long long step, int p, int c)
{
    const unsigned long long* f = xchg_flag(X.mailboxes[X.rank], X, (int)(step & 1ull), p, c);
    if (ld_acquire_sys_u64(f) >= step + 1) return true;
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys_u64(f) < step + 1) {
        if (X.timeout_ns && globaltimer_ns() - t0 > X.timeout_ns) { atomicExch(X.xstate + 2, 1ull); return false; }
    }
    return true;
}

// ---- cross-rank merge region (PeerMerge)
__device__ __forceinline__ size_t pm_block_floats(const PeerMerge& M) { return (size_t)M.W * M.Bl * (size_t)(M.d + 4); }
__device__ __forceinline__ float* pm_rows(char* box, const PeerMerge& M, int par) {
    return reinterpret_cast<float*>(box + M.off) + (size_t)par * pm_block_floats(M);
}
__device__ __forceinline__ unsigned long long* pm_flag(char* box, const PeerMerge& M, int par, int from) {
    return reinterpret_cast<unsigned long long*>(box + M.off + (size_t)2 * pm_block_floats(M) * sizeof(float)) + (size_t)par * M.W + from;
}
// One thread: wait until rank `from` has published its partials of `step` into my mailbox.  False on timeout.
__device__ __forceinline__ bool pm_wait(const PeerMerge& M, int par, int from, unsigned long long step)
{
    const unsigned long long* f = pm_flag(M.mailboxes[M.rank], M, par, from);
    if (ld_acquire_sys_u64(f) >= step + 1) return true;
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys_u64(f) < step + 1) {
        if (M.timeout_ns && globaltimer_ns() - t0 > M.timeout_ns) { atomicExch(M.mstate + 2, 1ull); return false; }
    }
    return true;
}
// loads of split partials: L2 (same GPU) or system-scope relaxed (a peer's mailbox over NVLink)
__device__ __forceinline__ float ld_part(const float* p, bool remote) {
    if (!remote) return __ldcg(p);
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_part(const int* p, bool remote) {
    if (!remote) return __ldcg(p);
    int v;
    asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_part(const float4* p, bool remote) {
    if (!remote) return __ldcg(p);
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ float ld_queue(const float* p) { return *p; }
__device__ __forceinline__ float ld_queue(const __nv_bfloat16* p) { return __bfloat162float(*p); }

}  // namespace gca
