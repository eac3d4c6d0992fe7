// K9: key all-gather over NVLink peer memory, one launch per step and rank (no NCCL call on the data path).
// Replaces Trainer._global_gather (tools/train_video_contrast_dis.py:182-187: world x ones_like + all_gather + cat)
// for the [B, d] momentum keys that every replica enqueues (train...:222, mem_moco.py:81-83).
//
// Every rank owns one "mailbox" allocation that all peers can address (symmetric memory, mapped over NVLink by the
// caller):   slots [2 parities][W ranks][B*d floats]   then   flags [2 parities][W ranks][XCHG_SLICES] (u64)
// Step s uses parity s & 1.  CTA (p, c) of rank r
//   1. stores slice c of the local keys into rank p's mailbox slot [s&1][r]          (remote 128-bit stores)
//   2. publishes  flag[s&1][r][c] = s + 1  in rank p's mailbox                        (fence + release store)
//   3. waits until its own mailbox shows flag[s&1][p][c] >= s + 1                      (rank p's slice has landed)
//   4. copies that slice from its own mailbox into all_k[p*B ...]                      (local, L1-bypassing loads)
// No CTA waits for another CTA of the same GPU, and a peer's push never depends on anything but that peer having
// finished its previous step, so there is no circular wait; double buffering by parity is enough because a rank can
// only be one step ahead (its step s+1 cannot finish before every peer has pushed step s+1, i.e. finished step s).
// The spin has an optional wall-clock bound: on expiry the kernel raises xstate[2] and returns instead of hanging.
// Flags are monotone step numbers, so nothing is ever reset; the step counter lives in device memory (xstate[0]) so
// a captured CUDA graph can be replayed.
#include "gca_common.cuh"
#include "launch_plan.cuh"

namespace gca {

constexpr int XCHG_THREADS = 512;

__global__ void __launch_bounds__(XCHG_THREADS)
keys_exchange_kernel(const float4* __restrict__ keys_local, const PeerXchg X, float4* __restrict__ all_k, const int parts)
{
    const int p = blockIdx.x;                         // peer this CTA talks to
    const int c = blockIdx.y;                         // slice of the key block
    const unsigned long long step = *reinterpret_cast<volatile unsigned long long*>(X.xstate);
    xchg_push_slice(X, keys_local, step, p, c);       // 1 + 2
    __shared__ int ok;
    if (threadIdx.x == 0) ok = xchg_wait_slice(X, step, p, c) ? 1 : 0;      // 3
    __syncthreads();
    if (ok) {                                         // 4. my mailbox -> all_k rows of rank p
        const int per = (X.n4 + XCHG_SLICES - 1) / XCHG_SLICES;
        const int lo = c * per, hi = min(X.n4, lo + per);
        const float4* src = xchg_slot(X.mailboxes[X.rank], X, (int)(step & 1ull), p);
        if (parts <= 1) {
            float4* out = all_k + (size_t)p * X.n4;
            for (int i = lo + threadIdx.x; i < hi; i += XCHG_THREADS) out[i] = ld_cg_f4(src + i);
        } else {
            // the local block is [parts][n4 / parts] (e.g. q rows | k rows); the output is part-major: [parts][W][n4 / parts]
            const int pn = X.n4 / parts;
            for (int i = lo + threadIdx.x; i < hi; i += XCHG_THREADS) {
                const int h = i / pn, j = i - h * pn;
                all_k[((size_t)h * X.W + p) * pn + j] = ld_cg_f4(src + i);
            }
        }
    }
    // step counter: the last CTA of this launch advances it (every CTA has read it before taking a ticket)
    __shared__ int last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long t = atomicAdd(X.xstate + 1, 1ull);
        last = (t == (unsigned long long)gridDim.x * gridDim.y - 1);
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        X.xstate[1] = 0;
        __threadfence();
        X.xstate[0] = step + 1;
    }
}

// The push half alone (steps 1 + 2), for the peer-fused replica step.  It runs on a side stream forked off the step's stream
// (a parallel branch when the step is captured into a CUDA graph) and is joined after the step's last launch: the
// system-scope release of a push CTA costs an NVLink round trip, and inside the prep launch it held back the queue sweep,
// which waits for the whole prep grid.  The flags are consumed by the enqueue CTAs at the end of the step (finalize.cu).
__global__ void __launch_bounds__(256)
keys_push_kernel(const float4* __restrict__ keys_local, const PeerXchg X)
{
    const unsigned long long step = *reinterpret_cast<volatile unsigned long long*>(X.xstate);
    xchg_push_slice(X, keys_local, step, blockIdx.x, blockIdx.y);
}

static thread_local bool t_push_forked = false;        // the key push of the step being issued went out on the side stream

int keys_push_fork(const float* keys_local, const PeerXchg& X, cudaStream_t st)
{
    cudaStream_t side;
    const int rc = side_stream_fork(st, &side);
    if (rc != GCA_OK) return rc;
    t_push_forked = true;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(X.W, XCHG_SLICES); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = side;
    GCA_CUDA(launch_ex(&cfg, keys_push_kernel, (const float4*)keys_local, X));
    GCA_LAUNCH_CHECK("keys_push_kernel");
    count_launch(1);
    return GCA_OK;
}

// joins the side stream if (and only if) this step's key push went out on it
int keys_push_join(cudaStream_t st)
{
    if (!t_push_forked) return GCA_OK;
    t_push_forked = false;
    return side_stream_join(st);
}

}  // namespace gca

extern "C" size_t gca_keys_exchange_bytes(int B, int d, int W)
{
    if (B <= 0 || d <= 0 || W <= 0) return 0;
    const size_t slots = (size_t)2 * W * B * d * sizeof(float);
    const size_t flags = (size_t)2 * W * gca::XCHG_SLICES * sizeof(unsigned long long);
    return gca::align_up(slots + flags, 256);
}

namespace gca {
// rows of every rank gathered through the mailboxes; parts > 1: the local block is [parts][B / parts, d] and the output is
// part-major, [parts][W][B / parts, d] (the q | k gather of the K-sharded step)
int keys_exchange_launch(const float* keys_local, int B, int d, int W, int rank, void* const* mailboxes, float* all_k,
                         long long* xstate, int timeout_ms, int parts, cudaStream_t st)
{
    PeerXchg X{};
    X.mailboxes = (char* const*)mailboxes; X.W = W; X.rank = rank; X.n4 = B * d / 4;
    X.xstate = (unsigned long long*)xstate;
    X.timeout_ns = timeout_ms > 0 ? (unsigned long long)timeout_ms * 1000000ull : 0ull;
    dim3 grid(W, XCHG_SLICES);
    keys_exchange_kernel<<<grid, XCHG_THREADS, 0, st>>>((const float4*)keys_local, X, (float4*)all_k, parts);
    GCA_LAUNCH_CHECK("keys_exchange_kernel");
    count_launch(1);
    return GCA_OK;
}
}  // namespace gca

extern "C" int gca_keys_exchange(const float* keys_local, int B, int d, int W, int rank, void* const* mailboxes,
                                 float* all_k, long long* xstate, int timeout_ms, void* stream)
{
    using namespace gca;
    GCA_CHECK_ARG(keys_local && mailboxes && all_k && xstate, "gca_keys_exchange: null pointer");
    GCA_CHECK_ARG(B > 0 && d > 0 && d % 4 == 0, "gca_keys_exchange: need B > 0 and d %% 4 == 0 (B=%d d=%d)", B, d);
    GCA_CHECK_ARG(W >= 1 && W <= 64 && rank >= 0 && rank < W, "gca_keys_exchange: bad rank %d of %d", rank, W);
    return keys_exchange_launch(keys_local, B, d, W, rank, mailboxes, all_k, xstate, timeout_ms, 1, (cudaStream_t)stream);
}
