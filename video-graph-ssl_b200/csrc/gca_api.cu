// C-ABI entry points of the InfoNCE head + library-wide helpers.  See include/gca_b200.h for the contract.
#include "gca_common.cuh"
#include "infonce_params.cuh"
#include "launch_plan.cuh"
#include <string.h>
#include <atomic>

namespace gca {

char* err_buf()
{
    static thread_local char buf[512] = "";
    return buf;
}

int set_err(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

static std::atomic<long long> g_launches{0};
static thread_local LaunchPlan* t_plan = nullptr;        // non-null between gca_plan_begin and gca_plan_end on this thread
LaunchPlan* plan_recording() { return t_plan; }
void count_launch(int n)
{
    if (t_plan) t_plan->counted += n;                    // recorded, not launched: gca_plan_run counts them when they run
    else g_launches.fetch_add(n, std::memory_order_relaxed);
}

int sm_count_cached()
{
    static int cached_dev = -1, cached_sms = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (dev != cached_dev) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
        cached_sms = sms;
        cached_dev = dev;
    }
    return cached_sms;
}

// One side stream per device for work that may run beside the caller's stream (a parallel branch when the caller is
// capturing a CUDA graph): fork = the side stream waits for everything issued so far on `st`; join = `st` waits for
// everything issued on the side stream since the fork.  Not re-entrant: one fork / join pair at a time per device.
struct SideLane { cudaStream_t stream; cudaEvent_t fork, join; bool ok; };
static SideLane* side_lane()
{
    static SideLane lanes[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    SideLane& L = lanes[dev];
    if (!L.ok) {                                         // (first use is an un-captured warm-up call)
        if (cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&L.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&L.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        L.ok = true;
    }
    return &L;
}

bool plan_is_side_stream(cudaStream_t st)
{
    SideLane* L = side_lane();
    return L != nullptr && st == L->stream;
}

int side_stream_fork(cudaStream_t st, cudaStream_t* side)
{
    SideLane* L = side_lane();
    if (!L) return set_err(GCA_ERR_CUDA, "could not create the side stream");
    if (t_plan) {                                        // recording: the fork becomes a plan operation
        PlanOp op; op.kind = PLAN_FORK;
        t_plan->ops.push_back(std::move(op));
        *side = L->stream;
        return GCA_OK;
    }
    GCA_CUDA(cudaEventRecord(L->fork, st));
    GCA_CUDA(cudaStreamWaitEvent(L->stream, L->fork, 0));
    *side = L->stream;
    return GCA_OK;
}

int side_stream_join(cudaStream_t st)
{
    SideLane* L = side_lane();
    if (!L) return set_err(GCA_ERR_CUDA, "could not create the side stream");
    if (t_plan) {
        PlanOp op; op.kind = PLAN_JOIN;
        t_plan->ops.push_back(std::move(op));
        return GCA_OK;
    }
    GCA_CUDA(cudaEventRecord(L->join, L->stream));
    GCA_CUDA(cudaStreamWaitEvent(st, L->join, 0));
    return GCA_OK;
}

int infonce_max_splits(int B)
{
    // every kernel family uses at most (#SMs / row blocks) splits with >= 32-row blocks; 1024 bounds finalize's smem
    (void)B;
    int sms = sm_count_cached();
    if (sms < 1) sms = 148;
    return sms < 1024 ? sms : 1024;
}

InfoNceWs infonce_ws_carve(void* base, int B, int d, int nsplit)
{
    InfoNceWs w{};
    w.nsplit = nsplit;
    w.Bpad = infonce_bpad(B);
    char* p = (char*)base;
    size_t off = 0;
    w.counter = (unsigned int*)(p + off);              off += 256;
    const size_t rows = (size_t)nsplit * w.Bpad;
    w.part_max = (float*)(p + off);                    off += align_up(rows * sizeof(float), 256);
    w.part_sum = (float*)(p + off);                    off += align_up(rows * sizeof(float), 256);
    w.part_cnt = (int*)(p + off);                      off += align_up(rows * sizeof(int), 256);
    w.pos_tmp  = (float*)(p + off);                    off += align_up((size_t)w.Bpad * sizeof(float), 256);
    w.pos_ws   = (float*)(p + off);                    off += align_up((size_t)w.Bpad * sizeof(float), 256);
    w.q_bf16   = (void*)(p + off);                     off += align_up((size_t)w.Bpad * d * 2, 1024);
    w.part_acc = (float*)(p + off);                    off += align_up(rows * d * sizeof(float), 256);
    w.k_hat    = (float*)(p + off);                    off += align_up((size_t)w.Bpad * d * sizeof(float), 256);
    w.inv_nq   = (float*)(p + off);                    off += align_up((size_t)w.Bpad * sizeof(float), 256);
    w.bytes = off;
    return w;
}

static int pick_algo(int algo, int dtype_queue, int d)
{
    // d == 128 runs on the tensor cores in either queue precision; other widths keep the CUDA-core fp32 kernel
    if (algo == GCA_ALGO_AUTO) return d != 128 ? GCA_ALGO_FFMA : (dtype_queue == GCA_BF16 ? GCA_ALGO_TCGEN05 : GCA_ALGO_TC32);
    return algo;
}

static int check_infonce_args(const char* fn, const void* q, const void* k, const void* queue, int dtype_queue, int B,
                              long long K, int d, float inv_T, int algo)
{
    GCA_CHECK_ARG(q && k && queue, "%s: null pointer", fn);
    GCA_CHECK_ARG(dtype_queue == GCA_F32 || dtype_queue == GCA_BF16, "%s: bad dtype_queue %d", fn, dtype_queue);
    GCA_CHECK_ARG(B >= 1 && K >= 1, "%s: need B >= 1 and K >= 1 (B=%d K=%lld)", fn, B, K);
    GCA_CHECK_ARG(inv_T > 0.f, "%s: inv_T must be > 0", fn);
    GCA_CHECK_ARG(algo == GCA_ALGO_AUTO || algo == GCA_ALGO_FFMA || algo == GCA_ALGO_TCGEN05 || algo == GCA_ALGO_TC32,
                  "%s: bad algo %d", fn, algo);
    const int a = pick_algo(algo, dtype_queue, d);
    if (a == GCA_ALGO_TCGEN05) {
        if (dtype_queue != GCA_BF16 || d != 128)
            return set_err(GCA_ERR_UNSUPPORTED, "%s: GCA_ALGO_TCGEN05 needs a bf16 queue and d == 128", fn);
    } else if (a == GCA_ALGO_TC32) {
        if (dtype_queue != GCA_F32 || d != 128)
            return set_err(GCA_ERR_UNSUPPORTED, "%s: GCA_ALGO_TC32 needs an fp32 queue and d == 128", fn);
    } else {
        if (d % 32 != 0 || d < 32 || d > 1024)
            return set_err(GCA_ERR_UNSUPPORTED, "%s: GCA_ALGO_FFMA needs d %% 32 == 0 and 32 <= d <= 1024 (d=%d)", fn, d);
    }
    return GCA_OK;
}

static int nsplit_for(int algo, int B, long long K, int d)
{
    if (algo == GCA_ALGO_TC32) return infonce_tc32_nsplit(B, K);
    return algo == GCA_ALGO_TCGEN05 ? infonce_tc_nsplit(B, K) : infonce_ffma_nsplit(B, K, d);
}

// launch the stream kernel of the chosen family; fills `ws`
static int run_stream(const float* q, const float* k, const void* queue, int dtype_queue, int B, long long K, int d,
                      float inv_T, int algo, const float* lse_fixed, bool want_acc, float* pos_out, float* logits_out,
                      void* workspace, size_t workspace_bytes, InfoNceWs* ws_out, cudaStream_t st, bool skip_prep = false,
                      const PeerXchg* px = nullptr, float* proj_k_hat = nullptr,
                      bool proj = false, int rank_cap = 0, int gather_Bl = 0)
{
    const int a = pick_algo(algo, dtype_queue, d);
    if (sm_count_cached() < 1) return set_err(GCA_ERR_CUDA, "no CUDA device available (this library has no CPU path)");
    const int nsplit = nsplit_for(a, B, K, d);
    InfoNceWs ws = infonce_ws_carve(workspace, B, d, nsplit);
    const size_t common = align_up(infonce_ws_carve(nullptr, B, d, infonce_max_splits(B)).bytes, 1024);
    const size_t need = (a == GCA_ALGO_TC32) ? common + infonce_tc32_extra_ws(B, K) : ws.bytes;
    if (!workspace || workspace_bytes < need)
        return set_err(GCA_ERR_WORKSPACE, "InfoNCE workspace too small: %zu bytes given, %zu needed", workspace_bytes, need);
    InfoNceStreamParams P{};
    P.q = q; P.k = k; P.queue = queue; P.B = B; P.K = K; P.d = d; P.inv_T = inv_T; P.lse_fixed = lse_fixed;
    P.counter = ws.counter; P.part_max = ws.part_max; P.part_sum = ws.part_sum; P.part_cnt = ws.part_cnt;
    P.part_acc = want_acc ? ws.part_acc : nullptr;
    P.nsplit = nsplit; P.Bpad = ws.Bpad; P.pos_out = pos_out ? pos_out : ws.pos_tmp; P.logits_out = logits_out; P.ld_logits = K + 1;
    P.q_bf16_ws = ws.q_bf16; P.pos_ws = ws.pos_ws; P.T_ = 1.f / inv_T; P.skip_prep = skip_prep ? 1 : 0;
    P.k_hat = ws.k_hat; P.inv_nq = ws.inv_nq;           // tcgen05 family: the prep kernel stages k (ffma family: unused)
    P.rank_cap = rank_cap; P.q_scale = 1.f;
    if (proj) {
        if (a != GCA_ALGO_TCGEN05)
            return set_err(GCA_ERR_UNSUPPORTED, "projection-tail fusion exists for the tcgen05 family only (bf16 queue, d == 128)");
        if (proj_k_hat) P.k_hat = proj_k_hat;
        P.normalize = 1;
    }
    if (px) {
        if (a != GCA_ALGO_TCGEN05)
            return set_err(GCA_ERR_UNSUPPORTED, "the peer-fused step exists for the tcgen05 family only (bf16 queue, d == 128)");
        P.xchg = *px;
        P.gather_Bl = gather_Bl;
    }
    *ws_out = ws;
    count_launch(1);
    if (a == GCA_ALGO_TC32) return infonce_tc32_launch(P, (char*)workspace + common, st);
    if (a == GCA_ALGO_TCGEN05) return infonce_tc_launch(P, lse_fixed != nullptr, st);
    return infonce_ffma_launch(P, dtype_queue, lse_fixed != nullptr, st);
}

}  // namespace gca

extern "C" int gca_version(void) { return GCA_ABI_VERSION; }
extern "C" const char* gca_last_error(void) { return gca::err_buf(); }
extern "C" long long gca_launch_count(void) { return gca::g_launches.load(); }
extern "C" int gca_sm_count(void)
{
    const int n = gca::sm_count_cached();
    return n > 0 ? n : gca::set_err(GCA_ERR_CUDA, "no CUDA device available");
}

// ---- launch plans (launch_plan.cuh) ----------------------------------------------------------------------------------
struct gca_plan { gca::LaunchPlan plan; };
static thread_local gca_plan* t_plan_owner = nullptr;      // the object whose .plan this thread records into

extern "C" int gca_plan_begin(void)
{
    using namespace gca;
    if (t_plan) return set_err(GCA_ERR_BAD_ARG, "gca_plan_begin: this thread is already recording a plan");
    if (sm_count_cached() < 1) return set_err(GCA_ERR_CUDA, "no CUDA device available (this library has no CPU path)");
    gca_plan* p = new gca_plan();
    cudaGetDevice(&p->plan.device);
    t_plan_owner = p;
    t_plan = &p->plan;
    return GCA_OK;
}

extern "C" int gca_plan_end(gca_plan** out)
{
    using namespace gca;
    if (!t_plan) return set_err(GCA_ERR_BAD_ARG, "gca_plan_end: this thread is not recording");
    gca_plan* p = t_plan_owner;
    t_plan = nullptr;
    t_plan_owner = nullptr;
    if (out) *out = nullptr;
    if (!out) { delete p; return set_err(GCA_ERR_BAD_ARG, "gca_plan_end: null output pointer"); }
    if (p->plan.recorded == 0 || p->plan.recorded != p->plan.counted) {
        const int rec = p->plan.recorded, cnt = p->plan.counted;
        delete p;
        return set_err(GCA_ERR_UNSUPPORTED, "gca_plan_end: %d launches recorded, %d issued -- only the tcgen05-family step entry "
                       "points (gca_moco_step, gca_moco_step_proj, gca_moco_step_peer on a bf16 queue with d == 128) are recordable", rec, cnt);
    }
    *out = p;
    return GCA_OK;
}

extern "C" int gca_plan_run(gca_plan* p, void* stream)
{
    using namespace gca;
    GCA_CHECK_ARG(p != nullptr, "gca_plan_run: null plan");
    if (t_plan) return set_err(GCA_ERR_BAD_ARG, "gca_plan_run: called while recording");
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev != p->plan.device)
        return set_err(GCA_ERR_BAD_ARG, "gca_plan_run: the plan was recorded on device %d, the current device is %d", p->plan.device, dev);
    cudaStream_t st = (cudaStream_t)stream, side = nullptr;
    for (PlanOp& op : p->plan.ops) {
        switch (op.kind) {
        case PLAN_LAUNCH: {
            cudaLaunchConfig_t cfg = op.cfg;
            cfg.attrs = op.attrs; cfg.numAttrs = (unsigned)op.nattrs;
            cfg.stream = op.lane ? side : st;
            if (op.lane && side == nullptr) return set_err(GCA_ERR_BAD_ARG, "gca_plan_run: side-stream launch without a fork");
            GCA_CUDA(cudaLaunchKernelExC(&cfg, op.func, op.args.data()));
            break;
        }
        case PLAN_FORK: { const int rc = side_stream_fork(st, &side); if (rc != GCA_OK) return rc; break; }
        case PLAN_JOIN: { const int rc = side_stream_join(st); if (rc != GCA_OK) return rc; break; }
        case PLAN_WAIT_EVENT: GCA_CUDA(cudaStreamWaitEvent(st, op.event, 0)); break;
        default: return set_err(GCA_ERR_BAD_ARG, "gca_plan_run: corrupt plan");
        }
    }
    count_launch(p->plan.recorded);
    return GCA_OK;
}

extern "C" int gca_plan_launches(const gca_plan* p) { return p ? p->plan.recorded : 0; }

extern "C" void gca_plan_destroy(gca_plan* p)
{
    if (p && gca::t_plan == &p->plan) { gca::t_plan = nullptr; t_plan_owner = nullptr; }
    delete p;
}

extern "C" int gca_workspace_set_done_flag(void* workspace, unsigned int* host_word, void* stream)
{
    using namespace gca;
    GCA_CHECK_ARG(workspace != nullptr, "gca_workspace_set_done_flag: null workspace");
    // control words 12..13 of the workspace (gca_common.cuh): the pointer the finalize launch publishes the step number to
    static thread_local unsigned long long staged[16];
    static thread_local int slot = 0;
    unsigned long long* src = &staged[slot++ & 15];          // (the async copy reads it later: one slot per recent call)
    *src = (unsigned long long)(uintptr_t)host_word;
    GCA_CUDA(cudaMemcpyAsync((char*)workspace + 12 * sizeof(unsigned int), src, sizeof(unsigned long long), cudaMemcpyHostToDevice,
                             (cudaStream_t)stream));
    GCA_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return GCA_OK;
}

extern "C" size_t gca_infonce_workspace_bytes(int B, long long K, int d, int dtype_queue, int algo)
{
    using namespace gca;
    (void)dtype_queue;
    if (B < 1 || d < 1) return 0;
    // sized for the largest split count any family may choose on this device (so one allocation serves all calls)
    const size_t common = infonce_ws_carve(nullptr, B, d, infonce_max_splits(B)).bytes;
    if (pick_algo(algo, dtype_queue, d) == GCA_ALGO_TC32 && K >= 1)
        return align_up(common, 1024) + infonce_tc32_extra_ws(B, K);                                // + queue / q planes
    return common;
}

static int infonce_fwd_impl(const char* fn, const float* q, const float* k, const void* queue, int dtype_queue, int B,
                            long long K, int d, float inv_T, int algo, float* loss_mean, float* loss_rows, float* lse,
                            float* pos_logit, int* rank_gt, int* top_hits, float* dq_unit, float* logits_out,
                            const float* enq_keys, int enq_N, long long enq_index, long long* enq_state, void* keys_ready_event,
                            void* workspace, size_t workspace_bytes, void* stream, const gca::PeerXchg* px = nullptr,
                            bool proj = false, float* k_hat_out = nullptr)
{
    using namespace gca;
    int rc = check_infonce_args(fn, q, k, queue, dtype_queue, B, K, d, inv_T, algo);
    if (rc != GCA_OK) return rc;
    GCA_CHECK_ARG(loss_rows && lse && pos_logit, "%s: loss_rows, lse, pos_logit are required", fn);
    GCA_CHECK_ARG(rank_gt || top_hits, "%s: rank_gt == NULL (top-k hits only) needs top_hits", fn);
    if (enq_keys || (proj && enq_N > 0)) {
        GCA_CHECK_ARG(enq_N >= 0 && enq_N <= K, "%s: N=%d rows do not fit a ring of %lld slots", fn, enq_N, K);
        GCA_CHECK_ARG(enq_state || (enq_index >= 0 && enq_index < K), "%s: pointer %lld outside [0, %lld)", fn, enq_index, K);
    }
    if (algo == GCA_ALGO_AUTO && logits_out && pick_algo(algo, dtype_queue, d) == GCA_ALGO_TC32)
        algo = GCA_ALGO_FFMA;                            // materialised logits: the CUDA-core kernel writes them
    cudaStream_t st = (cudaStream_t)stream;
    InfoNceWs ws;
    FinalizeParams F{};
    F.B = B; F.d = d; F.inv_T = inv_T; F.k = k; F.pos = pos_logit;
    F.lse = lse; F.loss_rows = loss_rows; F.rank_gt = rank_gt; F.dq = dq_unit; F.loss_mean = loss_mean; F.top_hits = top_hits;
    if (enq_keys || px || (proj && enq_N > 0)) {
        F.enq_queue = const_cast<void*>(queue); F.enq_dtype = dtype_queue; F.enq_K = K; F.enq_keys = enq_keys; F.enq_N = enq_N;
        F.enq_index = enq_index; F.enq_state = enq_state;
        if (px) F.xchg = *px;
    }
    rc = run_stream(q, k, queue, dtype_queue, B, K, d, inv_T, algo, nullptr, dq_unit != nullptr, pos_logit, logits_out,
                    workspace, workspace_bytes, &ws, st, false, px, k_hat_out, proj, rank_gt ? 0 : GCA_TOPK_RANK_CAP);
    if (rc != GCA_OK) return rc;
    if (proj) {                                          // the finalize kernel works on the normalised keys and maps dq to dzq
        const float* kh = k_hat_out ? k_hat_out : ws.k_hat;
        F.k = kh; F.zq = q; F.inv_nq = ws.inv_nq;
        if (F.enq_queue && F.enq_keys == nullptr) F.enq_keys = kh;
    } else if (pick_algo(algo, dtype_queue, d) == GCA_ALGO_TCGEN05 || pick_algo(algo, dtype_queue, d) == GCA_ALGO_TC32) {
        F.k = ws.k_hat;                                  // staged by the prep kernel: k is read once per step
    }
    F.counter = ws.counter; F.part_max = ws.part_max; F.part_sum = ws.part_sum; F.part_cnt = ws.part_cnt;
    F.part_acc = dq_unit ? ws.part_acc : nullptr;
    F.nsplit = ws.nsplit; F.Bpad = ws.Bpad;
    F.range_checked = ((pick_algo(algo, dtype_queue, d) == GCA_ALGO_TCGEN05 || pick_algo(algo, dtype_queue, d) == GCA_ALGO_TC32) &&
                       logits_out == nullptr) ? 1 : 0;
    if (keys_ready_event) {
        if (plan_recording()) {
            PlanOp op; op.kind = PLAN_WAIT_EVENT; op.event = (cudaEvent_t)keys_ready_event;
            plan_recording()->ops.push_back(std::move(op));
        } else {
            GCA_CUDA(cudaStreamWaitEvent(st, (cudaEvent_t)keys_ready_event, 0));
        }
    }
    rc = infonce_finalize_launch(F, FIN_FULL, st);
    if (rc == GCA_OK && px && pick_algo(algo, dtype_queue, d) == GCA_ALGO_TCGEN05) rc = keys_push_join(st);   // the side-stream key push
    return rc;
}

extern "C" int gca_infonce_fwd(const float* q, const float* k, const void* queue, int dtype_queue, int B, long long K,
                               int d, float inv_T, int algo, float* loss_mean, float* loss_rows, float* lse,
                               float* pos_logit, int* rank_gt, int* top_hits, float* dq_unit, float* logits_out, void* workspace,
                               size_t workspace_bytes, void* stream)
{
    return infonce_fwd_impl("gca_infonce_fwd", q, k, queue, dtype_queue, B, K, d, inv_T, algo, loss_mean, loss_rows, lse,
                            pos_logit, rank_gt, top_hits, dq_unit, logits_out, nullptr, 0, 0, nullptr, nullptr, workspace,
                            workspace_bytes, stream);
}

extern "C" int gca_moco_step(const float* q, const float* k, void* queue, int dtype_queue, int B, long long K, int d,
                             float inv_T, int algo, const float* enqueue_keys, int N, long long index, long long* state,
                             void* keys_ready_event, float* loss_mean, float* loss_rows, float* lse, float* pos_logit, int* rank_gt, int* top_hits,
                             float* dq_unit, void* workspace, size_t workspace_bytes, void* stream)
{
    GCA_CHECK_ARG(enqueue_keys, "gca_moco_step: enqueue_keys is required");
    return infonce_fwd_impl("gca_moco_step", q, k, queue, dtype_queue, B, K, d, inv_T, algo, loss_mean, loss_rows, lse,
                            pos_logit, rank_gt, top_hits, dq_unit, nullptr, enqueue_keys, N, index, state, keys_ready_event,
                            workspace, workspace_bytes, stream);
}

extern "C" int gca_moco_step_peer(const float* q, const float* k, void* queue, int dtype_queue, int B, long long K, int d,
                                  float inv_T, int algo, int W, int rank, void* const* mailboxes, long long* xstate,
                                  int timeout_ms, long long* state, float* loss_mean, float* loss_rows, float* lse,
                                  float* pos_logit, int* rank_gt, int* top_hits, float* dq_unit, void* workspace,
                                  size_t workspace_bytes, void* stream)
{
    using namespace gca;
    GCA_CHECK_ARG(mailboxes && xstate && state, "gca_moco_step_peer: mailboxes, xstate and state are required");
    GCA_CHECK_ARG(W >= 1 && W <= 64 && rank >= 0 && rank < W, "gca_moco_step_peer: bad rank %d of %d", rank, W);
    GCA_CHECK_ARG(B >= 1 && (long long)W * B <= K, "gca_moco_step_peer: %d x %d rows do not fit a ring of %lld slots", W, B, K);
    PeerXchg X{};
    X.mailboxes = (char* const*)mailboxes; X.W = W; X.rank = rank; X.n4 = B * d / 4;
    X.xstate = (unsigned long long*)xstate;
    X.timeout_ns = timeout_ms > 0 ? (unsigned long long)timeout_ms * 1000000ull : 0ull;
    return infonce_fwd_impl("gca_moco_step_peer", q, k, queue, dtype_queue, B, K, d, inv_T, algo, loss_mean, loss_rows, lse,
                            pos_logit, rank_gt, top_hits, dq_unit, nullptr, nullptr, W * B, 0, state, nullptr,
                            workspace, workspace_bytes, stream, &X);
}

extern "C" int gca_moco_step_proj(const float* zq, const float* zk, void* queue, int dtype_queue, int B, long long K, int d,
                                  float inv_T, int algo, const float* enqueue_keys, int N, long long index, long long* state,
                                  float* loss_mean, float* loss_rows, float* lse, float* pos_logit, int* rank_gt,
                                  int* top_hits, float* dz_unit, float* k_hat_out, void* workspace, size_t workspace_bytes,
                                  void* stream)
{
    using namespace gca;
    GCA_CHECK_ARG(N >= 0, "gca_moco_step_proj: N < 0");
    GCA_CHECK_ARG(enqueue_keys != nullptr || N == 0 || N == B,
                  "gca_moco_step_proj: without enqueue_keys the step enqueues its own B normalised keys (N must be 0 or B)");
    return infonce_fwd_impl("gca_moco_step_proj", zq, zk, queue, dtype_queue, B, K, d, inv_T, algo, loss_mean, loss_rows, lse,
                            pos_logit, rank_gt, top_hits, dz_unit, nullptr, enqueue_keys, N, index, state, nullptr,
                            workspace, workspace_bytes, stream, nullptr, true, k_hat_out);
}

extern "C" int gca_infonce_partials(const float* q, const float* k, const void* queue, int dtype_queue, int B,
                                    long long K, int d, float inv_T, int algo, int want_acc, void* workspace,
                                    size_t workspace_bytes, void* stream)
{
    using namespace gca;
    int rc = check_infonce_args("gca_infonce_partials", q, k, queue, dtype_queue, B, K, d, inv_T, algo);
    if (rc != GCA_OK) return rc;
    InfoNceWs ws;
    return run_stream(q, k, queue, dtype_queue, B, K, d, inv_T, algo, nullptr, (want_acc & 1) != 0, nullptr, nullptr, workspace,
                      workspace_bytes, &ws, (cudaStream_t)stream, (want_acc & 2) != 0);
}

extern "C" int gca_infonce_bwd(const float* q, const float* k, const void* queue, int dtype_queue, int B, long long K,
                               int d, float inv_T, int algo, const float* lse, float grad_scale, float* dq,
                               void* workspace, size_t workspace_bytes, void* stream)
{
    using namespace gca;
    int rc = check_infonce_args("gca_infonce_bwd", q, k, queue, dtype_queue, B, K, d, inv_T, algo);
    if (rc != GCA_OK) return rc;
    GCA_CHECK_ARG(lse && dq, "gca_infonce_bwd: lse and dq are required");
    if (algo == GCA_ALGO_AUTO && pick_algo(algo, dtype_queue, d) == GCA_ALGO_TC32) algo = GCA_ALGO_FFMA;   // (no two-pass variant)
    cudaStream_t st = (cudaStream_t)stream;
    InfoNceWs ws;
    // the stream kernel recomputes the positive logits into ws.pos_tmp
    rc = run_stream(q, k, queue, dtype_queue, B, K, d, inv_T, algo, lse, true, nullptr, nullptr, workspace,
                    workspace_bytes, &ws, st);
    if (rc != GCA_OK) return rc;
    FinalizeParams F{};
    F.counter = ws.counter; F.part_max = ws.part_max; F.part_sum = ws.part_sum; F.part_cnt = ws.part_cnt;
    F.part_acc = ws.part_acc; F.nsplit = ws.nsplit; F.Bpad = ws.Bpad; F.B = B; F.d = d; F.inv_T = inv_T; F.k = k;
    F.pos = ws.pos_tmp; F.lse_in = lse; F.grad_scale = grad_scale; F.dq = dq;
    return infonce_finalize_launch(F, FIN_BWD, st);
}

extern "C" int gca_infonce_shard_fwd(const float* q, const float* k, const void* shard, int dtype_queue, int B,
                                     long long K_shard, int d, float inv_T, int algo, float* pos_logit, float* part_max,
                                     float* part_sum, int* part_cnt, float* part_acc, void* workspace,
                                     size_t workspace_bytes, void* stream)
{
    using namespace gca;
    int rc = check_infonce_args("gca_infonce_shard_fwd", q, k, shard, dtype_queue, B, K_shard, d, inv_T, algo);
    if (rc != GCA_OK) return rc;
    GCA_CHECK_ARG(pos_logit && part_max && part_sum && part_cnt, "gca_infonce_shard_fwd: null output pointer");
    cudaStream_t st = (cudaStream_t)stream;
    InfoNceWs ws;
    rc = run_stream(q, k, shard, dtype_queue, B, K_shard, d, inv_T, algo, nullptr, part_acc != nullptr, pos_logit, nullptr,
                    workspace, workspace_bytes, &ws, st);
    if (rc != GCA_OK) return rc;
    FinalizeParams F{};
    F.counter = ws.counter; F.part_max = ws.part_max; F.part_sum = ws.part_sum; F.part_cnt = ws.part_cnt;
    F.part_acc = part_acc ? ws.part_acc : nullptr;
    F.nsplit = ws.nsplit; F.Bpad = ws.Bpad; F.B = B; F.d = d; F.inv_T = inv_T; F.k = k; F.pos = pos_logit;
    F.out_max = part_max; F.out_sum = part_sum; F.out_cnt = part_cnt; F.out_acc = part_acc;
    return infonce_finalize_launch(F, FIN_SHARD, st);
}

extern "C" size_t gca_shard_peer_bytes(int B_loc, int d, int W)
{
    if (B_loc <= 0 || d <= 0 || W <= 0) return 0;
    const size_t gather = gca_keys_exchange_bytes(2 * B_loc, d, W);
    const size_t rows = (size_t)2 * W * B_loc * (size_t)(d + 4) * sizeof(float);       // [2 parities][Bg * (d + 4)]
    const size_t flags = (size_t)2 * W * sizeof(unsigned long long);
    return gather + gca::align_up(rows + flags, 256);
}

extern "C" int gca_shard_step_peer(const float* qk_loc, void* shard, int dtype_queue, int B_loc, long long K, int d,
                                   float inv_T, int algo, int W, int rank, void* const* mailboxes, long long* pstate,
                                   int timeout_ms, long long* enq_state, float* qk_all, float* loss_mean, float* loss_rows,
                                   float* lse, float* pos_logit_all, int* rank_gt, int* top_hits, float* dq_unit,
                                   void* workspace, size_t workspace_bytes, void* stream)
{
    using namespace gca;
    GCA_CHECK_ARG(qk_loc && mailboxes && pstate && enq_state && qk_all, "gca_shard_step_peer: null pointer");
    GCA_CHECK_ARG(loss_mean && loss_rows && lse && pos_logit_all && rank_gt, "gca_shard_step_peer: null output pointer");
    GCA_CHECK_ARG(W >= 1 && W <= 64 && rank >= 0 && rank < W, "gca_shard_step_peer: bad rank %d of %d", rank, W);
    GCA_CHECK_ARG(B_loc >= 1 && K >= 1 && K % W == 0, "gca_shard_step_peer: K=%lld must divide evenly over %d ranks", K, W);
    GCA_CHECK_ARG((long long)W * B_loc <= K, "gca_shard_step_peer: %d x %d rows do not fit a ring of %lld slots", W, B_loc, K);
    const int Bg = W * B_loc;
    const long long Ks = K / W;
    const bool tc = pick_algo(algo, dtype_queue, d) == GCA_ALGO_TCGEN05;
    GCA_CHECK_ARG(tc || qk_all, "gca_shard_step_peer: qk_all is required outside the tcgen05 family");
    const float* q_all = tc ? qk_loc : qk_all;                          // (tcgen05 family: only the pointer check sees these)
    const float* k_all = tc ? qk_loc : qk_all + (size_t)Bg * d;
    int rc = check_infonce_args("gca_shard_step_peer", q_all, k_all, shard, dtype_queue, Bg, Ks, d, inv_T, algo);
    if (rc != GCA_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    PeerXchg X{};
    X.mailboxes = (char* const*)mailboxes; X.W = W; X.rank = rank; X.n4 = 2 * B_loc * d / 4;
    X.xstate = (unsigned long long*)pstate;
    X.timeout_ns = timeout_ms > 0 ? (unsigned long long)timeout_ms * 1000000ull : 0ull;
    // 1. q|k of every rank.  tcgen05 family: rides in the prep launch (push CTAs + row warps reading the mailbox);
    //    otherwise the stand-alone exchange kernel, output part-major (all q rows, then all k rows, rank-major inside)
    if (!tc) {
        rc = keys_exchange_launch(qk_loc, 2 * B_loc, d, W, rank, mailboxes, qk_all, pstate, timeout_ms, 2, st);
        if (rc != GCA_OK) return rc;
    }
    // 2. all rows against the local shard
    InfoNceWs ws;
    rc = run_stream(q_all, k_all, shard, dtype_queue, Bg, Ks, d, inv_T, algo, nullptr, dq_unit != nullptr, pos_logit_all,
                    nullptr, workspace, workspace_bytes, &ws, st, false, tc ? &X : nullptr, nullptr, false, 0, tc ? B_loc : 0);
    if (rc != GCA_OK) return rc;
    PeerMerge M{};
    M.mailboxes = (char* const*)mailboxes; M.off = gca_keys_exchange_bytes(2 * B_loc, d, W);
    M.W = W; M.rank = rank; M.Bl = B_loc; M.d = d;
    M.mstate = (unsigned long long*)pstate + 4;
    M.gather_state = tc ? (unsigned long long*)pstate : nullptr;
    M.timeout_ns = X.timeout_ns;
    // 3. split merge; every row's partial goes to its owner's mailbox
    FinalizeParams F{};
    F.counter = ws.counter; F.part_max = ws.part_max; F.part_sum = ws.part_sum; F.part_cnt = ws.part_cnt;
    F.part_acc = dq_unit ? ws.part_acc : nullptr;
    F.nsplit = ws.nsplit; F.Bpad = ws.Bpad; F.B = Bg; F.d = d; F.inv_T = inv_T; F.k = k_all; F.pos = pos_logit_all;
    F.merge = M;
    rc = infonce_finalize_launch(F, FIN_SHARD, st);
    if (rc != GCA_OK) return rc;
    // 4. cross-rank merge of the local rows; extra CTAs of the same launch enqueue the gathered keys that land in this
    //    rank's slots (tcgen05 family: the keys staged by the prep kernel)
    const float* keys = tc ? ws.k_hat : k_all;
    FinalizeParams G{};
    G.counter = ws.counter; G.nsplit = W; G.Bpad = ws.Bpad; G.B = B_loc; G.d = d; G.inv_T = inv_T;
    G.k = keys + (size_t)rank * B_loc * d;
    G.pos = pos_logit_all + (size_t)rank * B_loc;
    G.lse = lse; G.loss_rows = loss_rows; G.rank_gt = rank_gt; G.dq = dq_unit; G.loss_mean = loss_mean; G.top_hits = top_hits;
    G.range_checked = tc ? 1 : 0;
    G.merge = M; G.merge.wait = 1;
    G.enq_queue = shard; G.enq_dtype = dtype_queue; G.enq_K = K; G.enq_keys = keys; G.enq_N = Bg; G.enq_state = enq_state;
    G.enq_kbegin = (long long)rank * Ks; G.enq_kend = (long long)(rank + 1) * Ks;
    return infonce_finalize_launch(G, FIN_FULL, st);
}
