// Launch plans: a recorded sequence of this library's own kernel launches that can be re-issued with one call.
//
// A head step is 3-4 launches whose parameters do not change between steps (static buffers, device-resident ring
// pointer).  Re-issuing them from a recorded plan costs three cudaLaunchKernelExC calls on the host and keeps the
// programmatic dependencies between the launches AND across consecutive steps (the prep launch of step s+1 is resident
// while the finalize launch of step s drains), which a CUDA-graph launch per step cannot: measured 21.3 us per step
// against 24.2 us for graph replays at the headline shape (profiles/r02_notes.md section 10).
//
// Recording is thread-local (gca_plan_begin / gca_plan_end, gca_api.cu).  While a thread records, launch_ex() stores
// the launch instead of issuing it, the side-stream fork / join and event waits become plan operations, and
// count_launch() counts into the plan; gca_plan_end() refuses a plan whose count disagrees with what was recorded (some
// launch site bypassed the recorder).
#pragma once
#include <cuda_runtime.h>
#include <memory>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

namespace gca {

enum PlanOpKind { PLAN_LAUNCH = 0, PLAN_FORK = 1, PLAN_JOIN = 2, PLAN_WAIT_EVENT = 3 };

struct PlanOp {
    int kind = PLAN_LAUNCH;
    int lane = 0;                              // 0: the stream given to gca_plan_run, 1: the library's side stream
    const void* func = nullptr;
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attrs[4];
    int nattrs = 0;
    std::shared_ptr<void> storage;             // by-value copy of the kernel arguments
    std::vector<void*> args;                   // one pointer into `storage` per kernel parameter
    cudaEvent_t event = nullptr;               // PLAN_WAIT_EVENT
};

struct LaunchPlan {
    std::vector<PlanOp> ops;
    int recorded = 0;                          // launches stored by launch_ex()
    int counted = 0;                           // launches announced through count_launch() while recording
    int device = -1;
};

LaunchPlan* plan_recording();                  // the plan the calling thread is recording into, or nullptr (gca_api.cu)
bool plan_is_side_stream(cudaStream_t st);     // `st` is the library's side stream of the current device

template <typename Tup, size_t... I>
inline void plan_arg_pointers(Tup& t, std::vector<void*>& out, std::index_sequence<I...>)
{
    (out.push_back(const_cast<void*>(static_cast<const void*>(&std::get<I>(t)))), ...);
}

// cudaLaunchKernelEx, or -- while the calling thread records a plan -- the same launch stored for gca_plan_run
template <typename... KArgs, typename... Args>
inline cudaError_t launch_ex(const cudaLaunchConfig_t* cfg, void (*kern)(KArgs...), Args&&... args)
{
    LaunchPlan* plan = plan_recording();
    if (plan == nullptr) return cudaLaunchKernelEx(cfg, kern, std::forward<Args>(args)...);
    using Tup = std::tuple<std::decay_t<KArgs>...>;
    auto tup = std::make_shared<Tup>(std::decay_t<KArgs>(std::forward<Args>(args))...);
    PlanOp op;
    op.kind = PLAN_LAUNCH;
    op.lane = plan_is_side_stream(cfg->stream) ? 1 : 0;
    op.func = reinterpret_cast<const void*>(kern);
    op.cfg = *cfg;
    if (cfg->numAttrs > 4) return cudaErrorInvalidValue;         // (no launch site of this library sets more than three)
    op.nattrs = (int)cfg->numAttrs;
    for (int i = 0; i < op.nattrs; ++i) op.attrs[i] = cfg->attrs[i];
    op.cfg.attrs = nullptr;                    // re-pointed at op.attrs when the plan runs (ops move inside the vector)
    op.cfg.stream = nullptr;
    plan_arg_pointers(*tup, op.args, std::index_sequence_for<KArgs...>{});
    op.storage = tup;
    plan->ops.push_back(std::move(op));
    plan->recorded += 1;
    return cudaSuccess;
}

}  // namespace gca
