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

// ---------------------------------------------------------------------------------------------
// Workspace layout shared by the InfoNCE stream kernels and the finalize kernels.
//   [0, 256)                 : control block (block-completion counter, self-resetting)
//   part_max [nsplit, Bpad]  : natural-log max logit of the split        (-inf when the split is empty)
//   part_sum [nsplit, Bpad]  : sum exp(logit - part_max)
//   part_cnt [nsplit, Bpad]  : #negatives > positive in the split
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
    int    nsplit;
    int    Bpad;
    size_t bytes;
};

static inline int infonce_bpad(int B) { return (B + 127) / 128 * 128; }
// upper bound on the number of K-splits any kernel family uses (2 CTAs worth per SM, at least 1)
int infonce_max_splits(int B);
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

__device__ __forceinline__ float ld_queue(const float* p) { return *p; }
__device__ __forceinline__ float ld_queue(const __nv_bfloat16* p) { return __bfloat162float(*p); }

}  // namespace gca
