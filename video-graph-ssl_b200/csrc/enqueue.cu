// K3: in-place vectorised MoCo enqueue (ring buffer with wrap-around, optional K-shard ownership).
// Replaces lib/memory/mem_moco.py:17-27 (arange + fmod + index_copy_: 4 launches) with one launch.
#include "gca_common.cuh"

namespace gca {

// one thread moves 4 consecutive features of one key row: a 128-bit load, a 128-bit (fp32) or 64-bit (bf16) store
// kDevPtr: the ring pointer lives in device memory (state[0]; state[1] is a self-resetting ticket) so that a captured
// CUDA graph can be replayed step after step; the last CTA to finish advances it by N (mem_moco.py:14-15).
template <typename QT, bool kDevPtr>
__global__ void __launch_bounds__(256)
enqueue_kernel(QT* __restrict__ queue, long long K_global, long long k_begin, long long k_end, int d4,
               const float4* __restrict__ keys, int N, long long index, long long* state)
{
    if (kDevPtr) index = *reinterpret_cast<volatile long long*>(state);
    const long long total = (long long)N * d4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(i / d4);
        const int c4  = (int)(i - (long long)row * d4);
        long long slot = index + row;                 // index < K and row < N <= K  =>  slot < 2K
        if (slot >= K_global) slot -= K_global;
        if (slot < k_begin || slot >= k_end) continue;  // not owned by this shard
        const float4 v = __ldg(keys + i);
        const long long off = (slot - k_begin) * (long long)d4 + c4;
        if constexpr (sizeof(QT) == 4) {
            reinterpret_cast<float4*>(queue)[off] = v;
        } else {
            __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);   // round-to-nearest-even, like Tensor.to(bfloat16)
            __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&lo);
            pk.y = *reinterpret_cast<uint32_t*>(&hi);
            reinterpret_cast<uint2*>(queue)[off] = pk;
        }
    }
    if (kDevPtr) {
        __shared__ int last;
        __syncthreads();                                   // every thread of this CTA has read the pointer
        if (threadIdx.x == 0) {
            const unsigned long long t = atomicAdd(reinterpret_cast<unsigned long long*>(state + 1), 1ull);
            last = (t == gridDim.x - 1);
        }
        __syncthreads();
        if (last && threadIdx.x == 0) {
            long long nx = index + N;
            if (nx >= K_global) nx -= K_global;
            state[0] = nx;
            state[1] = 0;
        }
    }
}

}  // namespace gca

static int enqueue_impl(const char* fn, void* queue, int dtype_queue, long long K_global, long long k_begin,
                        long long k_end, int d, const float* keys, int N, long long index, long long* state, void* stream)
{
    using namespace gca;
    GCA_CHECK_ARG(queue && keys, "%s: null pointer", fn);
    GCA_CHECK_ARG(dtype_queue == GCA_F32 || dtype_queue == GCA_BF16, "%s: bad dtype_queue %d", fn, dtype_queue);
    GCA_CHECK_ARG(K_global > 0 && d > 0 && d % 4 == 0, "%s: need K > 0 and d %% 4 == 0 (K=%lld d=%d)", fn, K_global, d);
    GCA_CHECK_ARG(0 <= k_begin && k_begin <= k_end && k_end <= K_global, "%s: bad shard range [%lld, %lld) of %lld", fn,
                  k_begin, k_end, K_global);
    GCA_CHECK_ARG(N >= 0 && N <= K_global, "%s: N=%d rows do not fit a ring of %lld slots", fn, N, K_global);
    GCA_CHECK_ARG(state || (index >= 0 && index < K_global), "%s: pointer %lld outside [0, %lld)", fn, index, K_global);
    if (N == 0 || (k_begin == k_end && !state)) return GCA_OK;
    const int d4 = d / 4;
    const long long total = (long long)N * d4;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    cudaStream_t st = (cudaStream_t)stream;
    const float4* k4 = (const float4*)keys;
    if (dtype_queue == GCA_F32) {
        if (state) enqueue_kernel<float, true><<<blocks, 256, 0, st>>>((float*)queue, K_global, k_begin, k_end, d4, k4, N, 0, state);
        else       enqueue_kernel<float, false><<<blocks, 256, 0, st>>>((float*)queue, K_global, k_begin, k_end, d4, k4, N, index, nullptr);
    } else {
        if (state) enqueue_kernel<__nv_bfloat16, true><<<blocks, 256, 0, st>>>((__nv_bfloat16*)queue, K_global, k_begin, k_end, d4, k4, N, 0, state);
        else       enqueue_kernel<__nv_bfloat16, false><<<blocks, 256, 0, st>>>((__nv_bfloat16*)queue, K_global, k_begin, k_end, d4, k4, N, index, nullptr);
    }
    GCA_LAUNCH_CHECK("enqueue_kernel");
    count_launch(1);
    return GCA_OK;
}

extern "C" int gca_enqueue(void* queue, int dtype_queue, long long K_global, long long k_begin, long long k_end,
                           int d, const float* keys, int N, long long index, void* stream)
{
    return enqueue_impl("gca_enqueue", queue, dtype_queue, K_global, k_begin, k_end, d, keys, N, index, nullptr, stream);
}

extern "C" int gca_enqueue_devptr(void* queue, int dtype_queue, long long K_global, long long k_begin, long long k_end,
                                  int d, const float* keys, int N, long long* state, void* stream)
{
    GCA_CHECK_ARG(state, "gca_enqueue_devptr: null state");
    return enqueue_impl("gca_enqueue_devptr", queue, dtype_queue, K_global, k_begin, k_end, d, keys, N, 0, state, stream);
}
