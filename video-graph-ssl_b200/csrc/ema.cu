// Momentum (EMA) update of the key encoder as ONE multi-tensor launch:  p_ema = m * p_ema + (1 - m) * p  for every parameter.
// Replaces Trainer._momentum_update (tools/train_video_contrast_dis.py:176-180, called every step at :440), which issues a
// mul_ and an add_ per parameter tensor (2 x ~60-300 launches).  HBM-bound: 8 bytes read + 4 written per element.
// The caller passes a device table of chunk descriptors (built once per model); a CTA handles one chunk with 128-bit
// accesses (scalar head/tail where a tensor is not 16-byte aligned).
#include "gca_common.cuh"

namespace gca {

struct EmaChunk { float* ema; const float* src; long long n; };     // n elements (<= chunk size) starting at both pointers

__global__ void __launch_bounds__(256)
ema_update_kernel(const EmaChunk* __restrict__ chunks, int nchunks, float m, float one_minus_m)
{
    for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const EmaChunk ch = chunks[c];
        float* __restrict__ e = ch.ema;
        const float* __restrict__ s = ch.src;
        const long long n = ch.n;
        const bool aligned = ((reinterpret_cast<uintptr_t>(e) | reinterpret_cast<uintptr_t>(s)) & 15) == 0;
        if (aligned) {
            const long long n4 = n >> 2;
            float4* e4 = reinterpret_cast<float4*>(e);
            const float4* s4 = reinterpret_cast<const float4*>(s);
            for (long long i = threadIdx.x; i < n4; i += 256) {
                float4 a = e4[i];
                const float4 b = __ldg(s4 + i);
                // same association as the reference: (ema * m) + (1 - m) * p   (mul_ then add_(p, alpha=1-m))
                a.x = fmaf(one_minus_m, b.x, a.x * m); a.y = fmaf(one_minus_m, b.y, a.y * m);
                a.z = fmaf(one_minus_m, b.z, a.z * m); a.w = fmaf(one_minus_m, b.w, a.w * m);
                e4[i] = a;
            }
            for (long long i = (n4 << 2) + threadIdx.x; i < n; i += 256) e[i] = fmaf(one_minus_m, __ldg(s + i), e[i] * m);
        } else {
            for (long long i = threadIdx.x; i < n; i += 256) e[i] = fmaf(one_minus_m, __ldg(s + i), e[i] * m);
        }
    }
}

}  // namespace gca

extern "C" size_t gca_ema_chunk_bytes(void) { return sizeof(gca::EmaChunk); }

extern "C" int gca_ema_update(const void* chunk_table, int nchunks, float momentum, float one_minus_momentum, void* stream)
{
    using namespace gca;
    GCA_CHECK_ARG(chunk_table || nchunks == 0, "gca_ema_update: null chunk table");
    GCA_CHECK_ARG(nchunks >= 0, "gca_ema_update: nchunks < 0");
    GCA_CHECK_ARG(momentum >= 0.f && momentum <= 1.f, "gca_ema_update: momentum outside [0, 1]");
    if (nchunks == 0) return GCA_OK;
    int sms = sm_count_cached();
    if (sms < 1) return set_err(GCA_ERR_CUDA, "no CUDA device available (this library has no CPU path)");
    int blocks = nchunks < sms * 8 ? nchunks : sms * 8;
    ema_update_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const EmaChunk*)chunk_table, nchunks, momentum, one_minus_momentum);
    GCA_LAUNCH_CHECK("ema_update_kernel");
    count_launch(1);
    return GCA_OK;
}
