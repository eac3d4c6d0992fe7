// K11: BatchNorm1d (+ ReLU) epilogue of the SimSiam projection / prediction MLPs, forward and backward in one launch each.
// Replaces the nn.BatchNorm1d -> nn.ReLU(inplace) pairs of ProjectionMLP / PredictionMLP (lib/modeling/project_head.py:36-76;
// ATen: batch-norm statistics + apply + ReLU = 3 launches forward, 3-4 backward); the Linear in front stays a library GEMM.
//
// x is [B, C] row-major (the GEMM output).  One CTA owns 32 feature columns and all B rows: a warp reads one 128-byte row
// segment per load (fully coalesced), each thread keeps the values of its column in registers (up to BN_CACHE rows per thread,
// i.e. B <= 256; larger batches re-read, L2-resident), so x is read once and y written once: 8 * B * C bytes forward,
// 12 * B * C backward (x, dy in; dx out).  Statistics are two-pass (mean, then sum of squared deviations) with a fixed
// reduction order: rows in order inside a thread, warps in order across the CTA -- deterministic.
#include "gca_common.cuh"

namespace gca {

constexpr int BN_THREADS = 256;
constexpr int BN_WARPS = BN_THREADS / 32;
constexpr int BN_CACHE = 32;                     // rows cached per thread

// sum over the 8 warps of one column (lane), fixed order; `sm` is [BN_WARPS][32]
__device__ __forceinline__ float bn_col_sum(float v, float (*sm)[32], int warp, int lane)
{
    __syncthreads();                              // previous use of sm is over
    sm[warp][lane] = v;
    __syncthreads();
    float t = sm[0][lane];
#pragma unroll
    for (int w = 1; w < BN_WARPS; ++w) t += sm[w][lane];
    return t;
}

__global__ void __launch_bounds__(BN_THREADS)
bn1d_fwd_kernel(const float* __restrict__ x, int B, int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                float eps, float momentum, int relu, int use_running, float* running_mean, float* running_var,
                float* __restrict__ y, float* save_mean, float* save_invstd)
{
    __shared__ float sm[BN_WARPS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    const bool live = c < C;
    float v[BN_CACHE];
    const bool cached = B <= BN_WARPS * BN_CACHE;
    if (cached) {
#pragma unroll
        for (int i = 0; i < BN_CACHE; ++i) {
            const int r = warp + i * BN_WARPS;
            v[i] = (live && r < B) ? __ldg(x + (size_t)r * C + c) : 0.f;
        }
    }
    float mean, invstd;
    if (use_running) {                            // eval mode: the running statistics normalise, nothing is updated
        mean = live ? running_mean[c] : 0.f;
        invstd = live ? 1.f / sqrtf(running_var[c] + eps) : 0.f;
    } else {
        float s = 0.f;
        if (cached) {
#pragma unroll
            for (int i = 0; i < BN_CACHE; ++i) s += v[i];
        } else {
            for (int r = warp; r < B; r += BN_WARPS) s += live ? __ldg(x + (size_t)r * C + c) : 0.f;
        }
        mean = bn_col_sum(s, sm, warp, lane) / (float)B;
        float q = 0.f;
        if (cached) {
#pragma unroll
            for (int i = 0; i < BN_CACHE; ++i) {
                const int r = warp + i * BN_WARPS;
                const float dlt = v[i] - mean;
                if (r < B) q = fmaf(dlt, dlt, q);
            }
        } else {
            for (int r = warp; r < B; r += BN_WARPS) {
                const float dlt = (live ? __ldg(x + (size_t)r * C + c) : 0.f) - mean;
                q = fmaf(dlt, dlt, q);
            }
        }
        const float ssd = bn_col_sum(q, sm, warp, lane);
        const float var = ssd / (float)B;                          // biased: normalises the batch
        invstd = 1.f / sqrtf(var + eps);
        if (live && warp == 0) {
            if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
            if (running_var)  running_var[c]  = (1.f - momentum) * running_var[c] + momentum * (B > 1 ? ssd / (float)(B - 1) : var);
        }
    }
    if (live && warp == 0) {
        if (save_mean) save_mean[c] = mean;
        if (save_invstd) save_invstd[c] = invstd;
    }
    if (!live) return;
    const float g = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
    if (cached) {
#pragma unroll
        for (int i = 0; i < BN_CACHE; ++i) {
            const int r = warp + i * BN_WARPS;
            if (r < B) {
                float o = fmaf((v[i] - mean) * invstd, g, bt);
                if (relu) o = fmaxf(o, 0.f);
                y[(size_t)r * C + c] = o;
            }
        }
    } else {
        for (int r = warp; r < B; r += BN_WARPS) {
            float o = fmaf((__ldg(x + (size_t)r * C + c) - mean) * invstd, g, bt);
            if (relu) o = fmaxf(o, 0.f);
            y[(size_t)r * C + c] = o;
        }
    }
}

// dz = dy * [y > 0] (ReLU), dbeta = sum dz, dgamma = sum dz * xhat,
// training: dx = gamma * invstd * (dz - dbeta / B - xhat * dgamma / B);  eval: dx = gamma * invstd * dz
__global__ void __launch_bounds__(BN_THREADS)
bn1d_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, int B, int C, const float* __restrict__ gamma,
                const float* __restrict__ beta, const float* __restrict__ save_mean, const float* __restrict__ save_invstd,
                int relu, int use_running, float* __restrict__ dx, float* dgamma, float* dbeta)
{
    __shared__ float sm[BN_WARPS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    const bool live = c < C;
    const float mean = live ? save_mean[c] : 0.f, invstd = live ? save_invstd[c] : 0.f;
    const float g = (live && gamma) ? gamma[c] : 1.f, bt = (live && beta) ? beta[c] : 0.f;
    const bool cached = B <= BN_WARPS * BN_CACHE;
    float xh[BN_CACHE], dz[BN_CACHE];
    float sb = 0.f, sg = 0.f;
    if (cached) {
#pragma unroll
        for (int i = 0; i < BN_CACHE; ++i) {
            const int r = warp + i * BN_WARPS;
            const bool on = live && r < B;
            const float xv = on ? __ldg(x + (size_t)r * C + c) : 0.f;
            float d = on ? __ldg(dy + (size_t)r * C + c) : 0.f;
            xh[i] = (xv - mean) * invstd;
            if (relu && !(fmaf(xh[i], g, bt) > 0.f)) d = 0.f;
            dz[i] = d;
            sb += d;
            sg = fmaf(d, xh[i], sg);
        }
    } else {
        for (int r = warp; r < B; r += BN_WARPS) {
            const float xv = live ? __ldg(x + (size_t)r * C + c) : 0.f;
            float d = live ? __ldg(dy + (size_t)r * C + c) : 0.f;
            const float h = (xv - mean) * invstd;
            if (relu && !(fmaf(h, g, bt) > 0.f)) d = 0.f;
            sb += d;
            sg = fmaf(d, h, sg);
        }
    }
    const float db = bn_col_sum(sb, sm, warp, lane);
    const float dg = bn_col_sum(sg, sm, warp, lane);
    if (live && warp == 0) {
        if (dbeta) dbeta[c] = db;
        if (dgamma) dgamma[c] = dg;
    }
    if (!live || dx == nullptr) return;
    const float k1 = use_running ? 0.f : db / (float)B, k2 = use_running ? 0.f : dg / (float)B;
    const float gi = g * invstd;
    if (cached) {
#pragma unroll
        for (int i = 0; i < BN_CACHE; ++i) {
            const int r = warp + i * BN_WARPS;
            if (r < B) dx[(size_t)r * C + c] = gi * (dz[i] - k1 - xh[i] * k2);
        }
    } else {
        for (int r = warp; r < B; r += BN_WARPS) {
            const float h = (__ldg(x + (size_t)r * C + c) - mean) * invstd;
            float d = __ldg(dy + (size_t)r * C + c);
            if (relu && !(fmaf(h, g, bt) > 0.f)) d = 0.f;
            dx[(size_t)r * C + c] = gi * (d - k1 - h * k2);
        }
    }
}

}  // namespace gca

extern "C" int gca_bn1d_fwd(const float* x, int B, int C, const float* gamma, const float* beta, float eps, float momentum,
                            int relu, int use_running, float* running_mean, float* running_var, float* y, float* save_mean,
                            float* save_invstd, void* stream)
{
    using namespace gca;
    GCA_CHECK_ARG(x && y, "gca_bn1d_fwd: null pointer");
    GCA_CHECK_ARG(B >= 1 && C >= 1, "gca_bn1d_fwd: bad sizes B=%d C=%d", B, C);
    GCA_CHECK_ARG(eps >= 0.f, "gca_bn1d_fwd: eps < 0");
    GCA_CHECK_ARG(!use_running || (running_mean && running_var), "gca_bn1d_fwd: eval mode needs the running statistics");
    bn1d_fwd_kernel<<<(C + 31) / 32, BN_THREADS, 0, (cudaStream_t)stream>>>(x, B, C, gamma, beta, eps, momentum, relu, use_running,
                                                                          running_mean, running_var, y, save_mean, save_invstd);
    GCA_LAUNCH_CHECK("bn1d_fwd_kernel");
    count_launch(1);
    return GCA_OK;
}

extern "C" int gca_bn1d_bwd(const float* x, const float* dy, int B, int C, const float* gamma, const float* beta,
                            const float* save_mean, const float* save_invstd, int relu, int use_running, float* dx,
                            float* dgamma, float* dbeta, void* stream)
{
    using namespace gca;
    GCA_CHECK_ARG(x && dy && save_mean && save_invstd, "gca_bn1d_bwd: null pointer");
    GCA_CHECK_ARG(B >= 1 && C >= 1, "gca_bn1d_bwd: bad sizes B=%d C=%d", B, C);
    bn1d_bwd_kernel<<<(C + 31) / 32, BN_THREADS, 0, (cudaStream_t)stream>>>(x, dy, B, C, gamma, beta, save_mean, save_invstd, relu,
                                                                          use_running, dx, dgamma, dbeta);
    GCA_LAUNCH_CHECK("bn1d_bwd_kernel");
    count_launch(1);
    return GCA_OK;
}
