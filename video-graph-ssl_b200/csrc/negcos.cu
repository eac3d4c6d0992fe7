// K6: SimSiam negative-cosine loss, forward + unit gradient in one launch.
// Replaces D.forward 'v2' (lib/memory/criterion.py:53-62 == lib/modeling/graph_wrappers.py:99-108):
//   loss = -mean_b cos(p_b, stopgrad(z_b)),  cos as ATen's cosine_similarity: rows are divided by
//   max(||.||, eps) (eps = 1e-8) before the dot product.
#include "gca_common.cuh"

namespace gca {

constexpr int NC_THREADS = 128;          // 4 warps = 4 rows per CTA
constexpr float NC_EPS = 1e-8f;

__global__ void __launch_bounds__(NC_THREADS)
negcos_kernel(const float* __restrict__ p, const float* __restrict__ z, int B, int d,
              float* loss, float* cos_rows, float* dp_unit, unsigned int* counter)
{
    __shared__ float red[NC_THREADS / 32];
    __shared__ int flag;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = blockIdx.x * (NC_THREADS / 32) + warp;
    if (row < B) {
        const float* pr = p + (size_t)row * d;
        const float* zr = z + (size_t)row * d;
        float dot = 0.f, pp = 0.f, zz = 0.f;
        if (d % 4 == 0) {
            for (int c = lane * 4; c < d; c += 128) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(pr + c));
                const float4 b = __ldg(reinterpret_cast<const float4*>(zr + c));
                dot = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, dot))));
                pp  = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, fmaf(a.w, a.w, pp))));
                zz  = fmaf(b.x, b.x, fmaf(b.y, b.y, fmaf(b.z, b.z, fmaf(b.w, b.w, zz))));
            }
        } else {
            for (int c = lane; c < d; c += 32) {
                const float a = __ldg(pr + c), b = __ldg(zr + c);
                dot = fmaf(a, b, dot); pp = fmaf(a, a, pp); zz = fmaf(b, b, zz);
            }
        }
        dot = warp_sum(dot); pp = warp_sum(pp); zz = warp_sum(zz);
        const float pn = sqrtf(pp), zn = sqrtf(zz);
        const float ipn = 1.f / fmaxf(pn, NC_EPS), izn = 1.f / fmaxf(zn, NC_EPS);
        const float cosv = dot * ipn * izn;
        if (lane == 0) cos_rows[row] = cosv;
        if (dp_unit) {
            // d(-cos)/dp / B ;  d cos/dp = z^/max(|p|,eps) - [|p| >= eps] cos p / |p|^2
            const float gs = -1.f / (float)B;
            const float self = (pn >= NC_EPS) ? cosv / pp : 0.f;
            float* dr = dp_unit + (size_t)row * d;
            for (int c = lane; c < d; c += 32)
                dr[c] = gs * (__ldg(zr + c) * izn * ipn - self * __ldg(pr + c));
        }
    }
    // last CTA: loss = -mean(cos_rows) in a fixed order
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) flag = (atomicAdd(counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (flag) {
        __threadfence();
        float a = 0.f;
        for (int i = threadIdx.x; i < B; i += NC_THREADS) a += __ldcg(cos_rows + i);
        a = block_sum<NC_THREADS>(a, red);
        if (threadIdx.x == 0) { *loss = -a / (float)B; *counter = 0u; }
    }
}

}  // namespace gca

extern "C" size_t gca_negcos_workspace_bytes(int B, int d) { (void)d; return 256 + (size_t)(B > 0 ? B : 0) * sizeof(float); }

extern "C" int gca_negcos_fwd_bwd(const float* p, const float* z, int B, int d, float* loss, float* cos_rows,
                                  float* dp_unit, void* workspace, size_t workspace_bytes, void* stream)
{
    using namespace gca;
    GCA_CHECK_ARG(p && z && loss, "gca_negcos_fwd_bwd: null pointer");
    GCA_CHECK_ARG(B >= 1 && d >= 1, "gca_negcos_fwd_bwd: bad sizes B=%d d=%d", B, d);
    if (!workspace || workspace_bytes < gca_negcos_workspace_bytes(B, d))
        return set_err(GCA_ERR_WORKSPACE, "gca_negcos_fwd_bwd: workspace of %zu bytes needed", gca_negcos_workspace_bytes(B, d));
    cudaStream_t st = (cudaStream_t)stream;
    unsigned int* counter = (unsigned int*)workspace;
    float* rows = cos_rows ? cos_rows : (float*)((char*)workspace + 256);
    GCA_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
    const int blocks = (B + NC_THREADS / 32 - 1) / (NC_THREADS / 32);
    negcos_kernel<<<blocks, NC_THREADS, 0, st>>>(p, z, B, d, loss, rows, dp_unit, counter);
    GCA_LAUNCH_CHECK("negcos_kernel");
    count_launch(1);
    return GCA_OK;
}
