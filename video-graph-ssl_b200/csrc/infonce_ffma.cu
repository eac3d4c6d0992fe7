// InfoNCE stream, CUDA-core family (GCA_ALGO_FFMA): exact fp32 arithmetic for the parity mode, any
// d % 32 == 0 (d <= 1024), fp32 or bf16 queue, optional materialised logits.
//
// One CTA = (BM query rows) x (a contiguous range of 64-key tiles).  Per tile: S = q Q^T in registers
// (4x4 micro-tiles), online max / sum-exp per row with 16-lane shuffles, P -> smem, O += P Q with the
// running accumulator O[BM, d] resident in shared memory.  Split partials go to the workspace and are
// merged in fixed order by finalize.cu.  Replaces mem_moco.py:36-46 + criterion.py:44 + their autograd.
#include "gca_common.cuh"
#include "infonce_params.cuh"

namespace gca {

constexpr int FF_BN = 64;      // keys per tile
constexpr int FF_DK = 32;      // feature chunk for the S phase
constexpr int FF_DN = 64;      // feature chunk for the O phase
constexpr int FF_LDA = FF_DK + 1;
constexpr int FF_LDV = FF_DN + 4;

template <int BM>
constexpr size_t ffma_smem_bytes(int d, bool want_acc) {
    return sizeof(float) * ((size_t)BM * FF_LDA + (size_t)FF_BN * FF_LDA + (size_t)BM * FF_BN +
                            (size_t)FF_BN * FF_LDV + (want_acc ? (size_t)BM * d : 0) + BM);
}

template <typename QT, int BM, bool kFixedMax>
__global__ void __launch_bounds__(256)
infonce_ffma_kernel(const InfoNceStreamParams P)
{
    constexpr int MR = BM / 16;                       // rows per thread
    extern __shared__ __align__(16) float smem[];
    float* As  = smem;                                // [BM][FF_LDA]
    float* Bs  = As + BM * FF_LDA;                    // [FF_BN][FF_LDA]
    float* Ps  = Bs + FF_BN * FF_LDA;                 // [BM][FF_BN]
    float* Vs  = Ps + BM * FF_BN;                     // [FF_BN][FF_LDV]
    float* pos_s = Vs + FF_BN * FF_LDV;               // [BM]
    float* Os  = pos_s + BM;                          // [BM][d] (only when want_acc)

    const QT* __restrict__ queue = reinterpret_cast<const QT*>(P.queue);
    const float* __restrict__ q = P.q;
    const int d = P.d, B = P.B;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4;
    const int split = blockIdx.x, row0 = blockIdx.y * BM;
    const bool want_acc = (P.part_acc != nullptr);

    if (split == 0 && blockIdx.y == 0 && tid == 0) { for (int w = 0; w < 6; ++w) P.counter[w] = 0u; }     // re-arm the finalize control block

    // positive logit of each row, fp32, fixed summation order (identical in every split)
    for (int r = warp; r < BM; r += 8) {
        const int row = row0 + r;
        float a = 0.f;
        if (row < B)
            for (int c = lane; c < d; c += 32) a = fmaf(q[(size_t)row * d + c], P.k[(size_t)row * d + c], a);
        a = warp_sum(a) * P.inv_T;
        if (lane == 0) {
            pos_s[r] = a;
            if (split == 0 && row < B) {
                if (P.pos_out) P.pos_out[row] = a;
                if (P.logits_out) P.logits_out[(size_t)row * P.ld_logits] = a;      // column 0 = positive
            }
        }
    }
    if (want_acc) for (int i = tid; i < BM * d; i += 256) Os[i] = 0.f;
    __syncthreads();

    float pos_r[MR], m_run[MR], s_run[MR];
    int cnt[MR];
#pragma unroll
    for (int i = 0; i < MR; ++i) {
        pos_r[i] = pos_s[ty * MR + i];
        const int row = row0 + ty * MR + i;
        m_run[i] = kFixedMax ? ((row < B) ? P.lse_fixed[row] : 0.f) : -INFINITY;
        s_run[i] = 0.f; cnt[i] = 0;
    }

    const long long ntiles = (P.K + FF_BN - 1) / FF_BN;
    const long long t_begin = ntiles * split / P.nsplit, t_end = ntiles * (split + 1) / P.nsplit;

    for (long long t = t_begin; t < t_end; ++t) {
        const long long j0 = t * FF_BN;
        float acc[MR][4];
#pragma unroll
        for (int i = 0; i < MR; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;

        // ---------------- S phase: acc = q[rows] . Q[keys] over d, FF_DK features at a time
        for (int d0 = 0; d0 < d; d0 += FF_DK) {
            __syncthreads();
            for (int e = tid; e < BM * FF_DK; e += 256) {
                const int r = e / FF_DK, dd = e % FF_DK, row = row0 + r;
                As[r * FF_LDA + dd] = (row < B) ? q[(size_t)row * d + d0 + dd] : 0.f;
            }
            for (int e = tid; e < FF_BN * FF_DK; e += 256) {
                const int j = e / FF_DK, dd = e % FF_DK;
                const long long key = j0 + j;
                Bs[j * FF_LDA + dd] = (key < P.K) ? ld_queue(queue + (size_t)key * d + d0 + dd) : 0.f;
            }
            __syncthreads();
#pragma unroll 8
            for (int dd = 0; dd < FF_DK; ++dd) {
                float a[MR], b[4];
#pragma unroll
                for (int i = 0; i < MR; ++i) a[i] = As[(ty * MR + i) * FF_LDA + dd];
#pragma unroll
                for (int c = 0; c < 4; ++c) b[c] = Bs[(tx * 4 + c) * FF_LDA + dd];
#pragma unroll
                for (int i = 0; i < MR; ++i)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[i][c] = fmaf(a[i], b[c], acc[i][c]);
            }
        }

        // ---------------- logits of this tile, optional materialisation, mask, rank count
        float scale[MR];
#pragma unroll
        for (int i = 0; i < MR; ++i) {
            const int row = row0 + ty * MR + i;
            float tmax = -INFINITY;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const long long key = j0 + tx * 4 + c;
                float x = acc[i][c] * P.inv_T;
                if (key < P.K) {
                    if (P.logits_out && row < B) P.logits_out[(size_t)row * P.ld_logits + 1 + key] = x;
                    cnt[i] += (x > pos_r[i]) ? 1 : 0;
                } else {
                    x = -INFINITY;
                }
                acc[i][c] = x;
                tmax = fmaxf(tmax, x);
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
            float m_new;
            if (kFixedMax) { m_new = m_run[i]; scale[i] = 1.f; }
            else {
                m_new = fmaxf(m_run[i], tmax);
                scale[i] = (m_run[i] == -INFINITY) ? 0.f : __expf(m_run[i] - m_new);
            }
            float rs = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) { acc[i][c] = __expf(acc[i][c] - m_new); rs += acc[i][c]; }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
            s_run[i] = s_run[i] * scale[i] + rs;
            m_run[i] = m_new;
        }

        // ---------------- O phase: Os = Os * scale + P . Q[keys], FF_DN features at a time
        if (want_acc) {
#pragma unroll
            for (int i = 0; i < MR; ++i)
                *reinterpret_cast<float4*>(&Ps[(ty * MR + i) * FF_BN + tx * 4]) =
                    make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            for (int d0 = 0; d0 < d; d0 += FF_DN) {
                __syncthreads();                       // Ps visible / previous Vs consumed
                for (int e = tid; e < FF_BN * FF_DN; e += 256) {
                    const int j = e / FF_DN, c = e % FF_DN;
                    const long long key = j0 + j;
                    Vs[j * FF_LDV + c] = (key < P.K && d0 + c < d) ? ld_queue(queue + (size_t)key * d + d0 + c) : 0.f;
                }
                __syncthreads();
                float o[MR][4];
#pragma unroll
                for (int i = 0; i < MR; ++i)
#pragma unroll
                    for (int c = 0; c < 4; ++c) o[i][c] = 0.f;
#pragma unroll 8
                for (int j = 0; j < FF_BN; ++j) {
                    const float4 v = *reinterpret_cast<const float4*>(&Vs[j * FF_LDV + tx * 4]);
#pragma unroll
                    for (int i = 0; i < MR; ++i) {
                        const float p = Ps[(ty * MR + i) * FF_BN + j];
                        o[i][0] = fmaf(p, v.x, o[i][0]); o[i][1] = fmaf(p, v.y, o[i][1]);
                        o[i][2] = fmaf(p, v.z, o[i][2]); o[i][3] = fmaf(p, v.w, o[i][3]);
                    }
                }
                if (d0 + tx * 4 < d) {
#pragma unroll
                    for (int i = 0; i < MR; ++i) {
                        float4* dst = reinterpret_cast<float4*>(&Os[(size_t)(ty * MR + i) * d + d0 + tx * 4]);
                        float4 cur = *dst;
                        cur.x = fmaf(cur.x, scale[i], o[i][0]); cur.y = fmaf(cur.y, scale[i], o[i][1]);
                        cur.z = fmaf(cur.z, scale[i], o[i][2]); cur.w = fmaf(cur.w, scale[i], o[i][3]);
                        *dst = cur;
                    }
                }
            }
        }
    }

    // ---------------- split partials
#pragma unroll
    for (int i = 0; i < MR; ++i) {
        int c = cnt[i];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        const int r = row0 + ty * MR + i;
        if (tx == 0 && r < P.Bpad) {
            const size_t o = part_stat_index(split, r, P.nsplit);
            P.part_max[o] = m_run[i];
            P.part_sum[o] = s_run[i];
            P.part_cnt[o] = c;
        }
    }
    if (want_acc) {
        __syncthreads();
        for (int i = tid; i < BM * d; i += 256) {
            const int r = i / d, c = i - r * d, row = row0 + r;
            if (row < P.Bpad) P.part_acc[((size_t)split * P.Bpad + row) * d + c] = Os[i];
        }
    }
}

template <typename QT, int BM, bool kFixedMax>
static int launch_ffma(const InfoNceStreamParams& P, cudaStream_t st)
{
    const bool want_acc = P.part_acc != nullptr;
    const size_t smem = ffma_smem_bytes<BM>(P.d, want_acc);
    auto kern = infonce_ffma_kernel<QT, BM, kFixedMax>;
    GCA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(P.nsplit, (P.B + BM - 1) / BM);
    kern<<<grid, 256, smem, st>>>(P);
    GCA_LAUNCH_CHECK("infonce_ffma_kernel");
    return GCA_OK;
}

int infonce_ffma_pick_bm(int B, int d) { return (d > 512 || B <= 32) ? 32 : 64; }

int infonce_ffma_nsplit(int B, long long K, int d)
{
    const int bm = infonce_ffma_pick_bm(B, d);
    const int nblk = (B + bm - 1) / bm;
    long long ntiles = (K + FF_BN - 1) / FF_BN;
    int ns = sm_count_cached() / nblk;
    if (ns < 1) ns = 1;
    if (ns > ntiles) ns = (int)ntiles;
    return ns;
}

int infonce_ffma_launch(const InfoNceStreamParams& P, int dtype_queue, bool fixed_max, cudaStream_t st)
{
    const int bm = infonce_ffma_pick_bm(P.B, P.d);
#define GCA_FF(QT, BM) (fixed_max ? launch_ffma<QT, BM, true>(P, st) : launch_ffma<QT, BM, false>(P, st))
    if (dtype_queue == GCA_F32) return bm == 64 ? GCA_FF(float, 64) : GCA_FF(float, 32);
    return bm == 64 ? GCA_FF(__nv_bfloat16, 64) : GCA_FF(__nv_bfloat16, 32);
#undef GCA_FF
}

}  // namespace gca
