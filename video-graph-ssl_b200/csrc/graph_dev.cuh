// Device primitives and argument blocks of the temporal clip-graph head (shared by graph.cu / graph_bwd.cu).
#pragma once
#include "gca_common.cuh"

namespace gca {

constexpr int G_THREADS = 256;
constexpr int G_TMAXMAX = 32;
constexpr int G_CHUNK_FLOATS = 4096;            // per operand tile in pair_dots
constexpr float G_EPS = 1.1920928955078125e-07f;  // torch.finfo(float32).eps (clamp_probs)

struct GraphTheta { float w[G_TMAXMAX + 1]; };  // theta(hop) for hop <= max_hop, computed on the host in double

// ------------------------------------------------------------------------------------------------------------
// pair_dots: A, Bm point at one video's [Cn][T][S] block.  Result in out_tt[T*T] (shared).  `tiles` holds
// 2 * G_CHUNK_FLOATS floats, `red` holds G_THREADS * 4 floats.
// ------------------------------------------------------------------------------------------------------------
__device__ __noinline__ static void pair_dots(const float* __restrict__ A, const float* __restrict__ Bm, int Cn, int T, int S,
                          float* tiles, float* red, float* out_tt)
{
    const int tid = threadIdx.x;
    const int npairs = T * T;
    const int PP = npairs < G_THREADS ? npairs : G_THREADS;     // pair lanes
    const int G = G_THREADS / PP;                               // element groups
    const int NP = (npairs + PP - 1) / PP;                      // pairs per lane (<= 4)
    const int plane = tid % PP, group = tid / PP;
    const bool active = group < G;
    int pi[4], pj[4];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int p = plane + k * PP;
        pi[k] = (p < npairs) ? p / T : 0;
        pj[k] = (p < npairs) ? p % T : 0;
    }
    float* tA = tiles;
    float* tB = tiles + G_CHUNK_FLOATS;
    int sc_ = S < 256 ? S : 256;                                // s-chunk; rows padded to an odd stride so the
    if (T * (sc_ | 1) > G_CHUNK_FLOATS) sc_ = G_CHUNK_FLOATS / T - 1;   // pair reads are bank-conflict-free
    const int use_sc = sc_, use_ld = sc_ | 1;
    int nC = G_CHUNK_FLOATS / (T * use_ld);                     // channels staged per tile
    if (nC < 1) nC = 1;

    for (int s0 = 0; s0 < S; s0 += use_sc) {
        const int sce = (S - s0 < use_sc) ? (S - s0) : use_sc;
        for (int c0 = 0; c0 < Cn; c0 += nC) {
            const int nce = (Cn - c0 < nC) ? (Cn - c0) : nC;
            __syncthreads();
            const int n_load = nce * T * sce;
            for (int idx = tid; idx < n_load; idx += G_THREADS) {
                const int row = idx / sce, s = idx - row * sce;       // row = c * T + i
                const size_t g = ((size_t)c0 * T + row) * S + s0 + s;
                tA[row * use_ld + s] = __ldg(A + g);
                tB[row * use_ld + s] = __ldg(Bm + g);
            }
            __syncthreads();
            if (active) {
                int c = 0, s = group;
                while (s >= sce) { s -= sce; ++c; }
                while (c < nce) {
                    const float* ra = tA + c * T * use_ld + s;
                    const float* rb = tB + c * T * use_ld + s;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k < NP) acc[k] = fmaf(ra[pi[k] * use_ld], rb[pj[k] * use_ld], acc[k]);
                    s += G;
                    while (s >= sce) { s -= sce; ++c; }
                }
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) red[k * G_THREADS + tid] = acc[k];
    __syncthreads();
    for (int p = tid; p < npairs; p += G_THREADS) {
        const int k = p / PP, pl = p - k * PP;
        float v = 0.f;
        for (int g = 0; g < G; ++g) v += red[k * G_THREADS + g * PP + pl];
        out_tt[p] = v;
    }
    __syncthreads();
}

// edge weight: theta(|i-j|) within max_hop, else 0  (temporal_graph.py:25-36, 204-210)
__device__ __forceinline__ float edge_w(const GraphTheta& th, int i, int j, int max_hop)
{
    const int hop = i > j ? i - j : j - i;
    return hop <= max_hop ? th.w[hop] : 0.f;
}

// forward T x T element work: logits (smem) -> sim, adj, s (smem + global)
__device__ static void adj_forward(float* lg /*in: logits, out: s*/, float* sim_s, float* adj_s, const float* u,
                            const GraphTheta& th, int T, int max_hop, float inv_temp,
                            float* sim_g, float* adj_g, float* s_g)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = warp; i < T; i += G_THREADS / 32) {
        const float x = lane < T ? lg[i * T + lane] : -INFINITY;
        const float m = warp_max(x);
        const float e = lane < T ? expf(x - m) : 0.f;
        const float sum = warp_sum(e);
        if (lane < T) {
            const int o = i * T + lane;
            const float sim = e / sum;                                  // F.softmax(dim=-1), :176
            const float adj = sim * edge_w(th, i, lane, max_hop);       // :204-210
            const float p  = fminf(fmaxf(adj, G_EPS), 1.f - G_EPS);     // clamp_probs
            const float uc = fminf(fmaxf(u[o], G_EPS), 1.f - G_EPS);          // (u may live in shared memory)
            const float z = (logf(uc) - log1pf(-uc) + logf(p) - log1pf(-p)) * inv_temp;   // LogitRelaxedBernoulli.rsample
            const float s = 1.f / (1.f + expf(-z));                     // SigmoidTransform
            sim_s[o] = sim; adj_s[o] = adj; lg[o] = s;
            sim_g[o] = sim; adj_g[o] = adj; s_g[o] = s;
        }
    }
    __syncthreads();
}

// backward T x T element work: ds (smem, in place -> d_logit)
__device__ static void adj_backward(float* ds, const float* __restrict__ sim_g, const float* __restrict__ adj_g,
                             const float* __restrict__ s_g, const GraphTheta& th, int T, int max_hop, float inv_temp)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = warp; i < T; i += G_THREADS / 32) {
        float sim = 0.f, dsim = 0.f;
        if (lane < T) {
            const int o = i * T + lane;
            sim = __ldg(sim_g + o);
            const float adj = __ldg(adj_g + o), s = __ldg(s_g + o);
            const float p = fminf(fmaxf(adj, G_EPS), 1.f - G_EPS);
            const bool inside = (adj >= G_EPS) && (adj <= 1.f - G_EPS);   // clamp passes gradient on [min, max]
            const float dadj = inside ? ds[o] * s * (1.f - s) * inv_temp / (p * (1.f - p)) : 0.f;
            dsim = dadj * edge_w(th, i, lane, max_hop);
        }
        const float dot = warp_sum(dsim * sim);
        if (lane < T) ds[i * T + lane] = sim * (dsim - dot);            // softmax backward
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------------------
// aggregate: items [item_begin, item_end) step item_stride of one video; item = (channel, VEC-wide column group)
// ------------------------------------------------------------------------------------------------------------
template <int TMAX, int VEC, bool kSmemIn = false>
__device__ __forceinline__ void aggregate_items(const float* __restrict__ in, float* __restrict__ out, const float* M,
                                                bool transpose, bool skip, int Cn, int T, int S,
                                                int item_begin, int item_stride)
{
    const int SV = S / VEC;
    const int n_items = Cn * SV;
    for (int item = item_begin; item < n_items; item += item_stride) {
        const int c = item / SV, sv = item - c * SV;
        const size_t base = (size_t)c * T * S + (size_t)sv * VEC;
        float x[TMAX][VEC];
#pragma unroll
        for (int j = 0; j < TMAX; ++j) {
            if (j < T) {
                if constexpr (VEC == 4) {
                    const float4* src = reinterpret_cast<const float4*>(in + base + (size_t)j * S);
                    const float4 v = kSmemIn ? *src : __ldg(src);          // (__ldg is a global-memory load)
                    x[j][0] = v.x; x[j][1] = v.y; x[j][2] = v.z; x[j][3] = v.w;
                } else {
                    x[j][0] = kSmemIn ? in[base + (size_t)j * S] : __ldg(in + base + (size_t)j * S);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < TMAX; ++i) {
            if (i < T) {
                float y[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) y[v] = skip ? x[i][v] : 0.f;
#pragma unroll
                for (int j = 0; j < TMAX; ++j) {
                    if (j < T) {
                        const float m = transpose ? M[j * T + i] : M[i * T + j];
#pragma unroll
                        for (int v = 0; v < VEC; ++v) y[v] = fmaf(m, x[j][v], y[v]);
                    }
                }
                if constexpr (VEC == 4)
                    *reinterpret_cast<float4*>(out + base + (size_t)i * S) = make_float4(y[0], y[1], y[2], y[3]);
                else
                    out[base + (size_t)i * S] = y[0];
            }
        }
    }
}

struct AggJob { const float* in; float* out; const float* M; int transpose; int skip; int Cn; int S; };
struct AggJobs { AggJob j[3]; int T; };

struct GraphArgs {
    const float* gq; const float* gk; int Cq; int S;
    const float* support; int C; int HW;
    int T; int B;
    const float* u;
    int max_hop; float inv_temp;
    float* sim; float* adj; float* s;          // [B, T, T]
    float* y;                                  // fwd out
    const float* dy;                           // bwd in
    float* d_gq; float* d_gk; float* d_support;
    float* dl;                                 // [B, T, T] scratch (bwd, split path)
    GraphTheta th;
};

constexpr int G_SMEM_FLOATS = 2 * G_CHUNK_FLOATS + 4 * G_THREADS + 3 * G_TMAXMAX * G_TMAXMAX;

static inline int pick_tmax(int T) { return T <= 4 ? 4 : T <= 8 ? 8 : T <= 16 ? 16 : 32; }

// host launchers
void fill_theta(GraphTheta& th, float alpha, int max_hop);
int graph_agg_launch(const AggJobs& jobs, int njobs, int B, cudaStream_t st);   // graph.cu
int graph_bwd_adj_launch(const GraphArgs& a, bool fused, cudaStream_t st);   // graph_bwd.cu
size_t graph_split_scratch_floats(int B, int T);                              // graph_split.cu
int graph_split_adj_launch(const GraphArgs& a, bool bwd, float* scratch, cudaStream_t st);
bool graph_smem_fits(const GraphArgs& a, bool bwd);                           // graph_smem.cu
int graph_smem_launch(const GraphArgs& a, bool bwd, cudaStream_t st);

}  // namespace gca
