// Device primitives and argument blocks of the temporal clip-graph head (shared by graph.cu / graph_bwd.cu).
#pragma once
#include "gca_common.cuh"

namespace gca {

constexpr int G_THREADS = 256;
constexpr int G_TMAXMAX = 32;
constexpr int G_CHUNK_FLOATS = 4096;            // per operand tile in pair_dots
constexpr float G_EPS = 1.1920928955078125e-07f;  // torch.finfo(float32).eps (clamp_probs)

struct GraphTheta { float w[G_TMAXMAX + 1]; };  // theta(hop) for hop <= max_hop, computed on the host in double
// default-OFF variants of the T x T chain (include/gca_b200.h, GCA_GRAPH_*): flags == 0 is the reference's arithmetic
struct GraphOpts { unsigned flags; float tau; int topk; float p_drop; };

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

// re-sampled edge weight from the (masked) adjacency entry: the reference's relaxed Bernoulli, or the hard seeded edge drop
__device__ __forceinline__ float resample_edge(float adj, float u, float inv_temp, const GraphOpts& opt)
{
    if (opt.flags & GCA_GRAPH_EDGE_DROP) return (u >= opt.p_drop) ? adj : 0.f;     // keep mask = [u >= p_drop]
    const float p  = fminf(fmaxf(adj, G_EPS), 1.f - G_EPS);     // clamp_probs
    const float uc = fminf(fmaxf(u, G_EPS), 1.f - G_EPS);
    const float z = (logf(uc) - log1pf(-uc) + logf(p) - log1pf(-p)) * inv_temp;   // LogitRelaxedBernoulli.rsample
    return 1.f / (1.f + expf(-z));                              // SigmoidTransform
}

// forward T x T element work: logits (smem) -> sim, adj, s (smem + global)
__device__ static void adj_forward(float* lg /*in: logits, out: s*/, float* sim_s, float* adj_s, const float* u,
                            const GraphTheta& th, int T, int max_hop, float inv_temp,
                            float* sim_g, float* adj_g, float* s_g, const GraphOpts opt = GraphOpts{0u, 0.f, 0, 0.f})
{
    __shared__ float deg_s[G_TMAXMAX];                                  // row sums of s (GCA_GRAPH_SYMNORM)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = warp; i < T; i += G_THREADS / 32) {
        const float x = lane < T ? lg[i * T + lane] : -INFINITY;
        const float m = warp_max(x);
        const float e = lane < T ? expf(x - m) : 0.f;
        const float sum = warp_sum(e);
        const float sim = e / sum;                                      // F.softmax(dim=-1), :176
        float adj = lane < T ? sim * edge_w(th, i, lane, max_hop) : -INFINITY;      // :204-210
        if (opt.flags & GCA_GRAPH_THRESHOLD) { if (lane < T && adj < opt.tau) adj = 0.f; }
        if (opt.flags & GCA_GRAPH_TOPK) {
            // rank of this entry in its row: larger value first, ties towards the lower column (deterministic, integer work)
            int rank = 0;
            for (int mcol = 0; mcol < T; ++mcol) {
                const float o = __shfl_sync(0xffffffffu, adj, mcol);
                rank += (o > adj || (o == adj && mcol < lane)) ? 1 : 0;
            }
            if (lane < T && rank >= opt.topk) adj = 0.f;
        }
        float sv = 0.f;
        if (lane < T) sv = resample_edge(adj, u[i * T + lane], inv_temp, opt);
        if (opt.flags & GCA_GRAPH_SYMNORM) {
            const float d = warp_sum(sv);
            if (lane == 0) deg_s[i] = d;
        }
        if (lane < T) {
            const int o = i * T + lane;
            sim_s[o] = sim; adj_s[o] = adj; lg[o] = sv;
            sim_g[o] = sim; adj_g[o] = adj; s_g[o] = sv;
        }
    }
    __syncthreads();
    if (opt.flags & GCA_GRAPH_SYMNORM) {
        // s <- D^-1/2 s D^-1/2,  D = diag(row sums of s)  (degrees below G_EPS count as G_EPS)
        for (int p = threadIdx.x; p < T * T; p += G_THREADS) {
            const int i = p / T, j = p - i * T;
            const float v = lg[p] * rsqrtf(fmaxf(deg_s[i], G_EPS)) * rsqrtf(fmaxf(deg_s[j], G_EPS));
            lg[p] = v; s_g[p] = v;
        }
        __syncthreads();
    }
}

// backward T x T element work: ds (smem, in place -> d_logit).  `u` is only read by the variants that need the pre-
// normalisation s again (GCA_GRAPH_SYMNORM) or the keep mask (GCA_GRAPH_EDGE_DROP).
__device__ static void adj_backward(float* ds, const float* __restrict__ sim_g, const float* __restrict__ adj_g,
                             const float* __restrict__ s_g, const GraphTheta& th, int T, int max_hop, float inv_temp,
                             const GraphOpts opt = GraphOpts{0u, 0.f, 0, 0.f}, const float* __restrict__ u = nullptr)
{
    __shared__ float sraw_s[G_TMAXMAX * G_TMAXMAX];                     // GCA_GRAPH_SYMNORM: s before the normalisation
    __shared__ float deg_s[G_TMAXMAX], cvec_s[G_TMAXMAX];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool symnorm = (opt.flags & GCA_GRAPH_SYMNORM) != 0;
    if (symnorm) {
        // s' = s_ij r_i r_j, r = d^-1/2, d_i = sum_j s_ij:  dL/ds_ij = g_ij r_i r_j - 1/2 d_i^-3/2 c_i,
        // c_i = dL/dr_i = sum_j g_ij s_ij r_j + sum_k g_ki s_ki r_k      (g = dL/ds')
        for (int i = warp; i < T; i += G_THREADS / 32) {
            float sv = 0.f;
            if (lane < T) sv = resample_edge(__ldg(adj_g + i * T + lane), __ldg(u + i * T + lane), inv_temp, opt);
            const float d = warp_sum(sv);
            if (lane < T) sraw_s[i * T + lane] = sv;
            if (lane == 0) deg_s[i] = fmaxf(d, G_EPS);
        }
        __syncthreads();
        for (int i = warp; i < T; i += G_THREADS / 32) {
            float c = 0.f;
            if (lane < T) {
                const float rj = rsqrtf(deg_s[lane]);
                c = ds[i * T + lane] * sraw_s[i * T + lane] * rj + ds[lane * T + i] * sraw_s[lane * T + i] * rj;
            }
            c = warp_sum(c);
            if (lane == 0) cvec_s[i] = c;
        }
        __syncthreads();
        for (int p = threadIdx.x; p < T * T; p += G_THREADS) {
            const int i = p / T, j = p - i * T;
            const float di = deg_s[i];
            const float below = (di <= G_EPS) ? 0.f : 0.5f * cvec_s[i] / (di * sqrtf(di));      // clamped degree: no gradient
            ds[p] = ds[p] * rsqrtf(di) * rsqrtf(deg_s[j]) - below;
        }
        __syncthreads();
    }
    for (int i = warp; i < T; i += G_THREADS / 32) {
        float sim = 0.f, dsim = 0.f;
        if (lane < T) {
            const int o = i * T + lane;
            sim = __ldg(sim_g + o);
            const float adj = __ldg(adj_g + o);
            float dadj;
            if (opt.flags & GCA_GRAPH_EDGE_DROP) {
                dadj = (__ldg(u + o) >= opt.p_drop) ? ds[o] : 0.f;
                if (adj == 0.f) dadj = 0.f;                             // removed by the hop mask / threshold / top-k
            } else {
                const float s = symnorm ? sraw_s[o] : __ldg(s_g + o);
                const float p = fminf(fmaxf(adj, G_EPS), 1.f - G_EPS);
                const bool inside = (adj >= G_EPS) && (adj <= 1.f - G_EPS);   // clamp passes gradient on [min, max]
                dadj = inside ? ds[o] * s * (1.f - s) * inv_temp / (p * (1.f - p)) : 0.f;
            }
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

struct AggJob { const float* in; float* out; const float* M; int transpose; int skip; int Cn; int S;
                int reverse; };   // reverse: videos in descending order (re-read what the previous kernel touched last while it is in L2)
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
    GraphOpts opt;                             // default-OFF variants (flags == 0: the reference's arithmetic)
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
