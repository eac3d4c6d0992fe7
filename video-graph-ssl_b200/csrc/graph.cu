// K4: temporal clip-graph head, forward (lib/ops/module_wrappers/temporal_graph.py:150-239 between the 1x1x1
// convolutions), the grid-parallel aggregation kernel and the C-ABI entry points of both directions.
// Two device primitives (graph_dev.cuh) cover everything:
//   pair_dots : M[i][j] = sum_{c,s} A[c][i][s] * B[c][j][s]        (similarity logits; ds in backward)
//   aggregate : out[c][i][:] = sum_j M(i,j) in[c][j][:] (+ in[c][i][:])   (GCN aggregation; all three input grads)
// Small videos (embedding-level heads) run ONE CTA per video with everything in shared memory; large feature maps
// split into an adjacency kernel (one CTA per video) and a grid-parallel aggregation kernel with 128-bit loads.
// All sums have a fixed order (deterministic).
#include "graph_dev.cuh"
#include <stdlib.h>

namespace gca {

// forward, one CTA per video.  kAgg: also aggregate (fused path); otherwise only sim/adj/s are produced.
template <int TMAX, int VH, bool kAgg>
__global__ void __launch_bounds__(G_THREADS)
graph_fwd_kernel(const GraphArgs a)
{
    extern __shared__ __align__(16) float gsm[];
    float* tiles = gsm;
    float* red = tiles + 2 * G_CHUNK_FLOATS;
    float* m0 = red + 4 * G_THREADS;                  // logits -> s
    float* m1 = m0 + G_TMAXMAX * G_TMAXMAX;           // sim
    float* m2 = m1 + G_TMAXMAX * G_TMAXMAX;           // adj
    const int b = blockIdx.x, T = a.T;
    const size_t tt = (size_t)b * T * T;
    pair_dots(a.gq + (size_t)b * a.Cq * T * a.S, a.gk + (size_t)b * a.Cq * T * a.S, a.Cq, T, a.S, tiles, red, m0);
    adj_forward(m0, m1, m2, a.u + tt, a.th, T, a.max_hop, a.inv_temp, a.sim + tt, a.adj + tt, a.s + tt, a.opt);
    if constexpr (kAgg) {
        const size_t off = (size_t)b * a.C * T * a.HW;
        aggregate_items<TMAX, VH>(a.support + off, a.y + off, m0, false, true, a.C, T, a.HW, threadIdx.x, G_THREADS);
    }
}

// grid-parallel aggregation for large feature maps: grid = (chunks, B, jobs)
// T <= 8 with 128-bit columns (the shipped shapes): a thread's T row loads do not depend on the mixing matrix, so they are
// issued BEFORE the matrix is fetched and the block synchronises -- one memory round trip per CTA instead of two -- and the
// register budget (64) lets 4 CTAs per SM keep ~128 KB of loads in flight.
template <int TMAX>
__global__ void __launch_bounds__(G_THREADS, 4)
graph_agg_vec4_kernel(const AggJobs jobs)
{
    __shared__ float M[TMAX * TMAX];
    const AggJob jb = jobs.j[blockIdx.z];
    const int T = jobs.T, b = jb.reverse ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y, S = jb.S;
    const size_t off = (size_t)b * jb.Cn * T * S;
    const int stride = gridDim.x * G_THREADS;
    const int SV = S / 4, n_items = jb.Cn * SV;
    const float* in = jb.in + off;
    float* out = jb.out + off;
    float4 x[TMAX];
    int item = blockIdx.x * G_THREADS + threadIdx.x;
    size_t base = 0;
    if (item < n_items) {
        const int c = item / SV, sv = item - c * SV;
        base = (size_t)c * T * S + (size_t)sv * 4;
#pragma unroll
        for (int j = 0; j < TMAX; ++j) if (j < T) x[j] = __ldg(reinterpret_cast<const float4*>(in + base + (size_t)j * S));
    }
    // M is stored [i][j] in shared memory already transposed if the job asks for it, padded to TMAX columns
    for (int p = threadIdx.x; p < TMAX * TMAX; p += G_THREADS) {
        const int i = p / TMAX, j = p - i * TMAX;
        float m = 0.f;
        if (i < T && j < T) m = __ldg(jb.M + (size_t)b * T * T + (jb.transpose ? j * T + i : i * T + j));
        M[p] = m;
    }
    __syncthreads();
    while (item < n_items) {
#pragma unroll
        for (int i = 0; i < TMAX; ++i) {
            if (i < T) {
                float4 y = jb.skip ? x[i] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int j = 0; j < TMAX; ++j) {
                    if (j < T) {
                        const float m = M[i * TMAX + j];
                        y.x = fmaf(m, x[j].x, y.x); y.y = fmaf(m, x[j].y, y.y);
                        y.z = fmaf(m, x[j].z, y.z); y.w = fmaf(m, x[j].w, y.w);
                    }
                }
                *reinterpret_cast<float4*>(out + base + (size_t)i * S) = y;
            }
        }
        item += stride;
        if (item < n_items) {
            const int c = item / SV, sv = item - c * SV;
            base = (size_t)c * T * S + (size_t)sv * 4;
#pragma unroll
            for (int j = 0; j < TMAX; ++j) if (j < T) x[j] = __ldg(reinterpret_cast<const float4*>(in + base + (size_t)j * S));
        }
    }
}

// T <= 8, rows that are not 128-bit addressable (e.g. the 7x7 projections): scalar columns, four independent items per
// thread so that 32 loads are in flight before the mixing matrix is needed
constexpr int AGG_ILP = 4;
template <int TMAX>
__global__ void __launch_bounds__(G_THREADS, 2)
graph_agg_scalar_kernel(const AggJobs jobs)
{
    __shared__ float M[TMAX * TMAX];
    const AggJob jb = jobs.j[blockIdx.z];
    const int T = jobs.T, b = jb.reverse ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y, S = jb.S;
    const size_t off = (size_t)b * jb.Cn * T * S;
    const int n_items = jb.Cn * S;
    const float* in = jb.in + off;
    float* out = jb.out + off;
    const int stride = gridDim.x * G_THREADS;
    float x[AGG_ILP][TMAX];
    size_t base[AGG_ILP];
    bool ok[AGG_ILP];
    int item0 = blockIdx.x * G_THREADS + threadIdx.x;
    auto load = [&]() {
#pragma unroll
        for (int u = 0; u < AGG_ILP; ++u) {
            const int item = item0 + u * stride;
            ok[u] = item < n_items;
            if (ok[u]) {
                const int c = item / S, sv = item - c * S;
                base[u] = (size_t)c * T * S + sv;
#pragma unroll
                for (int j = 0; j < TMAX; ++j) if (j < T) x[u][j] = __ldg(in + base[u] + (size_t)j * S);
            }
        }
    };
    load();
    for (int p = threadIdx.x; p < TMAX * TMAX; p += G_THREADS) {
        const int i = p / TMAX, j = p - i * TMAX;
        float m = 0.f;
        if (i < T && j < T) m = __ldg(jb.M + (size_t)b * T * T + (jb.transpose ? j * T + i : i * T + j));
        M[p] = m;
    }
    __syncthreads();
    while (item0 < n_items) {
#pragma unroll
        for (int u = 0; u < AGG_ILP; ++u) {
            if (ok[u]) {
#pragma unroll
                for (int i = 0; i < TMAX; ++i) {
                    if (i < T) {
                        float y = jb.skip ? x[u][i] : 0.f;
#pragma unroll
                        for (int j = 0; j < TMAX; ++j) if (j < T) y = fmaf(M[i * TMAX + j], x[u][j], y);
                        out[base[u] + (size_t)i * S] = y;
                    }
                }
            }
        }
        item0 += AGG_ILP * stride;
        if (item0 < n_items) load();
    }
}

template <int TMAX>
__global__ void __launch_bounds__(G_THREADS, (TMAX <= 8) ? 2 : 1)
graph_agg_kernel(const AggJobs jobs)
{
    __shared__ float M[G_TMAXMAX * G_TMAXMAX];
    const AggJob jb = jobs.j[blockIdx.z];
    const int T = jobs.T, b = jb.reverse ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
    for (int p = threadIdx.x; p < T * T; p += G_THREADS) M[p] = __ldg(jb.M + (size_t)b * T * T + p);
    __syncthreads();
    const size_t off = (size_t)b * jb.Cn * T * jb.S;
    const int begin = blockIdx.x * G_THREADS + threadIdx.x, stride = gridDim.x * G_THREADS;
    if (TMAX <= 16 && jb.S % 4 == 0)
        aggregate_items<TMAX, (TMAX <= 16 ? 4 : 1)>(jb.in + off, jb.out + off, M, jb.transpose != 0, jb.skip != 0, jb.Cn, T,
                                                    jb.S, begin, stride);
    else
        aggregate_items<TMAX, 1>(jb.in + off, jb.out + off, M, jb.transpose != 0, jb.skip != 0, jb.Cn, T, jb.S, begin, stride);
}

void fill_theta(GraphTheta& th, float alpha, int max_hop)
{
    for (int h = 0; h <= G_TMAXMAX; ++h) {
        const double e = exp(-(double)h);
        th.w[h] = (h <= max_hop) ? (float)(e / (1.0 + e * e) + (double)alpha) : 0.f;   // temporal_graph.py:206
    }
}

template <int TM, int VH, bool kAgg>
static int launch_fwd_one(const GraphArgs& a, cudaStream_t st)
{
    const size_t smem = G_SMEM_FLOATS * sizeof(float);
    auto kern = graph_fwd_kernel<TM, VH, kAgg>;
    GCA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<a.B, G_THREADS, smem, st>>>(a);
    GCA_LAUNCH_CHECK("graph_fwd_kernel");
    count_launch(1);
    return GCA_OK;
}

static int graph_fwd_adj_launch(const GraphArgs& a, bool fused, cudaStream_t st)
{
    if (!fused) return launch_fwd_one<4, 1, false>(a, st);
    const bool v4 = (a.HW % 4 == 0);
    switch (pick_tmax(a.T)) {
        case 4:  return v4 ? launch_fwd_one<4, 4, true>(a, st) : launch_fwd_one<4, 1, true>(a, st);
        case 8:  return v4 ? launch_fwd_one<8, 4, true>(a, st) : launch_fwd_one<8, 1, true>(a, st);
        case 16: return v4 ? launch_fwd_one<16, 4, true>(a, st) : launch_fwd_one<16, 1, true>(a, st);
        default: return launch_fwd_one<32, 1, true>(a, st);
    }
}

static int agg_launch_group(const AggJobs& jobs, int njobs, int B, bool vec4, cudaStream_t st)
{
    long long max_items = 0;
    for (int i = 0; i < njobs; ++i) {
        const long long it = (long long)jobs.j[i].Cn * (jobs.j[i].S % 4 == 0 ? jobs.j[i].S / 4 : jobs.j[i].S);
        if (it > max_items) max_items = it;
    }
    int chunks = (int)((max_items + G_THREADS - 1) / G_THREADS);
    if (chunks < 1) chunks = 1;
    if (chunks > 256) chunks = 256;
    dim3 grid(chunks, B, njobs);
    if (vec4) {
        if (jobs.T <= 4) graph_agg_vec4_kernel<4><<<grid, G_THREADS, 0, st>>>(jobs);
        else             graph_agg_vec4_kernel<8><<<grid, G_THREADS, 0, st>>>(jobs);
        GCA_LAUNCH_CHECK("graph_agg_vec4_kernel");
        count_launch(1);
        return GCA_OK;
    }
    if (jobs.T <= 8) {
        long long mi = 0;
        for (int i = 0; i < njobs; ++i) { const long long it = (long long)jobs.j[i].Cn * jobs.j[i].S; if (it > mi) mi = it; }
        int ch = (int)((mi + (long long)G_THREADS * AGG_ILP - 1) / ((long long)G_THREADS * AGG_ILP));
        // about two resident CTAs per SM over the whole grid; each thread then loops with its next 32 loads issued right
        // behind the current stores (one wave, no per-CTA ramp per item)
        const int want = (2 * sm_count_cached() + B * njobs - 1) / (B * njobs);
        if (ch > want) ch = want;
        if (ch < 1) ch = 1;
        if (ch > 256) ch = 256;
        dim3 g2(ch, B, njobs);
        if (jobs.T <= 4) graph_agg_scalar_kernel<4><<<g2, G_THREADS, 0, st>>>(jobs);
        else             graph_agg_scalar_kernel<8><<<g2, G_THREADS, 0, st>>>(jobs);
        GCA_LAUNCH_CHECK("graph_agg_scalar_kernel");
        count_launch(1);
        return GCA_OK;
    }
    switch (pick_tmax(jobs.T)) {
        case 4:  graph_agg_kernel<4><<<grid, G_THREADS, 0, st>>>(jobs); break;
        case 8:  graph_agg_kernel<8><<<grid, G_THREADS, 0, st>>>(jobs); break;
        case 16: graph_agg_kernel<16><<<grid, G_THREADS, 0, st>>>(jobs); break;
        default: graph_agg_kernel<32><<<grid, G_THREADS, 0, st>>>(jobs); break;
    }
    GCA_LAUNCH_CHECK("graph_agg_kernel");
    count_launch(1);
    return GCA_OK;
}

// jobs whose rows are 128-bit addressable (T <= 8) go to the vec4 kernel, the rest to the generic one: two launches at most
int graph_agg_launch(const AggJobs& jobs, int njobs, int B, cudaStream_t st)
{
    AggJobs fast{}, slow{};
    fast.T = slow.T = jobs.T;
    int nf = 0, ns = 0;
    for (int i = 0; i < njobs; ++i) {
        const AggJob& jb = jobs.j[i];
        const bool v = jobs.T <= 8 && jb.S % 4 == 0 &&
                       ((reinterpret_cast<uintptr_t>(jb.in) | reinterpret_cast<uintptr_t>(jb.out)) & 15) == 0;
        if (v) fast.j[nf++] = jb; else slow.j[ns++] = jb;
    }
    if (nf && ns) {
        // the two groups are independent (different outputs): the scalar group runs on a forked side stream (a parallel
        // branch under graph capture) so that its lower-bandwidth loads hide under the 128-bit group
        cudaStream_t side;
        int rc = side_stream_fork(st, &side);
        if (rc != GCA_OK) return rc;
        rc = agg_launch_group(slow, ns, B, false, side);
        if (rc != GCA_OK) return rc;
        rc = agg_launch_group(fast, nf, B, true, st);
        if (rc != GCA_OK) return rc;
        return side_stream_join(st);
    }
    if (nf) { const int rc = agg_launch_group(fast, nf, B, true, st); if (rc != GCA_OK) return rc; }
    if (ns) { const int rc = agg_launch_group(slow, ns, B, false, st); if (rc != GCA_OK) return rc; }
    return GCA_OK;
}

// one CTA per video is the right shape while a video's tensors are small; above this the aggregation is spread
// over the whole grid
static bool use_fused(const GraphArgs& a)
{
    const size_t per_video = ((size_t)a.C * a.HW + 2ull * a.Cq * a.S) * a.T * sizeof(float);
    return per_video <= 256u * 1024u;
}

static int check_graph_opts(const char* fn, const GcaGraphOpts* o, int T, const float* u, GraphOpts* out)
{
    *out = GraphOpts{0u, 0.f, 0, 0.f};
    if (!o || o->flags == 0u) return GCA_OK;
    const unsigned all = GCA_GRAPH_THRESHOLD | GCA_GRAPH_TOPK | GCA_GRAPH_EDGE_DROP | GCA_GRAPH_SYMNORM;
    if (o->flags & ~all) return set_err(GCA_ERR_UNSUPPORTED, "%s: flags 0x%x not supported", fn, o->flags);
    GCA_CHECK_ARG(!(o->flags & GCA_GRAPH_TOPK) || (o->topk >= 1 && o->topk <= T), "%s: topk=%d outside [1, T=%d]", fn, o->topk, T);
    GCA_CHECK_ARG(!(o->flags & GCA_GRAPH_EDGE_DROP) || (o->p_drop >= 0.f && o->p_drop <= 1.f), "%s: p_drop outside [0, 1]", fn);
    GCA_CHECK_ARG(!(o->flags & GCA_GRAPH_THRESHOLD) || o->tau >= 0.f, "%s: tau < 0", fn);
    GCA_CHECK_ARG(!(o->flags & (GCA_GRAPH_EDGE_DROP | GCA_GRAPH_SYMNORM)) || u, "%s: these variants need the uniforms u", fn);
    *out = GraphOpts{o->flags, o->tau, o->topk, o->p_drop};
    return GCA_OK;
}

static int check_graph_args(const char* fn, int Cq, int S, int C, int HW, int T, int B, int max_hop, float temperature,
                            unsigned flags)
{
    GCA_CHECK_ARG(Cq >= 1 && S >= 1 && C >= 1 && HW >= 1 && B >= 1, "%s: bad sizes", fn);
    GCA_CHECK_ARG(T >= 1 && T <= G_TMAXMAX, "%s: T=%d outside [1, %d]", fn, T, G_TMAXMAX);
    GCA_CHECK_ARG(max_hop >= 0, "%s: max_hop < 0", fn);
    GCA_CHECK_ARG(temperature > 0.f, "%s: temperature must be > 0", fn);
    if (flags != GCA_GRAPH_REFERENCE) return set_err(GCA_ERR_UNSUPPORTED, "%s: flags 0x%x not supported", fn, flags);
    return GCA_OK;
}

}  // namespace gca

extern "C" size_t gca_graph_workspace_bytes(int B, int T)
{
    // [B, T, T] d_logit + the per-chunk pair-dot partials of the split path
    return (B > 0 && T > 0) ? ((size_t)B * T * T + gca::graph_split_scratch_floats(B, T)) * sizeof(float) : 0;
}

extern "C" int gca_graph_fwd(const float* gq, const float* gk, int Cq, int S, const float* support, int C, int HW,
                             int T, int B, const float* u, float alpha, int max_hop, float temperature, unsigned flags,
                             float* sim, float* adj, float* s, float* y, void* workspace, size_t workspace_bytes,
                             void* stream)
{
    if (flags != GCA_GRAPH_REFERENCE)
        return gca::set_err(GCA_ERR_UNSUPPORTED, "gca_graph_fwd: flags 0x%x need gca_graph_fwd_ex (parameters in GcaGraphOpts)", flags);
    return gca_graph_fwd_ex(gq, gk, Cq, S, support, C, HW, T, B, u, alpha, max_hop, temperature, nullptr, sim, adj, s, y,
                            workspace, workspace_bytes, stream);
}

extern "C" int gca_graph_fwd_ex(const float* gq, const float* gk, int Cq, int S, const float* support, int C, int HW,
                                int T, int B, const float* u, float alpha, int max_hop, float temperature,
                                const GcaGraphOpts* opts, float* sim, float* adj, float* s, float* y, void* workspace,
                                size_t workspace_bytes, void* stream)
{
    using namespace gca;
    GCA_CHECK_ARG(gq && gk && support && u && sim && adj && s && y, "gca_graph_fwd: null pointer");
    int rc = check_graph_args("gca_graph_fwd", Cq, S, C, HW, T, B, max_hop, temperature, 0u);
    if (rc != GCA_OK) return rc;
    GraphOpts gopt;
    rc = check_graph_opts("gca_graph_fwd_ex", opts, T, u, &gopt);
    if (rc != GCA_OK) return rc;
    GraphArgs a{};
    a.opt = gopt;
    a.gq = gq; a.gk = gk; a.Cq = Cq; a.S = S; a.support = support; a.C = C; a.HW = HW; a.T = T; a.B = B; a.u = u;
    a.max_hop = max_hop; a.inv_temp = 1.f / temperature; a.sim = sim; a.adj = adj; a.s = s; a.y = y;
    fill_theta(a.th, alpha, max_hop);
    cudaStream_t st = (cudaStream_t)stream;
    if (gopt.flags == 0u && graph_smem_fits(a, false)) return graph_smem_launch(a, false, st);
    if (use_fused(a)) return graph_fwd_adj_launch(a, true, st);
    if (!workspace || workspace_bytes < gca_graph_workspace_bytes(B, T))
        return set_err(GCA_ERR_WORKSPACE, "gca_graph_fwd: workspace of %zu bytes needed", gca_graph_workspace_bytes(B, T));
    rc = graph_split_adj_launch(a, false, (float*)workspace + (size_t)B * T * T, st);
    if (rc != GCA_OK) return rc;
    AggJobs jobs{};
    jobs.T = T;
    jobs.j[0] = AggJob{support, y, s, 0, 1, C, HW, 0};
    return graph_agg_launch(jobs, 1, B, st);
}

extern "C" int gca_graph_bwd(const float* gq, const float* gk, int Cq, int S, const float* support, int C, int HW,
                             int T, int B, const float* sim, const float* adj, const float* s, const float* dy,
                             float alpha, int max_hop, float temperature, unsigned flags,
                             float* d_gq, float* d_gk, float* d_support, void* workspace, size_t workspace_bytes,
                             void* stream)
{
    if (flags != GCA_GRAPH_REFERENCE)
        return gca::set_err(GCA_ERR_UNSUPPORTED, "gca_graph_bwd: flags 0x%x need gca_graph_bwd_ex (parameters in GcaGraphOpts)", flags);
    return gca_graph_bwd_ex(gq, gk, Cq, S, support, C, HW, T, B, sim, adj, s, dy, nullptr, alpha, max_hop, temperature, nullptr,
                            d_gq, d_gk, d_support, workspace, workspace_bytes, stream);
}

extern "C" int gca_graph_bwd_ex(const float* gq, const float* gk, int Cq, int S, const float* support, int C, int HW,
                                int T, int B, const float* sim, const float* adj, const float* s, const float* dy,
                                const float* u, float alpha, int max_hop, float temperature, const GcaGraphOpts* opts,
                                float* d_gq, float* d_gk, float* d_support, void* workspace, size_t workspace_bytes,
                                void* stream)
{
    using namespace gca;
    GCA_CHECK_ARG(gq && gk && support && sim && adj && s && dy && d_gq && d_gk && d_support, "gca_graph_bwd: null pointer");
    int rc = check_graph_args("gca_graph_bwd", Cq, S, C, HW, T, B, max_hop, temperature, 0u);
    if (rc != GCA_OK) return rc;
    GraphOpts gopt;
    rc = check_graph_opts("gca_graph_bwd_ex", opts, T, u, &gopt);
    if (rc != GCA_OK) return rc;
    GraphArgs a{};
    a.opt = gopt; a.u = gopt.flags ? u : nullptr;
    a.gq = gq; a.gk = gk; a.Cq = Cq; a.S = S; a.support = support; a.C = C; a.HW = HW; a.T = T; a.B = B;
    a.max_hop = max_hop; a.inv_temp = 1.f / temperature;
    a.sim = const_cast<float*>(sim); a.adj = const_cast<float*>(adj); a.s = const_cast<float*>(s);
    a.dy = dy; a.d_gq = d_gq; a.d_gk = d_gk; a.d_support = d_support;
    fill_theta(a.th, alpha, max_hop);
    cudaStream_t st = (cudaStream_t)stream;
    if (gopt.flags == 0u && graph_smem_fits(a, true)) return graph_smem_launch(a, true, st);
    if (use_fused(a)) return graph_bwd_adj_launch(a, true, st);
    if (!workspace || workspace_bytes < gca_graph_workspace_bytes(B, T))
        return set_err(GCA_ERR_WORKSPACE, "gca_graph_bwd: workspace of %zu bytes needed", gca_graph_workspace_bytes(B, T));
    a.dl = (float*)workspace;
    // d_support = s^T dy + dy uses the FORWARD's s: it does not depend on the backward chain, so it runs on a forked side
    // stream (a parallel branch under graph capture) BESIDE the pair-dot kernel.  Both stream dy video by video in ascending
    // order; whichever is behind finds the lines in L2, so dy comes from HBM once per backward instead of twice (an in-kernel
    // fusion of the two was measured slower: the extra stores and FMAs cost the pair-dot kernel its loads in flight).
    static int early_on = -1;                         // GCA_GRAPH_NOEARLY=1: the aggregation waits for the chain (A/B timing)
    if (early_on < 0) { const char* e = getenv("GCA_GRAPH_NOEARLY"); early_on = (e && e[0] == '1') ? 0 : 1; }
    const bool early = early_on && T <= 8 && HW % 4 == 0 &&
                       ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(d_support)) & 15) == 0;
    AggJobs first{};
    first.T = T;
    first.j[0] = AggJob{dy, d_support, s, 1, 1, C, HW, 0};
    if (early) {
        cudaStream_t side;
        rc = side_stream_fork(st, &side);
        if (rc != GCA_OK) return rc;
        rc = graph_agg_launch(first, 1, B, side);
        if (rc != GCA_OK) return rc;
    }
    rc = graph_split_adj_launch(a, true, (float*)workspace + (size_t)B * T * T, st);
    if (rc != GCA_OK) return rc;
    AggJobs jobs{};
    jobs.T = T;
    int nj = 0;
    if (!early) jobs.j[nj++] = AggJob{dy, d_support, s, 1, 1, C, HW, 1};      // (descending: the tail of dy is still in L2)
    jobs.j[nj++] = AggJob{gk, d_gq, a.dl, 0, 0, Cq, S, 0};
    jobs.j[nj++] = AggJob{gq, d_gk, a.dl, 1, 0, Cq, S, 0};
    rc = graph_agg_launch(jobs, nj, B, st);
    if (rc != GCA_OK) return rc;
    return early ? side_stream_join(st) : GCA_OK;
}
