// Parameter block shared by the two InfoNCE stream kernel families (ffma, tcgen05).
#pragma once
#include "gca_common.cuh"

namespace gca {

struct InfoNceStreamParams {
    const float* q;          // [B, d] fp32
    const float* k;          // [B, d] fp32 (positive keys)
    const void*  queue;      // [K, d] fp32 or bf16
    int          B;
    long long    K;
    int          d;
    float        inv_T;
    const float* lse_fixed;  // [B] row log-sum-exp (fixed-max / backward-recompute mode) or NULL
    // split partials (workspace)
    unsigned int* counter;
    float* part_max;         // [nsplit, Bpad]
    float* part_sum;         // [nsplit, Bpad]
    int*   part_cnt;         // [nsplit, Bpad]
    float* part_acc;         // [nsplit, Bpad, d] or NULL (no gradient wanted)
    int    nsplit;
    int    Bpad;
    float* pos_out;          // [B] or NULL
    float* logits_out;       // [B, ld_logits] or NULL (ffma family only)
    long long ld_logits;
    // tcgen05 family only: per-step scratch filled by infonce_prep_kernel
    void*  q_bf16_ws;        // [Bpad, 128] bf16
    float* pos_ws;           // [Bpad] positive logits (natural-log units)
    float  T_;               // 1 / inv_T
    int    skip_prep;        // reuse q_bf16_ws / pos_ws from the previous launch on this workspace (profiling)
    PeerXchg xchg;           // tcgen05 family: extra CTAs of the prep kernel push k[B, d] to every peer's mailbox (off if null)
    // projection-tail fusion (tcgen05 family): q and k are UN-normalised; the prep kernel L2-normalises both rows
    float* k_hat;            // [B.., d] out: the positive keys staged for the finalize kernel (normalised if `normalize`)
    float* inv_nq;           // [Bpad] out: 1 / max(||q||, 1e-12) (normalize only)
    int    normalize;
    int    rank_cap;         // tcgen05 single-pass kernel: > 0 = the caller only needs ranks below this value (top-k hits)
    int    gather_Bl;        // > 0 (K-sharded step over peer memory): q = this rank's [q_loc; k_loc] block of 2 * gather_Bl rows;
                             // the prep launch pushes it to every peer and reads all B = W * gather_Bl rows from the mailbox
    float  q_scale;          // factor folded into the bf16 queries by the prep kernel (1, or log2(e)/T for infonce_tcx)
    int    no_prep_wait;     // infonce_tcx: the prep launch triggers this kernel only after its rows are written (late trigger):
                             // do not wait for its completion (push CTAs of a peer exchange ride in it)
};

// ffma family (infonce_ffma.cu)
int infonce_ffma_nsplit(int B, long long K, int d);
int infonce_ffma_launch(const InfoNceStreamParams& P, int dtype_queue, bool fixed_max, cudaStream_t st);
// tcgen05 family (infonce_tc.cu)
int infonce_tc_nsplit(int B, long long K);
int infonce_tc_launch(const InfoNceStreamParams& P, bool fixed_max, cudaStream_t st);

// tcgen05 fp32-grade family (infonce_tc32.cu): `extra` = infonce_tc32_extra_ws() bytes behind the common workspace
int infonce_tc32_nsplit(int B, long long K);
size_t infonce_tc32_extra_ws(int B, long long K);
int infonce_tc32_launch(const InfoNceStreamParams& P, void* extra, cudaStream_t st);

// finalize.cu
enum FinalizeMode { FIN_FULL = 0, FIN_SHARD = 1, FIN_BWD = 2 };
struct FinalizeParams {
    // split partials
    unsigned int* counter;
    const float* part_max; const float* part_sum; const int* part_cnt; const float* part_acc;
    int nsplit, Bpad, B, d;
    float inv_T;
    const float* k;          // [B, d]
    const float* pos;        // [B]
    const float* lse_in;     // FIN_BWD: given lse
    float grad_scale;        // FIN_BWD
    // outputs
    float* lse; float* loss_rows; int* rank_gt; float* dq; float* loss_mean;     // FIN_FULL / FIN_BWD (dq)
    int* top_hits;                                                               // FIN_FULL: [2] rows with rank < 1, < 5
    float* out_max; float* out_sum; int* out_cnt; float* out_acc;                // FIN_SHARD
    // optional fused enqueue (FIN_FULL): keys [enq_N, d] into the full ring queue [enq_K, d]
    void* enq_queue; int enq_dtype; long long enq_K; const float* enq_keys; int enq_N;
    long long enq_index; long long* enq_state;
    long long enq_kbegin, enq_kend;   // enq_kend > 0: enq_queue is the shard holding global slots [enq_kbegin, enq_kend)
    PeerXchg xchg;                 // enq_keys == NULL and xchg on: the rows come from this rank's mailbox (all W*B of them)
    unsigned long long* timebuf;   // bring-up only (tools/tc_timeline.py): entry / exit time stamps
    const float* zq; const float* inv_nq;   // projection-tail fusion: dq is pushed back through q = zq / ||zq|| (NULL = off)
    // K-sharded step over peer memory (gca_shard_step_peer; protocol at PeerMerge in gca_common.cuh).  FIN_SHARD launch: row
    // b's merged partial goes straight into its OWNER's mailbox (remote stores + one released row count per CTA).  FIN_FULL
    // launch with merge.wait: the "splits" are the W ranks' partials of this rank's rows, read from this rank's mailbox.
    PeerMerge merge;               // off if mailboxes == nullptr
    int range_checked;             // the stream kernel reports out-of-range logits in control word 6 (tcgen05 family)
    int pk_nb, pk_frac;            // set by infonce_finalize_launch: packed loss/hits/ticket word (0 = unpacked path)
};
int infonce_finalize_launch(const FinalizeParams& F, int mode, cudaStream_t st);
bool pdl_enabled();                    // programmatic dependent launch between prep -> stream -> finalize (default on)
unsigned long long* debug_timebuf();   // GCA_TC_TIMEBUF env (bring-up only), nullptr normally

}  // namespace gca
