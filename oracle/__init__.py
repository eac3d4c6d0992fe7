"""CPU oracle for the GCA contrastive hot path.  TEST INFRASTRUCTURE ONLY.

Every function in this package is a CPU restatement (torch / numpy, fp32 or fp64) of the
reference's algorithm for one piece of the hot path, citing the reference file:line it follows
(paths relative to the upstream repo ACMMM2021-Anonymous/video-graph-ssl).

Who may import this package: `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs, and there only as the checker / the timed CPU baseline.
The product (`video-graph-ssl_b200/`) never imports it and has no CPU fallback.

Parity status: PINNED.  The reference is pure Python and imports in the build container, so
`oracle/gen_golden.py` runs the reference's own modules (`lib.memory`, `lib.ops`,
`lib.evaluation`, `tools/video_retrieval.py` arithmetic) on seeded inputs, checks every
restatement here against them (bit-equal where stated) and writes the outputs to
`tests/golden/*.npz`; `tests/test_oracle_golden.py` re-checks the restatements against those
fixtures on every run (the reference itself does not exist on the GPU box).
The reference ships no tests or golden vectors of its own.
"""
from .ring import ring_slots, enqueue, advance_pointer                      # noqa: F401
from .infonce import (logits_full, infonce_loss, infonce_grad_q, infonce_step,   # noqa: F401
                      lse_partials, merge_partials, positive_rank, topk_accuracy,
                      sharded_infonce)
from .graph import hop_distance, hop_weights, graph_forward, graph_forward_backward  # noqa: F401
from .negcos import neg_cosine, neg_cosine_grad                               # noqa: F401
from .retrieval import cosine_topk, recall_hits                               # noqa: F401
from . import bank                                                            # noqa: F401  (NPID instance bank: mem_bank.py)
