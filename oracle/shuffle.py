"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's ShuffleBN bookkeeping, single process.

Follows Trainer._shuffle_bn (tools/train_video_contrast_dis.py:189-231) on a list of per-rank tensors: the
all_gather + cat becomes a torch.cat of the list, everything else is the reference's indexing, one rank at a time.
Only tests/ may import this.  Parity: pinned by construction (pure integer indexing of the reference's own
expressions, checked in tests against torch.randperm / argsort with the same generator state).
"""
import torch


def shuffle_bn_all_ranks(xs, encoder, shuffle_ids):
    """xs: list over ranks of [bsz, ...] tensors (one node).  Returns ([k_rank0, k_rank1, ...], all_k, [this_x...])."""
    world, bsz = len(xs), xs[0].shape[0]
    node_x = torch.cat(xs, dim=0)                                       # train...:203-206
    reverse_ids = torch.argsort(shuffle_ids)                            # :209
    this_xs, ks = [], []
    for r in range(world):
        this_ids = shuffle_ids[r * bsz:(r + 1) * bsz]                   # :213
        this_x = node_x[this_ids]                                       # :215
        this_xs.append(this_x)
        ks.append(encoder(this_x))                                      # :216-219
    all_k = torch.cat(ks, dim=0)                                        # :222 (_global_gather, :182-187)
    out = []
    for r in range(world):
        this_ids = reverse_ids[r * bsz:(r + 1) * bsz]                   # :228
        out.append(all_k[this_ids])                                     # :229 (single node: node_k == all_k)
    return out, all_k, this_xs
