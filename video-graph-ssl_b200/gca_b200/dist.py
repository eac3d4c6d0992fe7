"""MoCo queue sharded along K over a torch.distributed process group (one process per GPU, NCCL over NVLink).

The reference replicates the queue on every rank: it is broadcast once (train_video_contrast_dis.py:233-242) and
every rank enqueues the same gathered keys (mem_moco.py:81-83, train...:222), so all replicas stay identical.
Sharding therefore changes no observable result: rank r owns global slots [r*K/W, (r+1)*K/W), every rank scores
ALL gathered query rows against its shard, and the online-softmax partials are merged with two small collectives
(SURVEY.md section 8e, Appendix A.4):

    all-gather q, k                      [B_loc, d] -> [B_glob, d]       (the key gather the reference already does)
    shard kernel                         (max, sum-exp, count, acc) of every row against the local shard
    all-gather of the [3, B_glob] stats  12*B_glob bytes per rank
    combine kernel                       lse, loss rows, rank; acc *= exp(max_r - lse)
    reduce-scatter(sum) of acc           [B_glob, d] -> [B_loc, d]
    finish kernel                        dq_unit, mean loss over the local rows (DDP averages across ranks)

Compute goes through `ShardCompute` (the C ABI).  The collectives are plain torch.distributed calls; on the gloo
backend (CPU tests) reduce-scatter is emulated with all-reduce + slice.
"""
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from . import functional as GF
from ._lib import ptr
from .memory.moco_queue import FusedLogits


class ShardCompute(object):
    """The four per-rank compute steps on the CUDA library (include/gca_b200.h, 'sharded' section)."""

    def shard_fwd(self, q_all, k_all, shard, T, algo, want_grad):
        GF._need_cuda(q_all, k_all, shard)
        B, d = q_all.shape
        dev = q_all.device
        qd = GF.queue_dtype_code(shard)
        pos = torch.empty(B, dtype=torch.float32, device=dev)
        stats = torch.empty(3, B, dtype=torch.float32, device=dev)          # max, sum, count (int32 bits)
        acc = torch.empty(B, d, dtype=torch.float32, device=dev) if want_grad else None
        ws = GF.workspace(dev, GF.infonce_workspace_bytes(B, shard.shape[0], d, qd, algo), "infonce")
        _lib.call("gca_infonce_shard_fwd", ptr(q_all), ptr(k_all), ptr(shard), qd, B, shard.shape[0], d, 1.0 / T,
                  _lib.ALGO[algo], ptr(pos), ptr(stats[0]), ptr(stats[1]), ptr(stats[2]), ptr(acc), ptr(ws), ws.numel(),
                  GF._stream(q_all))
        return pos, stats, acc

    def shard_combine(self, all_stats, rank_id, pos, acc):
        W, _, B = all_stats.shape
        dev = pos.device
        lse = torch.empty(B, dtype=torch.float32, device=dev)
        loss_rows = torch.empty(B, dtype=torch.float32, device=dev)
        rank_gt = torch.empty(B, dtype=torch.int32, device=dev)
        st = all_stats.permute(1, 0, 2).contiguous()                        # [3, W, B]
        d = acc.shape[1] if acc is not None else 1
        _lib.call("gca_infonce_shard_combine", ptr(st[0]), ptr(st[1]), ptr(st[2]), W, rank_id, B, d, ptr(pos), ptr(lse),
                  ptr(loss_rows), ptr(rank_gt), ptr(acc), GF._stream(pos))
        return lse, loss_rows, rank_gt

    def shard_finish(self, acc_loc, k_loc, pos_loc, lse_loc, loss_rows_loc, T):
        B_loc = pos_loc.shape[0]
        dev = pos_loc.device
        d = k_loc.shape[1]
        dq_unit = torch.empty(B_loc, d, dtype=torch.float32, device=dev) if acc_loc is not None else None
        loss = torch.empty((), dtype=torch.float32, device=dev)
        _lib.call("gca_infonce_shard_finish", ptr(acc_loc), ptr(k_loc), ptr(pos_loc), ptr(lse_loc), ptr(loss_rows_loc),
                  B_loc, d, 1.0 / T, ptr(dq_unit), ptr(loss), GF._stream(pos_loc))
        return dq_unit, loss

    def enqueue(self, shard, keys, index, K, k_begin):
        return GF.enqueue_(shard, keys, index, K_global=K, k_begin=k_begin)


def _all_gather_rows(x, group):
    W = dist.get_world_size(group)
    out = torch.empty((W * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


def _reduce_scatter_rows(x, group):
    """[W*n, d] summed over ranks, rank r keeps rows [r*n, (r+1)*n)."""
    W, r = dist.get_world_size(group), dist.get_rank(group)
    n = x.shape[0] // W
    if dist.get_backend(group) == "gloo":                   # gloo has no reduce-scatter
        y = x.clone()
        dist.all_reduce(y, group=group)
        return y[r * n:(r + 1) * n].contiguous()
    out = torch.empty((n,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.reduce_scatter_tensor(out, x.contiguous(), group=group)
    return out


class _ShardedInfoNCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, shard, owner):
        """q, k: the LOCAL rows, k in the same (un-shuffled) order as q: row b of k is the positive of row b of q
        (mem_moco.py:36-38).  Both are gathered here; a caller-supplied `all_k` is never used for the positives -- the one
        `_shuffle_bn` returns is in SHUFFLED row order (train_video_contrast_dis.py:222, 231) and only feeds the enqueue."""
        group, T, comp = owner.group, owner.T, owner.compute
        W, r = dist.get_world_size(group), dist.get_rank(group)
        want = ctx.needs_input_grad[0]
        B_loc = q.shape[0]
        q32 = q.detach().float().contiguous()
        q_all = _all_gather_rows(q32, group)
        k_all = _all_gather_rows(k.detach().float().contiguous(), group)
        pos, stats, acc = comp.shard_fwd(q_all, k_all, shard, T, owner.algo, want)
        all_stats = _all_gather_rows(stats.unsqueeze(0), group)              # [W, 3, B_glob]
        lse, loss_rows, rank_gt = comp.shard_combine(all_stats, r, pos, acc)
        acc_loc = _reduce_scatter_rows(acc, group) if want else None
        sl = slice(r * B_loc, (r + 1) * B_loc)
        dq_unit, loss = comp.shard_finish(acc_loc, k_all[sl].contiguous(), pos[sl].contiguous(), lse[sl].contiguous(),
                                          loss_rows[sl].contiguous(), T)
        if want:
            ctx.save_for_backward(dq_unit)
        ctx.in_dtype = q.dtype
        outs = (loss, loss_rows[sl].contiguous(), lse[sl].contiguous(), pos[sl].contiguous(), rank_gt[sl].contiguous(),
                k_all)
        ctx.mark_non_differentiable(*outs[1:])
        return outs

    @staticmethod
    def backward(ctx, g_loss, *_):
        (dq_unit,) = ctx.saved_tensors
        return (dq_unit * g_loss).to(ctx.in_dtype), None, None, None


class ShardedRGBMoCo(nn.Module):
    """`RGBMoCo` with the queue split along K over `group`.  `forward(q, k, all_k=None)` takes the LOCAL q, k
    (and optionally the already gathered keys, as the trainer passes them) and returns (FusedLogits, labels) for the
    local rows; the loss is the mean over local rows, exactly what each DDP replica of the reference computes."""

    def __init__(self, n_dim, K=65536, T=0.07, group=None, queue_dtype="fp32", algo="auto", compute=None, device=None):
        super(ShardedRGBMoCo, self).__init__()
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if K % self.world != 0:
            raise ValueError("K=%d must divide evenly over %d ranks" % (K, self.world))
        self.K, self.T, self.index = K, T, 0
        self.algo = algo
        self.compute = compute if compute is not None else ShardCompute()
        self.k_begin = self.rank * (K // self.world)
        # every rank draws the full queue like the reference (mem_moco.py:57-58); rank 0's draw wins
        # (train_video_contrast_dis.py:233-242), then each rank keeps only the slots it owns
        full = F.normalize(torch.randn(K, n_dim))
        if device is not None:
            full = full.to(device)
        dist.broadcast(full, 0, group=self.group)
        shard = full[self.k_begin:self.k_begin + K // self.world].clone()
        self.register_buffer('memory', shard.to(torch.bfloat16) if queue_dtype == "bf16" else shard)

    def forward(self, q, k, all_k=None):
        k = k.detach()
        if all_k is not None and all_k.shape[0] != self.world * q.shape[0]:
            raise ValueError("all_k must hold the keys of every rank (%d rows), got %d" % (self.world * q.shape[0], all_k.shape[0]))
        # positives: the local k, gathered in rank order (row b of k belongs to row b of q).  Enqueue: the caller's all_k in
        # ITS row order when given (the trainer's is in ShuffleBN order, train...:222 -- every replica of the reference
        # enqueues exactly those rows, mem_moco.py:81-82), else the rank-ordered gather.
        loss, loss_rows, lse, pos, rank_gt, k_all = _ShardedInfoNCE.apply(q, k, self.memory, self)
        rows = all_k.detach().float() if all_k is not None else k_all
        with torch.no_grad():
            self.index = self.compute.enqueue(self.memory, rows, self.index, self.K, self.k_begin)
        labels = torch.zeros(q.shape[0], dtype=torch.long, device=q.device)
        return FusedLogits(loss, loss_rows, lse, pos, rank_gt, q.shape[0], self.K), labels

    def gather_full_queue(self):
        """The whole [K, d] queue in fp32 (checkpoint format of the reference, train...:278)."""
        return _all_gather_rows(self.memory.float(), self.group)

    # ---- checkpoints -------------------------------------------------------------------------------------------
    # `state_dict()` stays local and non-collective (the trainer saves on rank 0 only, train...:270-285): it holds this
    # rank's shard, labelled with the slots it covers ('shard_begin', 'shard_world') so that it can only be loaded back into
    # the rank that owns them -- an unmodified trainer that saves rank 0's state_dict() and resumes on every rank gets an
    # error instead of W copies of rank 0's shard.  `full_state_dict()` is the upstream format -- one fp32 [K, d] `memory` --
    # and is COLLECTIVE: every rank calls it, any rank may write the result (INTEGRATION.md shows the trainer change).
    # `load_state_dict` takes either; a full queue is cut down to the owned slots.
    def full_state_dict(self, include_pointer=False):
        """{'memory': fp32 [K, d]} exactly as `RGBMoCo.state_dict()` / the reference would save it.  With
        include_pointer=True the ring pointer travels as 'index' (the reference drops it on resume)."""
        sd = {"memory": self.gather_full_queue().cpu()}
        if include_pointer:
            sd["index"] = torch.tensor(int(self.index), dtype=torch.int64)
        return sd

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        super(ShardedRGBMoCo, self)._save_to_state_dict(destination, prefix, keep_vars)
        if self.memory.dtype == torch.bfloat16:
            destination[prefix + "memory"] = self.memory.float()
        if self.world > 1:
            destination[prefix + "shard_begin"] = torch.tensor(int(self.k_begin), dtype=torch.int64)
            destination[prefix + "shard_world"] = torch.tensor(int(self.world), dtype=torch.int64)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        key = prefix + "memory"
        begin = state_dict.pop(prefix + "shard_begin", None)
        sworld = state_dict.pop(prefix + "shard_world", None)
        if key in state_dict:
            mem = state_dict[key]
            Ks = self.memory.shape[0]
            if mem.shape[0] == self.K and self.world > 1:            # a full (upstream-format) queue: keep the owned slots
                mem = mem[self.k_begin:self.k_begin + Ks]
            elif mem.shape[0] != Ks:
                raise ValueError("checkpoint queue has %d rows; expected the full %d or this rank's %d"
                                 % (mem.shape[0], self.K, Ks))
            elif self.world > 1:
                # a per-rank shard: only the rank that owns those slots may take it
                if begin is None or sworld is None:
                    raise ValueError("checkpoint holds a %d-row queue shard without its slot range; save the sharded queue "
                                     "with full_state_dict() (collective) or state_dict() of this module" % Ks)
                if int(begin) != self.k_begin or int(sworld) != self.world:
                    raise ValueError("checkpoint shard covers slots from %d (world %d); this rank owns slots from %d (world "
                                     "%d) -- a sharded queue must be saved with full_state_dict() (collective) or per rank"
                                     % (int(begin), int(sworld), self.k_begin, self.world))
            state_dict[key] = mem.to(self.memory.dtype)
        if prefix + "index" in state_dict:
            self.index = int(state_dict.pop(prefix + "index")) % self.K
        super(ShardedRGBMoCo, self)._load_from_state_dict(state_dict, prefix, *args, **kwargs)


# ---------------------------------------------------------------------------------------------------------------
# ShuffleBN clip exchange (train_video_contrast_dis.py:189-231)
# ---------------------------------------------------------------------------------------------------------------

def shuffle_plan(shuffle_ids, bsz, world, rank):
    """Host-side routing tables for one rank of the ShuffleBN exchange.

    The reference all-gathers every rank's clips and then indexes `node_x[shuffle_ids[rank*bsz:(rank+1)*bsz]]`
    (train...:203-218), moving world x the payload.  Row g of `node_x` lives on rank g // bsz as local row g % bsz, so
    only the rows a peer asks for need to travel.  Returns (all int64, CPU):
      send_rows   local row numbers in send order: grouped by destination rank, inside a group in the order the
                  destination's id list names them
      send_counts rows per destination                     [world]
      recv_counts rows per source                          [world]
      recv_place  recv_place[j] = position in `this_x` of the j-th received row (receive order: by source rank)
    """
    ids = shuffle_ids.to("cpu", torch.int64).view(world, bsz)
    src = torch.div(ids, bsz, rounding_mode="floor")
    send_rows, send_counts = [], []
    for dst in range(world):
        mine = ids[dst][src[dst] == rank] - rank * bsz
        send_rows.append(mine)
        send_counts.append(int(mine.numel()))
    order = torch.argsort(src[rank], stable=True)
    recv_counts = torch.bincount(src[rank], minlength=world).tolist()
    return torch.cat(send_rows), send_counts, recv_counts, order


class ShuffleBN(object):
    """`Trainer._shuffle_bn` (train_video_contrast_dis.py:189-231) with an all-to-all instead of gather-everything.

    `__call__(x, model_ema)` returns `(k, all_k)` exactly like the reference: x is shuffled across the ranks of
    `local_group` with a permutation drawn by `torch.randperm` on every rank and overwritten by rank 0's
    (train...:208-211), the momentum encoder runs on the shuffled rows, the keys are gathered over `global_group`
    and un-shuffled.  The permutation, its inverse and every row placement are integer work: bit-identical to
    the reference for the same generator state.
    """

    def __init__(self, local_group=None, global_group=None, node_rank=0, payload_dtype=None):
        """payload_dtype: None -> the clips travel in their own dtype (bit-identical to the reference's gather + index);
        torch.bfloat16 -> rows are rounded to bf16 for the NVLink hop and widened again on arrival (half the bytes of the
        largest exchange of the step; the momentum encoder then sees bf16-rounded pixels -- opt-in, it changes bits)."""
        self.local_group = local_group if local_group is not None else dist.group.WORLD
        self.global_group = global_group if global_group is not None else dist.group.WORLD
        self.node_rank = node_rank
        self.payload_dtype = payload_dtype
        self.last_shuffle_ids = None
        self._side = None

    def exchange(self, x, shuffle_ids):
        """Rows `shuffle_ids[rank*bsz:(rank+1)*bsz]` of the (virtual) concatenation of every rank's x."""
        g = self.local_group
        W, r = dist.get_world_size(g), dist.get_rank(g)
        bsz = x.shape[0]
        send_rows, send_counts, recv_counts, place = shuffle_plan(shuffle_ids, bsz, W, r)
        x2 = x.contiguous().view(bsz, -1)
        send = x2.index_select(0, send_rows.to(x.device))
        if self.payload_dtype is not None:
            send = send.to(self.payload_dtype)
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, output_split_sizes=recv_counts, input_split_sizes=send_counts, group=g)
        out = torch.empty_like(x2)
        out.index_copy_(0, place.to(x.device), recv.to(x2.dtype))
        return out.view(x.shape)

    def __call__(self, x, model_ema):
        g = self.local_group
        W, r = dist.get_world_size(g), dist.get_rank(g)
        bsz = x.shape[0]
        shuffle_ids = torch.randperm(bsz * W).to(x.device)                 # train...:208 (CPU generator, as upstream)
        gg = self.global_group
        root = 0 if gg is dist.group.WORLD else dist.get_global_rank(gg, 0)
        dist.broadcast(shuffle_ids, root, group=gg)                         # rank 0's draw wins (train...:210-211)
        # ONE broadcast: the inverse permutation is a function of the permutation, so every rank sorts rank 0's ids itself
        # (integer work, identical everywhere) instead of receiving rank 0's argsort in a second collective (train...:209, 211)
        reverse_ids = torch.argsort(shuffle_ids)
        self.last_shuffle_ids = shuffle_ids
        with torch.no_grad():
            this_x = self.exchange(x, shuffle_ids)
            k = model_ema(this_x)
        all_k = _all_gather_rows(k.contiguous(), self.global_group)
        node_k = all_k[self.node_rank * W * bsz:(self.node_rank + 1) * W * bsz]
        return node_k[reverse_ids[r * bsz:(r + 1) * bsz]], all_k

    def launch(self, x, model_ema):
        """The whole key branch -- clip exchange, momentum-encoder forward, key gather, un-shuffle -- on a side CUDA stream,
        so that it overlaps the caller's `model(x1)` on the current stream: nothing in it depends on the online encoder
        (train...:407-410 run the two branches back to back).  Returns a handle; `handle.wait()` orders the current stream
        after the branch and gives `(k, all_k)`."""
        dev = x.device
        if self._side is None:
            self._side = torch.cuda.Stream(dev)
        cur = torch.cuda.current_stream(dev)
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            k, all_k = self(x, model_ema)
            done = torch.cuda.Event()
            done.record(self._side)
        x.record_stream(self._side)
        return _PendingKeys(k, all_k, done, dev)


class _PendingKeys(object):
    def __init__(self, k, all_k, done, dev):
        self.k, self.all_k, self.done, self.dev = k, all_k, done, dev

    def wait(self):
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(self.done)
        self.k.record_stream(cur)
        self.all_k.record_stream(cur)
        return self.k, self.all_k
