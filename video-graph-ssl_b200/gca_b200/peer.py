"""Key all-gather over NVLink peer memory (gca_keys_exchange), the NCCL-free variant of `Trainer._global_gather`
(tools/train_video_contrast_dis.py:182-187) for the momentum keys every replica enqueues (train...:222).

Each rank allocates one mailbox in symmetric memory (torch.distributed._symmetric_memory: a CUDA allocation every
peer of the node maps over NVLink); one kernel launch per step pushes the local keys into every peer's mailbox,
publishes a step-numbered flag, waits for the peers' flags and copies the received rows out.  The launch is
CUDA-graph capturable and carries no host synchronisation.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib
from ._lib import ptr


class PeerKeyExchange(object):
    """`ex(keys_local, all_k_out)` == `all_k_out[:] = torch.cat(all_gather(keys_local))`, stream-ordered.

    Collective constructor (every rank of `group`, same B and d).  All ranks must call the exchange the same number
    of times.  `timeout_ms` bounds the in-kernel wait for a peer (0 = wait for ever); `check()` raises if a wait
    ever expired."""

    def __init__(self, batch, n_dim, group=None, device=None, timeout_ms=10000):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.W, self.r = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.B, self.d = int(batch), int(n_dim)
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("PeerKeyExchange moves keys between GPUs; there is no CPU path")
        self.device = dev
        nbytes = int(_lib.load().gca_keys_exchange_bytes(self.B, self.d, self.W))
        self.mailbox = symm_mem.empty(nbytes, dtype=torch.uint8, device=dev)
        self.mailbox.zero_()
        self.handle = symm_mem.rendezvous(self.mailbox, self.group)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        if len(ptrs) != self.W or ptrs[self.r] != self.mailbox.data_ptr():
            raise RuntimeError("symmetric-memory rendezvous returned an unexpected peer table: %r" % (ptrs,))
        self.table = torch.tensor(ptrs, dtype=torch.int64, device=dev)
        self.xstate = torch.zeros(4, dtype=torch.int64, device=dev)
        self.timeout_ms = int(timeout_ms)
        torch.cuda.synchronize(dev)
        dist.barrier(group=self.group)                       # every mailbox is zeroed before anyone pushes

    def __call__(self, keys_local, all_k_out, stream=None):
        if keys_local.shape != (self.B, self.d) or all_k_out.shape != (self.W * self.B, self.d):
            raise ValueError("expected keys [%d, %d] and output [%d, %d]" % (self.B, self.d, self.W * self.B, self.d))
        if keys_local.dtype != torch.float32 or all_k_out.dtype != torch.float32:
            raise TypeError("keys travel as fp32")
        if not (keys_local.is_contiguous() and all_k_out.is_contiguous()):
            raise ValueError("buffers must be contiguous")
        if stream is None:
            stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.call("gca_keys_exchange", ptr(keys_local), self.B, self.d, self.W, self.r, ptr(self.table), ptr(all_k_out),
                  ptr(self.xstate), self.timeout_ms, stream)
        return all_k_out

    def steps_done(self):
        return int(self.xstate[0])

    def check(self):
        if int(self.xstate[2]) != 0:
            raise RuntimeError("gca_keys_exchange: a peer did not deliver its keys within %d ms" % self.timeout_ms)


class PeerShardLink(object):
    """Symmetric-memory mailbox + device state for the K-sharded head step over peer memory (gca_shard_step_peer): the q|k
    gather and the cross-rank merge of the partials both travel through it, no NCCL call on the data path.
    Collective constructor (every rank of `group`, same local batch and d); all ranks must issue the same step sequence."""

    def __init__(self, batch_local, n_dim, group=None, device=None, timeout_ms=10000):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.W, self.r = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.Bl, self.d = int(batch_local), int(n_dim)
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("PeerShardLink moves rows between GPUs; there is no CPU path")
        self.device = dev
        nbytes = int(_lib.load().gca_shard_peer_bytes(self.Bl, self.d, self.W))
        self.mailbox = symm_mem.empty(nbytes, dtype=torch.uint8, device=dev)
        self.mailbox.zero_()
        self.handle = symm_mem.rendezvous(self.mailbox, self.group)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        if len(ptrs) != self.W or ptrs[self.r] != self.mailbox.data_ptr():
            raise RuntimeError("symmetric-memory rendezvous returned an unexpected peer table: %r" % (ptrs,))
        self.table = torch.tensor(ptrs, dtype=torch.int64, device=dev)
        self.pstate = torch.zeros(8, dtype=torch.int64, device=dev)
        self.timeout_ms = int(timeout_ms)
        torch.cuda.synchronize(dev)
        dist.barrier(group=self.group)                       # every mailbox is zeroed before anyone pushes

    def check(self):
        st = self.pstate.cpu()
        if int(st[2]) != 0 or int(st[6]) != 0:
            raise RuntimeError("gca_shard_step_peer: a peer did not deliver within %d ms (gather flag %d, merge flag %d)"
                               % (self.timeout_ms, int(st[2]), int(st[6])))
