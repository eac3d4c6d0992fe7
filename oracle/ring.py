"""Oracle: MoCo ring-buffer bookkeeping (integer work, bit-exact).  TEST INFRASTRUCTURE ONLY.

Restates lib/memory/mem_moco.py:14-27 (`_update_pointer`, `_update_memory`).
"""
import numpy as np
import torch


def ring_slots(index: int, n: int, K: int) -> np.ndarray:
    """Slot ids written by one enqueue of `n` rows starting at pointer `index`.

    mem_moco.py:24-26: `fmod(arange(n) + index, K).long()` -- a float fmod of small
    non-negative integers, i.e. plain integer modulo.
    """
    return (np.arange(n, dtype=np.int64) + int(index)) % int(K)


def advance_pointer(index: int, n: int, K: int) -> int:
    """mem_moco.py:14-15."""
    return (int(index) + int(n)) % int(K)


def enqueue(memory: torch.Tensor, keys: torch.Tensor, index: int) -> int:
    """In-place enqueue of `keys[n, d]` into the full queue `memory[K, d]`; returns the new pointer.

    mem_moco.py:17-27 (`index_copy_` on the ring slots) then :83 (pointer update).  `keys` are cast
    to memory's dtype with round-to-nearest-even (bf16 queue mode; a no-op for the fp32 queue).
    n > K makes the reference's `index_copy_` write duplicate indices (undefined order): rejected.
    """
    K, n = memory.shape[0], keys.shape[0]
    if n > K:
        raise ValueError("enqueue of more rows than the queue holds is undefined in the reference")
    memory.index_copy_(0, torch.from_numpy(ring_slots(index, n, K)), keys.detach().to(memory.dtype))
    return advance_pointer(index, n, K)


def enqueue_sharded(shard: torch.Tensor, keys: torch.Tensor, index: int, K: int, k_begin: int) -> int:
    """`enqueue` for a rank owning global slots [k_begin, k_begin + len(shard)) of a ring of size K.

    Every rank sees the same gathered `keys` (train_video_contrast_dis.py:222, mem_moco.py:81-83);
    with the queue split along K a rank writes only the slots it owns.  The pointer
    is global and advances identically on every rank.
    """
    n = keys.shape[0]
    if n > K:
        raise ValueError("enqueue of more rows than the queue holds is undefined in the reference")
    slots = ring_slots(index, n, K)
    own = (slots >= k_begin) & (slots < k_begin + shard.shape[0])
    if own.any():
        rows = torch.from_numpy(np.nonzero(own)[0])
        shard.index_copy_(0, torch.from_numpy(slots[own] - k_begin), keys.detach()[rows].to(shard.dtype))
    return advance_pointer(index, n, K)
