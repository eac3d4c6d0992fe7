#!/usr/bin/env python
"""Bring-up aid (torchrun, >= 2 GPUs): %globaltimer timeline of one K-sharded head step over peer memory
(prep -> stream -> split merge + push -> cross-rank merge + enqueue).  Not part of the product or the tests."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-graph-ssl_b200"))
import torch
import torch.distributed as dist
import torch.nn.functional as F

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
buf = torch.zeros(32 * 1024 + 64, dtype=torch.int64, device=dev)
os.environ["GCA_TC_TIMEBUF"] = hex(buf.data_ptr())
from gca_b200.dist import ShardedRGBMoCo
from gca_b200.graphed import GraphedShardedStep
from gca_b200.peer import PeerShardLink

Bl = int(sys.argv[1]) if len(sys.argv) > 1 else 256
K = 1 << 20
torch.manual_seed(1)
moco = ShardedRGBMoCo(128, K=K, T=0.07, queue_dtype="bf16", device=dev)
link = PeerShardLink(Bl, 128, device=dev)
gs = GraphedShardedStep(moco, Bl, link=link).capture()
gs.q.copy_(F.normalize(torch.randn(Bl, 128)).to(dev)); gs.k.copy_(F.normalize(torch.randn(Bl, 128)).to(dev))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
sync = torch.zeros(1, device=dev)
BIG = 1 << 62
for it in range(30):
    flush.fill_(it & 1)
    dist.all_reduce(sync)
    if it == 29:
        buf.zero_()
        for w in (0, 2, 8):
            buf[32 * 1000 + w] = BIG
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); gs.graph.replay(); b.record()
torch.cuda.synchronize()
t = buf.cpu()
nblk = (2 * Bl * world // 128) * 0 + 148
cta = t[:32 * 148].view(148, 32).double()
live = cta[:, 0] > 0
p_in, p_out = float(t[32 * 1000 + 2]), float(t[32 * 1000 + 4])
s_in, s_out = float(cta[live, 0].min()), float(cta[live, 8].max())
f_in, f_out = float(t[32 * 1000 + 0]), float(t[32 * 1000 + 3])
m_in, m_out, m_last = float(t[32 * 1000 + 8]), float(t[32 * 1000 + 11]), float(t[32 * 1000 + 9])
rel = lambda x: (x - p_in) / 1e3
for r in range(world):
    dist.barrier()
    if r == rank:
        print("rank %d: event %.1f us | prep %.1f..%.1f | stream %.1f..%.1f (main loop done mean %.1f) | split merge+push %.1f..%.1f | "
              "cross-rank merge+enqueue %.1f..%.1f (ticket %.1f)" % (rank, a.elapsed_time(b) * 1e3, 0.0, rel(p_out), rel(s_in), rel(s_out),
              rel(float(cta[live, 5].mean())), rel(f_in), rel(f_out), rel(m_in), rel(m_out), rel(m_last)), flush=True)
dist.barrier()
os._exit(0)
