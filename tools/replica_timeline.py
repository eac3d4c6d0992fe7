#!/usr/bin/env python
"""Bring-up aid (torchrun, >= 2 GPUs): %globaltimer timeline of one data-parallel head step with the key exchange fused over
peer memory (push on the side stream | prep -> stream -> finalize + enqueue from the mailbox).  Not part of the product."""
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
import gca_b200
from gca_b200.graphed import GraphedReplicaStep
from gca_b200.peer import PeerKeyExchange

B, K = 256, 65536
torch.manual_seed(1)
moco = gca_b200.RGBMoCo(128, K=K, queue_dtype="bf16").to(dev)
ex = PeerKeyExchange(B, 128, device=dev)
gs = GraphedReplicaStep(moco, B, exchange=ex, want_rank=False).capture()
torch.manual_seed(5 + rank)
gs.q.copy_(F.normalize(torch.randn(B, 128)).to(dev)); gs.k.copy_(F.normalize(torch.randn(B, 128)).to(dev))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
sync = torch.zeros(1, device=dev)
BIG = 1 << 62
for it in range(40):
    flush.fill_(it & 1)
    dist.all_reduce(sync)
    if it == 39:
        buf.zero_()
        for w in (0, 2):
            buf[32 * 1000 + w] = BIG
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); gs.graph.replay(); b.record()
torch.cuda.synchronize()
t = buf.cpu()
cta = t[:32 * 148].view(148, 32).double()
live = cta[:, 0] > 0
p_in, p_out = float(t[32 * 1000 + 2]), float(t[32 * 1000 + 4])
rel = lambda x: (float(x) - p_in) / 1e3
fb = t[32 * 400: 32 * 400 + 4 * 256].view(256, 4).double()
for r in range(world):
    dist.barrier()
    if r == rank:
        print("rank %d: event %.1f us | prep 0.0..%.1f | stream %.1f..%.1f (first S seen %.1f, main loop done %.1f) | finalize entry %.1f, "
              "rows done %.1f..%.1f, ticket %.1f | enqueue CTAs: flags seen %.1f, done %.1f" % (
                  rank, a.elapsed_time(b) * 1e3, rel(p_out), rel(cta[live, 0].min()), rel(cta[live, 8].max()), rel(cta[live, 4].mean()),
                  rel(cta[live, 5].mean()), rel(t[32 * 1000 + 0]), rel(fb[:, 2].min()), rel(fb[:, 2].max()), rel(t[32 * 1000 + 1]),
                  rel(t[32 * 1000 + 6]), rel(t[32 * 1000 + 5])), flush=True)
dist.barrier()
os._exit(0)
