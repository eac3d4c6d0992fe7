#!/usr/bin/env python
"""Pre-train step (BASELINE config 2: visual_moco.yaml shape) around the B200 head: clips/s with the reference's
eager head + per-tensor EMA loop against gca_b200's fused head + one-launch EMA, same backbone, same inputs.

The step is the reference's (tools/train_video_contrast_dis.py:398-440): chunk the 6-channel clip pair, ShuffleBN +
momentum encoder on x2 (no grad), encoder on x1, contrast + criterion, backward, SGD step, accuracy, loss.item(),
momentum update.  The backbone is the reference's R3D-18 re-stated (7x7x7 stem, 4 stages of 2 basic blocks, 512-d,
shortcut B; lib/modeling/backbone/backbone_3d/resnet.py:108-222 -- identical parameter shapes, 33.2 M parameters, and
feature map: tests/test_host_logic_cpu.py checks them against a list taken from the reference's own class) plus the MLP
head (project_head.py:13-33) running on cuDNN under bf16 autocast -- library code that the hot path leaves untouched; it
is here only as the surrounding workload.  Synthetic Kinetics-shaped clips randn(B, 6, 16, 112, 112), random-init weights.

    python tools/pretrain_step.py --batch 64 --steps 10 --warmup 3            # one GPU
    torchrun --nproc-per-node N ... tools/pretrain_step.py                    # DDP + ShuffleBN all-to-all + key gather
Prints one JSON line per head variant.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-graph-ssl_b200"))


class Block(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.c1 = nn.Conv3d(cin, cout, 3, stride, 1, bias=False)
        self.b1 = nn.BatchNorm3d(cout)
        self.c2 = nn.Conv3d(cout, cout, 3, 1, 1, bias=False)
        self.b2 = nn.BatchNorm3d(cout)
        self.down = None
        if stride != 1 or cin != cout:
            self.down = nn.Sequential(nn.Conv3d(cin, cout, 1, stride, bias=False), nn.BatchNorm3d(cout))

    def forward(self, x):
        r = x if self.down is None else self.down(x)
        y = F.relu(self.b1(self.c1(x)), inplace=True)
        return F.relu(self.b2(self.c2(y)) + r, inplace=True)


class R3D18Shape(nn.Module):
    """Encoder + MLP projection head -> 128-d unit rows (the `model` / `model_ema` of the trainer)."""

    def __init__(self, feat_dim=128):
        super().__init__()
        self.stem = nn.Sequential(nn.Conv3d(3, 64, 7, (1, 2, 2), 3, bias=False), nn.BatchNorm3d(64), nn.ReLU(inplace=True),
                                  nn.MaxPool3d(3, 2, 1))
        chans, layers, cin = [64, 128, 256, 512], [], 64
        for i, c in enumerate(chans):
            layers += [Block(cin, c, 1 if i == 0 else 2), Block(c, c, 1)]
            cin = c
        self.layers = nn.Sequential(*layers)
        self.head = nn.Sequential(nn.Linear(512, 512), nn.ReLU(inplace=True), nn.Linear(512, feat_dim))

    def forward(self, x):
        x = self.layers(self.stem(x))
        x = F.adaptive_avg_pool3d(x, 1).flatten(1)
        return F.normalize(self.head(x).float(), dim=1)


class EagerMoCo(nn.Module):
    """The reference head's op sequence in eager PyTorch on the GPU (mem_moco.py:60-88, mem_moco.py:8-34 enqueue,
    nce.py:58-66 loss, utils accuracy): what an unmodified checkout runs per step."""

    def __init__(self, n_dim, K, T):
        super().__init__()
        self.K, self.T, self.index = K, T, 0
        self.register_buffer("memory", F.normalize(torch.randn(K, n_dim)))

    def forward(self, q, k, all_k=None):
        k = k.detach()
        queue = self.memory.clone().detach()
        pos = (q * k).sum(1, keepdim=True)
        neg = q @ queue.t()
        out = torch.cat((pos, neg), dim=1) / self.T
        out = out.squeeze().contiguous()
        keys = all_k if all_k is not None else k
        with torch.no_grad():
            n = keys.shape[0]
            ids = (torch.arange(n, device=q.device) + self.index) % self.K
            self.memory.index_copy_(0, ids, keys)
            self.index = (self.index + n) % self.K
        return out, torch.zeros(q.shape[0], dtype=torch.long, device=q.device)


def eager_criterion(out):
    label = torch.zeros(out.shape[0], dtype=torch.long, device=out.device)
    return F.cross_entropy(out, label)


def accuracy(output, target, topk=(1, 5)):
    maxk = max(topk)
    _, pred = output.topk(maxk, 1, True, True)
    correct = pred.t().eq(target.view(1, -1).expand_as(pred.t()))
    return [correct[:k].reshape(-1).float().sum(0) * (100.0 / target.size(0)) for k in topk]


def loop_ema(model, model_ema, m):
    for p1, p2 in zip(model.parameters(), model_ema.parameters()):
        p2.data.mul_(m).add_(p1.detach().data, alpha=1 - m)


def run(variant, args, rank, world, dev):
    from gca_b200.memory import RGBMoCo, NCESoftmaxLoss
    from gca_b200.ema import MomentumUpdater
    from gca_b200.dist import ShuffleBN
    torch.manual_seed(1)
    model = R3D18Shape().to(dev).to(memory_format=torch.channels_last_3d)
    model_ema = R3D18Shape().to(dev).to(memory_format=torch.channels_last_3d)
    model_ema.load_state_dict(model.state_dict())
    for p in model_ema.parameters():
        p.requires_grad_(False)
    opt = torch.optim.SGD(model.parameters(), lr=0.06, momentum=0.9, weight_decay=5e-4)
    net = nn.parallel.DistributedDataParallel(model, device_ids=[dev.index]) if world > 1 else model
    if variant == "b200":
        contrast = RGBMoCo(128, K=args.K, T=0.07, queue_dtype="bf16").to(dev)
        criterion = NCESoftmaxLoss()
        ema = MomentumUpdater(model, model_ema)
        ema_step = lambda: ema.step(0.999)
    else:
        contrast = EagerMoCo(128, args.K, 0.07).to(dev)
        criterion = eager_criterion
        ema_step = lambda: loop_ema(model, model_ema, 0.999)
    sbn = ShuffleBN() if world > 1 else None
    model.train()
    model_ema.train()                                    # BN in train mode on the momentum encoder (train...:383-388)
    gen = torch.Generator(device=dev).manual_seed(7 + rank)
    images = torch.randn(args.batch, 6, 16, 112, 112, device=dev, generator=gen)
    head_ms = []

    def step(timed):
        x1, x2 = torch.chunk(images, 2, dim=1)
        if sbn is not None:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                feat_k, all_k = sbn(x2.contiguous(), model_ema)
        else:
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                feat_k = model_ema(x2)
            all_k = feat_k
        opt.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            feat_q = net(x1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        output, labels = contrast(feat_q, feat_k, all_k=all_k)
        loss = criterion(output)
        e1.record()
        loss.backward()
        opt.step()
        prec1, prec5 = accuracy(output.detach(), labels.detach(), topk=(1, 5))
        lv = loss.item()
        prec1.item(), prec5.item()
        ema_step()
        if timed:
            head_ms.append((e0, e1))
        return lv

    for _ in range(args.warmup):
        step(False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        lv = step(True)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    head = sorted(a.elapsed_time(b) for a, b in head_ms)[len(head_ms) // 2]
    # EMA alone
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        ema_step()
    b.record()
    torch.cuda.synchronize()
    res = {"workload": "pretrain step, R3D-18-shaped encoder, %d videos/GPU x 2 clips of 3x16x112x112, "
                       "queue %d x 128, bf16 autocast" % (args.batch, args.K),
           "variant": {"b200": "gca_b200 fused head (bf16 queue) + one-launch EMA",
                       "eager": "reference op sequence in eager PyTorch + per-tensor EMA loop"}[variant],
           "n_gpus": world, "ms_per_step": round(ms, 3),
           "clips_per_s": round(2 * args.batch * world / (ms * 1e-3), 1),
           "videos_per_s": round(args.batch * world / (ms * 1e-3), 1),
           "head_fwd_ms_median": round(head, 4), "ema_ms": round(a.elapsed_time(b) / 10, 4),
           "last_loss": round(lv, 5), "steps": args.steps, "warmup": args.warmup}
    if rank == 0 and not getattr(args, "quiet", False):
        print(json.dumps(res), flush=True)
    del net, model, model_ema, opt, contrast
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--K", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--variants", default="eager,b200")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.benchmark = True
    for v in args.variants.split(","):
        run(v, args, rank, world, dev)
        torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
