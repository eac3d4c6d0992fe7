"""Worker for the GraphedShardedStep parity check (NCCL, one process per GPU; world size from the torchrun env, 1 when
launched bare).  Runs the same steps through the eager ShardedRGBMoCo path and through the captured graph and compares
every result bit for bit (both go through the same kernels in the same order).  Exits with os._exit: tearing down an
NCCL communicator that was captured into a CUDA graph can block."""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "video-graph-ssl_b200"))
sys.path.insert(0, os.path.dirname(HERE))                              # the oracle (checker only)


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", str(29400 + os.getpid() % 500))
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from gca_b200.dist import ShardedRGBMoCo
    from gca_b200.graphed import GraphedShardedStep
    from gca_b200.memory.losses import NCESoftmaxLoss
    from gca_b200.peer import PeerShardLink
    K, d, T, Bl = 8192 * world, 128, 0.07, 128
    link = PeerShardLink(Bl, d, device=dev)            # one mailbox for every peer-memory step of this process
    for qdt in ("bf16", "fp32"):
        torch.manual_seed(3)
        a = ShardedRGBMoCo(d, K=K, T=T, queue_dtype=qdt, device=dev)
        torch.manual_seed(3)
        b = ShardedRGBMoCo(d, K=K, T=T, queue_dtype=qdt, device=dev)
        assert torch.equal(a.memory, b.memory)
        a.index = b.index = K - 3 * Bl * world + 64                  # the third step wraps around the ring
        gs = GraphedShardedStep(b, Bl).capture()
        assert torch.equal(a.memory, b.memory) and b.index == a.index   # capture leaves queue and pointer untouched
        # third copy: the same step with the q|k gather and the cross-rank merge over NVLink peer memory (no NCCL call)
        torch.manual_seed(3)
        c = ShardedRGBMoCo(d, K=K, T=T, queue_dtype=qdt, device=dev)
        c.index = a.index
        gp = GraphedShardedStep(c, Bl, link=link).capture()
        assert torch.equal(a.memory, c.memory) and c.index == a.index
        crit = NCESoftmaxLoss()
        gen = torch.Generator().manual_seed(11 + rank)
        for step in range(4):
            q = torch.nn.functional.normalize(torch.randn(Bl, d, generator=gen)).to(dev)
            k = torch.nn.functional.normalize(torch.randn(Bl, d, generator=gen)).to(dev)
            qa = q.clone().requires_grad_(True)
            full_before = a.gather_full_queue().double().cpu()          # the queue one replica of the reference would hold
            out, _ = a(qa, k)
            la = crit(out)
            la.backward()
            lb = gs.step(q, k)
            torch.cuda.synchronize()
            assert torch.equal(la.detach().reshape(1), lb), (step, float(la), float(lb))
            assert torch.equal(qa.grad, gs.dq), step
            assert torch.equal(out.rank, gs.rank) and torch.equal(out.lse, gs.lse), step
            assert a.index == b.index and int(gs.state[0]) == b.index, step
            assert torch.equal(a.memory, b.memory), step
            # peer-memory variant: same kernels for the sweep, another (fixed) association order in the cross-rank merge
            lc = gp.step(q, k)
            torch.cuda.synchronize()
            link.check()
            assert abs(float(lc) - float(lb)) <= 2e-6 * abs(float(lb)), (step, float(lc), float(lb))
            assert float((gp.dq - gs.dq).abs().max()) <= 2e-5 * float(gs.dq.abs().max()), step
            assert float((gp.lse - gs.lse).abs().max()) <= 1e-5, step
            assert torch.equal(gp.rank, gs.rank), step
            assert c.index == b.index and int(gp.state[0]) == b.index, step
            assert torch.equal(c.memory, b.memory), step
            hits_ref = [int((gs.rank < 1).sum()), int((gs.rank < 5).sum())]
            assert gp.hits.tolist() == hits_ref, (step, gp.hits.tolist(), hits_ref)
            # ... and against the ORACLE: the reference step of this rank's rows on the full (replicated) queue
            import oracle
            if qdt == "bf16":
                c2 = torch.tensor(1.0 / T, dtype=torch.float32) * torch.tensor(1.4426950408889634, dtype=torch.float32)
                rq = (q.cpu() * c2).to(torch.bfloat16).double() / c2.double()
                ltol, gtol = 2e-3, 1e-2
            else:
                rq, ltol, gtol = q.cpu().double(), 1e-5, 1e-3
            o = oracle.infonce_step(rq, k.cpu().double(), full_before.clone(), 0, T)
            assert abs(float(lb) - float(o["loss"])) <= ltol * abs(float(o["loss"])), (step, float(lb), float(o["loss"]))
            err = float((gs.dq.cpu().double() - o["dq"]).abs().max() / o["dq"].abs().max())
            assert err <= gtol, (step, err)
            neg = oracle.logits_full(rq, k.cpu().double(), full_before, T)[:, 1:]
            pos = ((q.cpu().double() * k.cpu().double()).sum(1) / T)[:, None]   # the kernels take the positive from the fp32 q.k
            ok = (neg - pos).abs().min(1).values > 3e-4
            assert torch.equal(gs.rank.cpu().long()[ok], (neg > pos).sum(1)[ok]), step
            assert abs(float(lc) - float(o["loss"])) <= ltol * abs(float(o["loss"])), (step, float(lc), float(o["loss"]))
            err = float((gp.dq.cpu().double() - o["dq"]).abs().max() / o["dq"].abs().max())
            assert err <= gtol, (step, err)
        assert gs.launches_per_step >= 5
    assert gp.launches_per_step in (4, 6), gp.launches_per_step      # bf16: 4; fp32 (tc32 + stand-alone gather): 6
    dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        print("SHARDED_GRAPH_OK world=%d launches_per_step=%d peer_launches_per_step=%d" % (world, gs.launches_per_step, gp.launches_per_step), flush=True)
    os._exit(0)


if __name__ == "__main__":
    main()
