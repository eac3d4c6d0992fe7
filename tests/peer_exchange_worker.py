"""Worker for the peer-memory key exchange check (gca_keys_exchange; one process per GPU, world size from the torchrun
env, 1 when launched bare).  Compares against NCCL all-gather bit for bit, eagerly and as a replayed CUDA graph with
rank-dependent delays so that ranks arrive at the exchange at different times; then times both variants."""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "video-graph-ssl_b200"))


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", str(29300 + os.getpid() % 500))
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from gca_b200.peer import PeerKeyExchange
    B, d = 256, 128
    ex = PeerKeyExchange(B, d, device=dev, timeout_ms=5000)
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    got, want = torch.empty(world * B, d, device=dev), torch.empty(world * B, d, device=dev)
    for step in range(9):                                        # odd count: the graph below starts on parity 1
        keys = torch.randn(B, d, device=dev, generator=gen)
        if step % 3 == rank % 3:
            torch.cuda._sleep(int(2e6) * (1 + rank))             # this rank shows up late
        ex(keys, got)
        dist.all_gather_into_tensor(want, keys)
        torch.cuda.synchronize()
        assert torch.equal(got, want), step
    assert ex.steps_done() == 9
    # captured: [optional delay] -> exchange, replayed with fresh keys
    keys = torch.randn(B, d, device=dev, generator=gen)
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ex(keys, got)                                            # warm-up outside capture (step 9)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ex(keys, got)
    for step in range(40):
        keys.copy_(torch.randn(B, d, device=dev, generator=gen))
        if (step + rank) % 4 == 0:
            torch.cuda._sleep(int(1e6))
        g.replay()
        dist.all_gather_into_tensor(want, keys)
        torch.cuda.synchronize()
        assert torch.equal(got, want), ("graph", step)
    ex.check()
    assert ex.steps_done() == 50
    # timing: replayed exchange vs replayed NCCL all-gather (device time per replay, back to back)
    gn = torch.cuda.CUDAGraph()
    dist.all_gather_into_tensor(want, keys)
    torch.cuda.synchronize()
    with torch.cuda.graph(gn):
        dist.all_gather_into_tensor(want, keys)
    res = {}
    for name, gr in (("p2p_kernel", g), ("nccl", gn)):
        for _ in range(20):
            gr.replay()
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(500):
            gr.replay()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / 500 * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = round(float(t), 2)
    ex.check()
    # the fused step (gca_moco_step_peer: push in the first launch, enqueue from the mailbox in the last) against the same
    # step with an NCCL all-gather: loss, gradient, ranks, queue and pointer must agree bit for bit on every rank
    import gca_b200
    from gca_b200.graphed import GraphedReplicaStep
    K = 8192
    mocos = []
    for _ in range(2):
        torch.manual_seed(5)
        mocos.append(gca_b200.RGBMoCo(d, K=K, T=0.07, queue_dtype="bf16").to(dev))
    for m in mocos:
        m.index = K - 3 * B * world + 64                             # wraps on the third step
    sa = GraphedReplicaStep(mocos[0], B).capture()                                   # NCCL
    sb = GraphedReplicaStep(mocos[1], B, exchange=ex, fuse_exchange=True).capture()  # fused peer exchange
    for step in range(6):
        q = torch.nn.functional.normalize(torch.randn(B, d, device=dev, generator=gen))
        k = torch.nn.functional.normalize(torch.randn(B, d, device=dev, generator=gen))
        if (step + rank) % 3 == 0:
            torch.cuda._sleep(int(1e6))
        la = sa.step(q, k).clone()
        lb = sb.step(q, k).clone()
        torch.cuda.synchronize()
        assert torch.equal(la, lb), ("fused loss", step, float(la), float(lb))
        assert torch.equal(sa.dq, sb.dq) and torch.equal(sa.rank, sb.rank) and torch.equal(sa.hits, sb.hits), step
        assert torch.equal(mocos[0].memory, mocos[1].memory), ("fused queue", step)
        assert int(sa.state[0]) == int(sb.state[0]) == mocos[0].index == mocos[1].index, step
    # back to back, no synchronisation between the steps (the launch-plan path chains consecutive steps by programmatic
    # dependent launch; the key push rides in the first launch with a late trigger), one rank delayed now and then: the
    # queue, the ring pointer and the last step's results must still equal the NCCL step's
    assert sb.plan is not None and sb.plan.launches == 3, "the fused replica step should run from a launch plan"
    qs = [torch.nn.functional.normalize(torch.randn(B, d, device=dev, generator=gen)) for _ in range(8)]
    ks = [torch.nn.functional.normalize(torch.randn(B, d, device=dev, generator=gen)) for _ in range(8)]
    for s_, tag in ((sa, "nccl"), (sb, "plan")):
        for step in range(40):
            if tag == "plan" and step % 7 == rank % 7:
                torch.cuda._sleep(int(3e5))
            s_.step(qs[step % 8], ks[step % 8])
        torch.cuda.synchronize()
    assert torch.equal(sa.loss, sb.loss) and torch.equal(sa.dq, sb.dq) and torch.equal(sa.hits, sb.hits), "back-to-back results"
    assert torch.equal(mocos[0].memory, mocos[1].memory), "back-to-back queue"
    assert int(sa.state[0]) == int(sb.state[0]) == mocos[0].index == mocos[1].index
    ex.check()
    dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        print("PEER_EXCHANGE_OK world=%d us_per_exchange=%s" % (world, res), flush=True)
    os._exit(0)


if __name__ == "__main__":
    main()
