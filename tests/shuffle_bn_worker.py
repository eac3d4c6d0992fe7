"""Worker for the ShuffleBN check over NCCL (torchrun --nproc-per-node 2; the fixture is for world size 2): the clip
exchange, the momentum-encoder forward on the shuffled rows, the key gather and the un-shuffle against
tests/golden/shuffle_bn_w2.npz, which was produced by EXECUTING the reference's own `Trainer._shuffle_bn`
(oracle/gen_golden_dist.py, tools/train_video_contrast_dis.py:189-231).  Also: the side-stream launch and the bf16 payload."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "video-graph-ssl_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", str(29600 + os.getpid() % 300))
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    g = dict(np.load(os.path.join(HERE, "golden", "shuffle_bn_w2.npz")))
    assert world == int(g["world"]), "the fixture was generated for world size %d" % int(g["world"])
    from gen_golden_dist import make_encoder                        # the fixture's encoder (test infrastructure)
    from gca_b200.dist import ShuffleBN
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    enc = make_encoder(int(g["encoder_seed"])).to(dev)
    sbn = ShuffleBN()
    for it in range(int(g["iters"])):
        x = torch.from_numpy(g["r%d_x%d" % (rank, it)]).to(dev)
        torch.manual_seed(7 + it + 50 * rank)                        # as in the generator: ranks draw different permutations
        state = torch.random.get_rng_state()
        k, all_k = sbn(x, enc)
        torch.cuda.synchronize()
        k_ref, all_ref = torch.from_numpy(g["r%d_k%d" % (rank, it)]), torch.from_numpy(g["r%d_all_k%d" % (rank, it)])
        # rows: which clip went where is integer work (exact); the values differ only by GPU vs CPU fp32 arithmetic of the encoder
        assert torch.allclose(all_k.cpu(), all_ref, rtol=1e-4, atol=1e-5), float((all_k.cpu() - all_ref).abs().max())
        assert torch.allclose(k.cpu(), k_ref, rtol=1e-4, atol=1e-5), float((k.cpu() - k_ref).abs().max())
        # the side-stream launch gives the same bits as the in-line call (same permutation: same generator state)
        torch.random.set_rng_state(state)
        h = sbn.launch(x, enc)
        busy = torch.randn(512, 512, device=dev) @ torch.randn(512, 512, device=dev)     # the caller's own work
        k2, all_k2 = h.wait()
        torch.cuda.synchronize()
        assert torch.equal(k2, k) and torch.equal(all_k2, all_k) and bool(torch.isfinite(busy).all())
        # bf16 payload: exactly the fp32 exchange of pre-rounded clips
        torch.random.set_rng_state(state)
        kb, all_kb = ShuffleBN(payload_dtype=torch.bfloat16)(x, enc)
        torch.random.set_rng_state(state)
        kr, all_kr = sbn(x.to(torch.bfloat16).float(), enc)
        torch.cuda.synchronize()
        assert torch.equal(kb, kr) and torch.equal(all_kb, all_kr)
    dist.barrier()
    if rank == 0:
        print("SHUFFLE_BN_NCCL_OK world=%d" % world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
