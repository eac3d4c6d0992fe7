#!/bin/bash
for bl in 256 128; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29563 tools/shard_timeline.py $bl 2>&1 | grep "rank "
done
