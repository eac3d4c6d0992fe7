#!/bin/bash
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29563 tools/replica_timeline.py 2>&1 | grep "rank \|Error\|error" | head
timeout 120 python tools/step_timeline.py 2>&1 | tail -4
