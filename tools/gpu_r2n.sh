#!/bin/bash
# 1-GPU: new tests (MLP / bn1d), whole suite
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest15.log 2>&1; tail -6 gpurun_out/r2_pytest15.log
