#!/bin/bash
for a in 0 1; do for b in 0 1; do GCA_GRAPH_NOSLAB=$a GCA_GRAPH_NOREV=$b timeout 100 python tools/graph_bwd_ab.py 2>&1 | tail -2; done; done
