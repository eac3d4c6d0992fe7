#!/bin/bash
for r in 8 4 2 1; do echo "== prep rows per CTA $r"; GCA_PREP_ROWS=$r HOSTIO=1 timeout 120 python tools/step_timeline.py 2>&1 | grep -E "warm L2|q block seen|prep last" | tail -3; done
