#!/bin/bash
timeout 120 python tools/tc_timeline.py 2>&1 | grep -v "cyc" | head -32
echo ---- NOPAIR
GCA_X_NOPAIR=1 timeout 120 python tools/tc_timeline.py 2>&1 | grep -v "cyc" | head -32
