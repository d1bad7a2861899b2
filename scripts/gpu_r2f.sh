#!/bin/bash
O=gpurun_out/r2f
mkdir -p $O
GLASS_B200_TC_TIMELINE=1 timeout 300 python scripts/tc_timeline.py > $O/tc_timeline.log 2>&1
