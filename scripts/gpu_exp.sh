#!/bin/bash
mkdir -p gpurun_out
for f in 0 1 2 3 7; do timeout -k 5 120 python scripts/tcn_clocks.py 256 20 $f 2>&1 | grep -v "^$"; done | tee gpurun_out/exp.txt
timeout -k 5 120 python scripts/tcn_clocks.py 59 20 0 2>&1 | grep -E "total|P2|A1" | tee -a gpurun_out/exp.txt
