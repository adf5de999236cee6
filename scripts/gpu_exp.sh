#!/bin/bash
mkdir -p gpurun_out
for f in 0 1 2 3 4 7; do timeout -k 5 120 python scripts/tcn_clocks.py 256 20 $f 2>&1 | grep -v "^$"; done | tee gpurun_out/exp.txt
