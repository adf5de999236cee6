#!/bin/bash
mkdir -p gpurun_out
for b in 29 59 118 256; do timeout -k 5 120 python scripts/tcn_clocks.py $b 20 0 2>&1 | grep -E "total|P2|A1 next|gap"; done | tee gpurun_out/exp.txt
for u in 59 118 256 512; do timeout -k 5 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --utts $u > gpurun_out/b_$u.json 2>gpurun_out/b_$u.err; python -c "
import json; d=json.load(open('gpurun_out/b_$u.json')); print($u, 'value %.0f ms/step %.3f e2e %.0f'%(d['value'], d['ms_per_step'], d['e2e']['value']), {k:round(v,3) for k,v in d['kernels_ms_per_step'].items()})"; done | tee -a gpurun_out/exp.txt
