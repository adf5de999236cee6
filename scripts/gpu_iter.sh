#!/bin/bash
# Quick iteration pass: parity tests (optional -k filter in $1), stage-kernel phase clocks, one bench line.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -k 5 ${TEST_TIMEOUT:-300} python -m pytest tests/test_gpu_network.py tests/test_gpu_signal.py -q -x --durations=5 ${1:+-k "$1"} > gpurun_out/tests.log 2>&1 ; echo "tests rc=$?" ; tail -${TEST_TAIL:-12} gpurun_out/tests.log
for f in ${CLOCK_FLAGS:-0}; do timeout -k 5 120 python scripts/tcn_clocks.py ${CLOCK_B:-256} 20 $f 2>&1 | grep -v "^$"; done > gpurun_out/clocks.txt 2>&1; cat gpurun_out/clocks.txt
timeout -k 5 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err ; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_iter.json'))
    print('value %.0f audio-s/s  ms/step %.3f  e2e %.0f  launches %d'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches']))
    print('kernels ms/step', {k:round(v,3) for k,v in d['kernels_ms_per_step'].items()})
    print('roofline', d['roofline']['achieved'], d['roofline']['frac'], 'stft', d['stft']['achieved'], d['stft']['frac'], 'enh GB/s', d['enhance_gbs'])
    print('clocks', d['clocks'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_iter.err').read()[-3000:])
PY
