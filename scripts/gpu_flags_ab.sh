#!/bin/bash
# A/B of the tile-level stage dependencies: bit-exactness against whole-launch dependencies, then one bench line per mode.
set -u
mkdir -p gpurun_out
timeout -k 5 100 python scripts/flags_check.py /tmp/a.npy 48 100000 && DXI_TCN_NO_FLAGS=1 timeout -k 5 100 python scripts/flags_check.py /tmp/b.npy 48 100000 && cmp /tmp/a.npy /tmp/b.npy && echo BITEXACT_BIG
timeout -k 5 60 python scripts/flags_check.py /tmp/c.npy 3 20000 same && DXI_TCN_NO_FLAGS=1 timeout -k 5 60 python scripts/flags_check.py /tmp/d.npy 3 20000 same && cmp /tmp/c.npy /tmp/d.npy && echo BITEXACT_SMALL
for mode in default NO_REVERSE NO_FLAGS ${EXTRA_MODES:-}; do
  env_kv=""; [ "$mode" != default ] && env_kv="DXI_TCN_$mode=1"
  env $env_kv timeout -k 5 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$mode.json 2> gpurun_out/bench_$mode.err
  python - "$mode" <<'PY'
import json, sys
m = sys.argv[1]
try:
    d = json.load(open('gpurun_out/bench_%s.json' % m))
    print(m, 'value %.0f  ms/step %.3f  e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']), {k: round(v, 3) for k, v in d['kernels_ms_per_step'].items()}, d['clocks'])
except Exception as e:
    print(m, 'bench parse failed', e); print(open('gpurun_out/bench_%s.err' % m).read()[-2000:])
PY
done
