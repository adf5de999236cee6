#!/bin/bash
# bench.py under a few DXI_TCN_* settings (tuning aid)
for cfg in "DXI_TCN_DBGFLAGS=0" "DXI_TCN_DBGFLAGS=16" "DXI_TCN_CHAIN=0" "DXI_TCN_CHAIN=0 DXI_TCN_DBGFLAGS=16" "DXI_TCN_DBGFLAGS=0" "DXI_TCN_DBGFLAGS=16"; do
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$cfg', 'ms/step %.3f' % d['ms_per_step'], {k:round(v,3) for k,v in d['kernels_ms_per_step'].items()}, 'e2e %.0f' % d['e2e']['value'])"
done
