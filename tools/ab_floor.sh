#!/bin/bash
# A/B the kernel variants under build/variants with the synthetic step-floor schedules and the bench
for lib in build/variants/lib_*.so; do
  echo "== $lib"
  TEBSCAT_LIB=$PWD/$lib TEBSCAT_LOCAL_MAX_SLOTS=0 python tools/step_floor.py 2>&1 | grep -E "16 warps, one R16 butterfly each|two R8|R16 unit"
  TEBSCAT_LIB=$PWD/$lib TEBSCAT_LOCAL_MAX_SLOTS=0 python bench.py --steps 3 --warmup 3 --no-phase 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', round(d['value']), d['ms_per_step'])"
done
