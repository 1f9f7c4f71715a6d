#!/bin/bash
# Developer tool (GPU box): run bench.py (and, with AB_CONFIGS=1, tools/bench_configs.py) once per variant built by
# tools/ab_build.py.
cp trpx_b200/libtrpx_b200.so /tmp/lib_orig.so
for v in "$@"; do
  cp trpx_b200/_variants/lib_$v.so trpx_b200/libtrpx_b200.so
  timeout 300 python bench.py --steps 3 --no-e2e --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/ab_$v.log 2>&1
  python -c "
import json; d=json.loads(open('gpurun_out/ab_$v.log').read().strip().splitlines()[-1]); print('$v', {k:round(v,3) for k,v in d['kernel_ms'].items()})"
  if [ -n "$AB_CONFIGS" ]; then timeout 600 python tools/bench_configs.py 2>&1 | grep -v "^ " | tail -6 | cut -c1-125; fi
done
cp /tmp/lib_orig.so trpx_b200/libtrpx_b200.so
