#!/bin/bash
# Developer tool (GPU box): the evidence committed under profiles/ -- GPU tests, the bench line, the reference arm,
# the ncu launch list of the device-resident bench and one `ncu --set full` capture of the three main kernels.
# Each ncu pass runs only after the same command has exited 0 without ncu.
set -x
R=${1:-r01}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${R}_pytest_gpu.txt
timeout 400 python bench.py > gpurun_out/${R}_bench_full.json 2> gpurun_out/${R}_bench_full.err || exit 1
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err
timeout 300 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/${R}_bench_device_only.json 2>/dev/null || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --no-e2e --no-cpu-baseline > gpurun_out/${R}_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:terse_encode|prolix_walk|prolix_unpack_seg' -c 3 \
    -o gpurun_out/${R}_full python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${R}_ncu_full.log 2>&1
timeout 300 python tools/bench_configs.py > gpurun_out/${R}_bench_configs.txt 2>&1
ls -la gpurun_out | tail -12
