#!/bin/bash
# Developer tool (GPU box): the evidence committed under profiles/.  One ncu use per invocation, and each ncu pass runs
# only after the same command has exited 0 without ncu.
#   final_profiles.sh r02 a   GPU tests, the bench lines of every config, the reference arm, the foreign-stack probe and
#                             the ncu launch list of the device-resident bench
#   final_profiles.sh r02 b   one `ncu --set full` capture of the three main kernels
set -x
R=${1:-r02}
PART=${2:-a}
if [ "$PART" = a ]; then
    timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${R}_pytest_gpu.txt
    timeout 600 python bench.py > gpurun_out/${R}_bench_full.json 2> gpurun_out/${R}_bench_full.err || exit 1
    timeout 400 python bench.py --impl reference > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err
    for c in c3 c4u8 c4u16 c5i16 c5i32; do
        timeout 600 python bench.py --config $c > gpurun_out/${R}_bench_$c.json 2> gpurun_out/${R}_bench_$c.err
    done
    timeout 300 python tools/foreign_probe.py 2000 > gpurun_out/${R}_foreign_stack.txt 2>&1
    timeout 300 python tools/foreign_probe.py 10000 >> gpurun_out/${R}_foreign_stack.txt 2>&1
    timeout 300 python tools/latency_probe.py > gpurun_out/${R}_latency.txt 2>&1
    timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${R}_bench_device_only.json 2>/dev/null || exit 1
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:terse|prolix|publish' --csv --log-file gpurun_out/${R}_launches.csv \
        python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${R}_ncu_launches.log 2>&1
else
    timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${R}_bench_device_only.json 2>/dev/null || exit 1
    timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:terse_encode|prolix_walk|prolix_unpack_seg' -s 6 -c 3 \
        -o gpurun_out/${R}_full -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${R}_ncu_full.log 2>&1
fi
ls -la gpurun_out | tail -12
