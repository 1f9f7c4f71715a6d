#!/bin/bash
# Developer tool (GPU box): the drop-in class on pageable memory (cxx/terse_bench) for several staging-thread counts.
python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch, bench
px = bench.synth_stack(torch, bench.CONFIGS["c2"], 0, 2000, torch.device("cuda", 0))
px.cpu().numpy().tofile("/dev/shm/trpx_probe.raw")
PY
for t in ${@:-3 6 8 12}; do
    echo "TRPX_STAGE_THREADS=$t: $(TRPX_STAGE_THREADS=$t cxx/terse_bench /dev/shm/trpx_probe.raw 262144 2000 3 | tail -1 | cut -c100-260)"
done
rm -f /dev/shm/trpx_probe.raw
