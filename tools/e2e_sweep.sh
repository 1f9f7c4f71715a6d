#!/bin/bash
# Developer tool (GPU box): sweep the host-pipeline knobs of the e2e measurement.
# usage: tools/e2e_sweep.sh "chunk batch_mb enc_lanes dec_lanes enc_threads dec_threads ramp enc_batch_mb" ...
for cfg in "$@"; do
  set -- $cfg
  f=gpurun_out/e2e_$1_$2_$3_$4_$5_$6_$7_$8
  TRPX_BATCH_MB=$2 TRPX_ENC_BATCH_MB=${8:-$2} TRPX_ENC_LANES=$3 TRPX_DEC_LANES=$4 timeout 200 python bench.py --steps 1 --no-cpu-baseline --e2e-chunk $1 --e2e-enc-threads ${5:-1} --e2e-dec-threads ${6:-1} --e2e-ramp ${7:-250} > $f.json 2> $f.err
  python - "$f.json" "$cfg" <<'PY' || tail -5 $f.err
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])["e2e"]
print(sys.argv[2], d["mode"], "streamed %.1f ms" % d["streamed"]["ms_per_step"], "sequential %.1f ms (enc %.1f dec %.1f)" % (d["sequential"]["ms_per_step"], d["sequential"]["encode_ms"], d["sequential"]["decode_ms"]))
PY
done
