#!/bin/bash
# A/B of an environment switch on one box: tools/ab_env.sh VAR v1 v2 ... (two alternations, headline bench only)
VAR=$1; shift
for i in 1 2; do for v in "$@"; do
  env $VAR=$v timeout 200 python bench.py --legs none --no-cpu-baseline > gpurun_out/abenv_${v}_$i.json 2>gpurun_out/abenv_${v}_$i.err
  python -c "
import json; d=json.loads(open('gpurun_out/abenv_${v}_$i.json').read().strip().splitlines()[-1]); print('bench $VAR=$v run $i', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"
done; done
