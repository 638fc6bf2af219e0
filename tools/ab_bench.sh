# A/B on one box: baseline lib vs current lib, alternating, 2 runs each
for i in 1 2; do for lib in base new; do
  if [ $lib = base ]; then export EESEG_LIB=ee_semantic_segmentation_b200/libeeseg_b200_base.so; else unset EESEG_LIB; fi
  timeout 200 python bench.py --legs none --no-cpu-baseline > gpurun_out/ab_$lib$i.json 2>gpurun_out/ab_$lib$i.err
  python -c "
import json; d=json.loads(open('gpurun_out/ab_$lib$i.json').read().strip().splitlines()[-1]); print('bench $lib $i', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d.get('operating_point',{}).get('value'))"
done; done
