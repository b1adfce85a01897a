#!/bin/bash
# tuning aid: run the dense workload through alternative builds of libkdot (kd_6d_pose_adlp_b200/lib/variants/*.so)
for f in kd_6d_pose_adlp_b200/lib/variants/*.so; do
  KDOT_LIB=$PWD/$f python bench.py --workload dense_b32 --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-60s %.3f ms  %.1f img/s  frac %.3f' % ('$f'.split('/')[-1], d['ms_per_step'], d['value'], d['roofline']['frac']))"
done
