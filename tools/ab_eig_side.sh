#!/bin/bash
# A/B of the side-stream bisection: fat CTAs that leave nm SMs alone (default) / thin CTAs on every SM / serialised
run() { python bench.py --no-cfg1 --no-multi --no-cpu --no-e2e --pod-steps 2 --sustained-steps 0 --steps 1 --warmup 3 2>/dev/null | tail -1 > gpurun_out/ab.json
  python -c "
import json
d=json.load(open('gpurun_out/ab.json')); print('$1', d['pod']['value'], d['pod']['parts_ms']['eig'])"; }
for i in 1 2; do
  run fat
  PODSDF_EIG_SIDE_FAT=0 run thin
done
PODSDF_EIG_SIDE=0 run serial
