#!/bin/bash
# leading eigenvalues: one CTA of 256 lanes per eigenvalue (default) against one warp per eigenvalue
run() { python bench.py --no-cfg1 --no-multi --no-cpu --no-e2e --pod-steps 2 --sustained-steps 0 --steps 1 --warmup 3 2>/dev/null | tail -1 > gpurun_out/ab.json
  python -c "
import json
d=json.load(open('gpurun_out/ab.json')); print('$1', d['pod']['value'], d['pod']['parts_ms']['eig'])"; }
for i in 1 2; do
  run block
  PODSDF_EIG_MULTISECT_WARP=1 run warp
done
