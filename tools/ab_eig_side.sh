#!/bin/bash
# A/B of PODSDF_EIG_SIDE (1 = bisection of the non-leading eigenvalues on a side stream, 0 = after the vectors)
for m in 1 0 1 0; do
  PODSDF_EIG_SIDE=$m python bench.py --no-cfg1 --no-multi --no-cpu --no-e2e --pod-steps 2 --sustained-steps 0 --steps 1 --warmup 3 2>/dev/null | tail -1 > gpurun_out/ab_$m.json
  python -c "
import json
d=json.load(open('gpurun_out/ab_$m.json')); print('side', $m, d['pod']['value'], d['pod']['parts_ms']['eig'])"
done
