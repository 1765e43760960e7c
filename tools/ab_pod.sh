#!/bin/bash
# PODFS leg of the bench, twice: prints snapshots/s and the eigen-solve time
for i in 1 2; do
  python bench.py --no-cfg1 --no-multi --no-cpu --no-e2e --pod-steps 2 --sustained-steps 0 --steps 1 --warmup 3 2>/dev/null | tail -1 > gpurun_out/ab.json
  python -c "
import json
d=json.load(open('gpurun_out/ab.json')); print('pod', d['pod']['value'], d['pod']['parts_ms']['eig'])"
done
