#!/bin/bash
# stage-2 time against the number of sweeps in flight (CTAs of the persistent chase kernel)
for v in 0 7; do for g in 148 164 200 240 296; do
  echo -n "chase variant $v, G $g: "
  PODSDF_CHASE=$v PODSDF_CHASE_G=$g timeout 200 python tools/time_sbr.py 10000 2 2>&1 | tail -1 | cut -c1-80
done; done
