#!/bin/bash
# A/B of the k-point lanes on the small BASELINE stand-ins: SGW_KLANES=1 (serial k loop) vs the default
for c in si c bn licl; do
  for l in 1 2 4 8; do
    echo "== $c lanes=$l"; SGW_KLANES=$l python tools/one_config.py $c 3 2>&1 | tail -2
  done
done
