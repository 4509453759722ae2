#!/bin/bash
# C2 fp32 step time for variants of the re-scoring kernel's launch configuration (compile-time): CTAs/SM x rows in flight
for cfg in "3 1" "4 0" "6 0" "4 1"; do
  set -- $cfg
  ICR_NVCC_DEFS="-DICR_RS_MINB=$1 -DICR_RS_WIDE=$2" python -m instacart_next_order_recommendation_b200.build --force > /dev/null 2>&1
  ms=$(python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-side --no-sharded 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4f %.4f' % (d['ms_per_step'], d['ms_per_step_min']))")
  echo "RS_MINB=$1 RS_WIDE=$2 step_ms min_ms: $ms"
done
