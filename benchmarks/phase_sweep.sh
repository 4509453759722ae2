#!/bin/bash
# C2 step time over the phase plan of the tensor path: rows scored densely first (tiles of 256) x growth of the sparse phases.
# usage: bash benchmarks/phase_sweep.sh [f32|bf16] > gpurun_out/phase_sweep.txt
dt=${1:-f32}
for d0 in 2 4 8; do
  for g in 4 6 8 12 16 24; do
    ms=$(ICR_K2_DENSE0=$d0 ICR_K2_GROWTH=$g python bench.py --dtype $dt --steps 10 --warmup 3 --no-cpu-baseline --no-side --no-sharded 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4f %.4f %d' % (d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['gpu_launches']/d['steps']))")
    echo "dense0=$d0 growth=$g step_ms kernel_ms launches: $ms"
  done
done
