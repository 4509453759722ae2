#!/bin/bash
# C2 step time over the bootstrap size of the single-launch K2 path (survivors per query the threshold is sized for)
dt=${1:-f32}
for t in 250 350 450 550 700 900 1200 1800 2600; do
  ms=$(ICR_K2_BOOT_TARGET=$t python bench.py --dtype $dt --steps 10 --warmup 3 --no-cpu-baseline --no-side --no-sharded 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4f %.4f %d' % (d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['gpu_launches']/d['steps']))")
  echo "target=$t step_ms kernel_ms launches: $ms"
done
