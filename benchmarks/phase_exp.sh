#!/bin/bash
# C2 step time against the phase plan of the tensor path: tiles of the dense first phase x growth factor of the sparse phases
cd ${GRAFT_REPO_ROOT:-.}
for d0 in 2 4 6 8; do for gr in 5 6 8 10 12 16; do
  ICR_K2_DENSE0=$d0 ICR_K2_GROWTH=$gr python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dense0=$d0 growth=$gr', round(d['ms_per_step'],4), round(d['ms_per_step_min'],4), 'gemm', round(d['roofline']['kernel_ms_per_step'],4), 'launches', d['gpu_launches']//20)"
done; done
