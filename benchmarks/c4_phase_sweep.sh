#!/bin/bash
# Phase plan of the queries-on-M kernel (K2) on one C4 shard at Q = 128 / 256: growth factor x dense first phase.
# Usage (GPU box): bash benchmarks/c4_phase_sweep.sh > gpurun_out/c4_phase_sweep.txt
export ICR_C4_QS=${ICR_C4_QS:-128,256}
for G in 8 32 64 128 512; do
  for D0 in 4 16; do
    echo "== ICR_K2_GROWTH=$G ICR_K2_DENSE0=$D0"
    ICR_K2_GROWTH=$G ICR_K2_DENSE0=$D0 timeout 120 python benchmarks/c4_shard.py 2>&1 | tail -4
  done
done
