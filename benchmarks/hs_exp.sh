#!/bin/bash
# on the GPU box: rebuild select_hist.cu with different warps-per-CTA, relink, time the C2 step
cd $GRAFT_REPO_ROOT
L=instacart_next_order_recommendation_b200/lib; C=instacart_next_order_recommendation_b200/csrc
cp $L/libicr_b200.so /tmp/lib_orig.so; cp $L/select_hist.cu.o /tmp/sel_orig.o
for W in 4 5 6 3 8; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DICR_HS_WARPS=$W -c $C/select_hist.cu -o $L/select_hist.cu.o 2>/dev/null
  nvcc -shared -cudart static -o $L/libicr_b200.so $L/*.o 2>/dev/null
  for dt in f32 bf16; do
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --dtype $dt 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('W=$W $dt', round(d['ms_per_step'],4), round(d['ms_per_step_min'],4), round(d['roofline']['kernel_ms_per_step'],4))"
  done
done
cp /tmp/lib_orig.so $L/libicr_b200.so; cp /tmp/sel_orig.o $L/select_hist.cu.o
