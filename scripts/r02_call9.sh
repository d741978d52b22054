#!/usr/bin/env bash
# dot_z in registers, dual H/dH generation, stagger knob for the persistent kernels; source-level ncu of RS forward+gradient
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu_r02i.log 2>&1; tail -3 $OUT/pytest_gpu_r02i.log
timeout 120 python scripts/kern_probe.py 10 > $OUT/kern_r02i.log 2>&1; cat $OUT/kern_r02i.log
for ns in 1500 3000 5000; do
  XL_STAGGER_NS=$ns timeout 120 python scripts/kern_probe.py 10 > $OUT/kern_r02i_st$ns.log 2>&1
  echo "== stagger $ns"; grep -E "^---|rs_cols" $OUT/kern_r02i_st$ns.log | head -8
done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:xl_kernel -s 8 -c 8 -o /tmp/prof_grad_r02i \
    python scripts/prof_rs.py 2048 grad 2 > $OUT/ncu_grad_r02i.log 2>&1
ncu -i /tmp/prof_grad_r02i.ncu-rep --page raw --csv > $OUT/ncu_r02i_grad_raw.csv 2>/dev/null
ncu -i /tmp/prof_grad_r02i.ncu-rep --page source --csv > /tmp/ncu_r02i_grad_source.csv 2>/dev/null
for k in XlHRows XlHCols XlRsRowsFwd XlRsColsAsync XlRsRowsInv XlRsRowsDual XlRsColsGzAsync; do
  python scripts/ncu_hot.py /tmp/ncu_r02i_grad_source.csv $k 2>&1 | tee -a $OUT/ncu_hot_r02i.txt
done
gzip -c /tmp/ncu_r02i_grad_source.csv > $OUT/ncu_r02i_grad_source.csv.gz; ls -la $OUT/ncu_r02i_grad_source.csv.gz
