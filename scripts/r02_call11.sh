#!/usr/bin/env bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu_r02k.log 2>&1; tail -3 $OUT/pytest_gpu_r02k.log
timeout 120 python scripts/kern_probe.py 10 > $OUT/kern_r02k.log 2>&1; grep -E "^---|rs_cols_gz" $OUT/kern_r02k.log
timeout 120 python scripts/sharp_focus_profile.py 1024 > $OUT/sharp_focus_profile_r02k.txt 2>&1; head -45 $OUT/sharp_focus_profile_r02k.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:xl_kernel -s 7 -c 7 -o /tmp/prof_vczt_r02k \
    python scripts/prof_rs.py 2048 vcztgrad 2 > $OUT/ncu_vczt_r02k.log 2>&1
ncu -i /tmp/prof_vczt_r02k.ncu-rep --page source --csv > /tmp/ncu_r02k_vczt_source.csv 2>/dev/null
ncu -i /tmp/prof_vczt_r02k.ncu-rep --page raw --csv > $OUT/ncu_r02k_vcztgrad_raw.csv 2>/dev/null
for k in "XlCztAxis<4096, 2, 0, 1>" "XlCztAxis<4096, 0, 1, 1>" "XlCztAxis<4096, 1, 0, 2>" "XlCztAxis<4096, 0, 1, 2>"; do
  python scripts/ncu_hot.py /tmp/ncu_r02k_vczt_source.csv "$k" 2>&1 | tee -a $OUT/ncu_hot_r02k.txt
done
