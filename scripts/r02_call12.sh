#!/usr/bin/env bash
# element kernels (xl_el_*), cfg 3 as one CUDA graph: parity on the device, timing, profile
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu_r02l.log 2>&1; tail -3 $OUT/pytest_gpu_r02l.log
timeout 120 python scripts/sharp_focus_table.py > $OUT/sharp_focus_r02l_eager.json 2> $OUT/sf.err; tail -c 420 $OUT/sharp_focus_r02l_eager.json; tail -3 $OUT/sf.err
timeout 120 python scripts/sharp_focus_table.py --graph > $OUT/sharp_focus_r02l_graph.json 2> $OUT/sfg.err; tail -c 420 $OUT/sharp_focus_r02l_graph.json; tail -5 $OUT/sfg.err
timeout 120 python scripts/sharp_focus_profile.py 1024 > $OUT/sharp_focus_profile_r02l.txt 2>&1; head -30 $OUT/sharp_focus_profile_r02l.txt
for m in grad; do
    timeout 200 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -x -k "element_kernels or sharp_focus" > $OUT/memcheck_elements_r02l.log 2>&1
    tail -3 $OUT/memcheck_elements_r02l.log
done
