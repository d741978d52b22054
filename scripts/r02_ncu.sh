#!/usr/bin/env bash
# ncu --set full of one forward+gradient iteration per operator family (second iteration), raw + source pages as CSV.
set -u
OUT=gpurun_out
TAG=${1:-r02}
mkdir -p $OUT
for m in grad cztgrad; do
    timeout 300 ncu --set full --import-source on --clock-control none -k regex:xl_kernel -s ${2:-9} -c ${3:-9} -o /tmp/prof_${m}_$TAG \
        python scripts/prof_rs.py 2048 $m 2 > $OUT/ncu_${m}_$TAG.log 2>&1
    ncu -i /tmp/prof_${m}_$TAG.ncu-rep --page raw --csv > $OUT/ncu_${TAG}_${m}_raw.csv 2>/dev/null
    ncu -i /tmp/prof_${m}_$TAG.ncu-rep --page source --csv > $OUT/ncu_${TAG}_${m}_source.csv 2>/dev/null
    ls -la /tmp/prof_${m}_$TAG.ncu-rep
done
du -sh $OUT
