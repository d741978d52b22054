#!/usr/bin/env bash
# One gpurun call's worth of evidence for a round (run it UNDER gpurun, 1 GPU, from the repo root):
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash scripts/measure_round.sh r02'
# then, back in the build container:   python scripts/make_profiles.py r02
# Every step has its own timeout; only CSV/JSON/log summaries are written to gpurun_out/ (the .ncu-rep files stay in /tmp:
# gpurun_out/ is limited to 64 MiB).  ~8 GPU-minutes.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
timeout 120 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; tail -2 $OUT/pytest_gpu_$TAG.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; tail -1 $OUT/smoke_$TAG.log
timeout 400 python bench.py > $OUT/bench_${TAG}_final.json 2> $OUT/bench_${TAG}_final.err; tail -c 400 $OUT/bench_${TAG}_final.json
timeout 90 python scripts/gpu_probe.py --nosmoke > $OUT/probe_$TAG.log 2>&1
# BASELINE configs 3 and 4 as whole tables (value + gradient of the loss), one GPU
timeout 90 python scripts/sharp_focus_table.py --graph > $OUT/sharp_focus_$TAG.json 2> $OUT/sharp_focus_$TAG.err; tail -c 300 $OUT/sharp_focus_$TAG.json
timeout 90 python scripts/four_f_sharded.py > $OUT/four_f_$TAG.json 2> $OUT/four_f_$TAG.err; tail -c 300 $OUT/four_f_$TAG.json
# launch list of the bench command itself (cold cache + serialised: shares only)
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -k xl_kernel -c 500 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
# full metrics of one forward+gradient per operator (second iteration of scripts/prof_rs.py): launches per iteration
# RS 8, VRS 9, CZT 6, VCZT 7, high-NA 5 (7 in the first call, which fills the cached tables)
for spec in rsgrad:grad:8:8 vrsgrad:vrsgrad:9:9 cztgrad:cztgrad:6:6 vcztgrad:vcztgrad:7:7 highna:highna:7:5; do
    IFS=: read name mode skip n <<< "$spec"
    timeout 300 ncu --set full --clock-control none -k regex:xl_kernel -s $skip -c $n -o /tmp/prof_${name}_$TAG \
        python scripts/prof_rs.py 2048 $mode 2 > $OUT/ncu_${name}.log 2>&1
    ncu -i /tmp/prof_${name}_$TAG.ncu-rep --page raw --csv > $OUT/ncu_${TAG}_${name}_raw.csv 2>/dev/null
done
# shared-memory race and out-of-bounds checks of every kernel family at a small size (the host emulation runs the phases of
# a CTA one after the other, so only the device can show a missing barrier)
for m in grad vrsgrad cztgrad vcztgrad highna; do
    timeout 200 compute-sanitizer --tool racecheck --print-limit 5 python scripts/prof_rs.py 512 $m 1 > $OUT/racecheck_${m}_$TAG.log 2>&1
    tail -2 $OUT/racecheck_${m}_$TAG.log
done
timeout 120 compute-sanitizer --tool memcheck --print-limit 5 python scripts/prof_rs.py 128 grad 1 > $OUT/memcheck_grad_$TAG.log 2>&1; tail -2 $OUT/memcheck_grad_$TAG.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/smi_$TAG.txt
du -sh $OUT
