#!/usr/bin/env bash
# Round 2, first GPU call: (1) the GPU test suite with the new 2048^2 device-vs-oracle tests, (2) bench + operator probe +
# whole-table timings of the round-1 library (baseline of this round), (3) racecheck / memcheck of every kernel family,
# (4) A/B of the experiment variants built by `scripts/ab_variants.sh build all` (parity gate on the device, then timings).
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scripts/r02_call1.sh'
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/smi_r02a.txt
timeout 900 python -m pytest tests -m gpu -x -q -s > $OUT/pytest_gpu_r02a.log 2>&1; tail -3 $OUT/pytest_gpu_r02a.log
grep "2048^2" $OUT/pytest_gpu_r02a.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_r02a.log 2>&1; tail -1 $OUT/smoke_r02a.log
timeout 200 python bench.py > $OUT/bench_r02a.json 2> $OUT/bench_r02a.err; tail -c 600 $OUT/bench_r02a.json
timeout 90 python scripts/sharp_focus_table.py > $OUT/sharp_focus_r02a.json 2> $OUT/sharp_focus_r02a.err; tail -c 300 $OUT/sharp_focus_r02a.json
timeout 90 python scripts/four_f_sharded.py > $OUT/four_f_r02a.json 2> $OUT/four_f_r02a.err; tail -c 300 $OUT/four_f_r02a.json
for m in grad vrsgrad cztgrad vczt; do
    timeout 150 compute-sanitizer --tool racecheck --print-limit 5 python scripts/prof_rs.py 128 $m 1 > $OUT/racecheck_${m}_r02a.log 2>&1
    echo "racecheck $m: $(tail -1 $OUT/racecheck_${m}_r02a.log)"
done
timeout 150 compute-sanitizer --tool memcheck --print-limit 5 python scripts/prof_rs.py 128 grad 1 > $OUT/memcheck_grad_r02a.log 2>&1
echo "memcheck grad: $(tail -1 $OUT/memcheck_grad_r02a.log)"
# ---- A/B of the variants
for lib in build/libxlprop_*.so; do
    [ -f "$lib" ] || continue
    name=$(basename "$lib" .so)
    case "$name" in *KEEP_SPECTRA*) export XL_KEEP_SPECTRA=1 ;; *) unset XL_KEEP_SPECTRA ;; esac
    XLPROP_LIB="$PWD/$lib" timeout 200 python -m pytest tests/test_gpu_parity.py -q -x -k "golden or four_f or 2048" > "$OUT/ab_${name}_parity.log" 2>&1
    echo "== $name parity: $(tail -1 "$OUT/ab_${name}_parity.log")"
done
unset XL_KEEP_SPECTRA
for pass in 1 2; do
    for lib in xlumina_b200/libxlprop.so build/libxlprop_*.so; do
        [ -f "$lib" ] || continue
        name=$(basename "$lib" .so)
        if [ "$pass" = 2 ] && [ "$name" != libxlprop ]; then
            case "$name" in *PERSIST*|*STAGE*|*KEEP*) ;; *) continue ;; esac      # second pass: product + the structural variants
        fi
        case "$name" in *KEEP_SPECTRA*) export XL_KEEP_SPECTRA=1 ;; *) unset XL_KEEP_SPECTRA ;; esac
        XLPROP_LIB="$PWD/$lib" timeout 120 python scripts/gpu_probe.py --nosmoke --only2048 > "$OUT/ab_${name}_$pass.log" 2>&1
    done
done
python scripts/ab_report.py $OUT | tee $OUT/ab_report_r02a.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv >> $OUT/smi_r02a.txt
du -sh $OUT
