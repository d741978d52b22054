#!/usr/bin/env bash
# A/B timing of experiment variants of the library (experiment queue).
#   1. here (build container):   bash scripts/ab_variants.sh build MACRO...   (e.g. XL_NO_SYNCWARP: CTA barriers instead of the warp-level ones)
#        -> build/libxlprop_<macro>.so per macro (build/ is git-ignored but travels with gpurun)
#   2. on the GPU:   gpurun --timeout 600 -- 'bash scripts/ab_variants.sh time'
#        -> gpurun_out/ab_<name>.log: scripts/gpu_probe.py timings of every operation for the product library and each variant,
#           interleaved twice so that clock drift shows up as a difference between the two passes of the SAME library.
# Any source edit reshuffles the code generation of the whole translation unit (build note): only
# compare libraries built from the same tree.
set -u
cd "$(dirname "$0")/.."
case "${1:-}" in
build)
    shift
    for m in "$@"; do
        case "$m" in
        *)       python -m xlumina_b200.build --exp "$m" --out "build/libxlprop_$m.so" || exit 1 ;;
        esac
    done
    ls -la build/libxlprop_*.so ;;
time)
    mkdir -p gpurun_out
    # a variant is timed only after it passes the golden parity tests on the device (XLPROP_LIB selects the library)
    for lib in build/libxlprop_*.so; do
        [ -f "$lib" ] || continue
        name=$(basename "$lib" .so)
        case "$name" in *KEEP_SPECTRA*) export XL_KEEP_SPECTRA=1 ;; *) unset XL_KEEP_SPECTRA ;; esac
        XLPROP_LIB="$PWD/$lib" timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "golden or four_f or 2048" > "gpurun_out/ab_${name}_parity.log" 2>&1
        echo "== $name parity: $(tail -1 "gpurun_out/ab_${name}_parity.log")"
    done
    for pass in 1 2; do
        for lib in xlumina_b200/libxlprop.so build/libxlprop_*.so; do
            [ -f "$lib" ] || continue
            name=$(basename "$lib" .so)
            case "$name" in *KEEP_SPECTRA*) export XL_KEEP_SPECTRA=1 ;; *) unset XL_KEEP_SPECTRA ;; esac
            XLPROP_LIB="$PWD/$lib" timeout 120 python scripts/gpu_probe.py --nosmoke > "gpurun_out/ab_${name}_$pass.log" 2>&1
            echo "== $name (pass $pass)"; grep -E "us$|us " "gpurun_out/ab_${name}_$pass.log" | head -30
        done
    done
    python scripts/ab_report.py gpurun_out | tee gpurun_out/ab_report.txt ;;
*)  echo "usage: $0 build MACRO... | time"; exit 2 ;;
esac
