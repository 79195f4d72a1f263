#!/usr/bin/env bash
# tools/refresh_profiles.sh -- run on the GPU box (under gpurun): regenerates everything that
# profiles/<prefix>_* is made from.
#   bash tools/refresh_profiles.sh bench|list|full [prefix]        (prefix default: r1_final)
# bench: both bench arms on the headline config + the other workloads; list: ncu launch list of one
# bench command; full: ncu --set full of its three kernels.  One profiler pass per call, and only after
# the same command has exited 0 without the profiler.  Outputs go to gpurun_out/ (scratch);
# tools/ncu_summary.py and the copy into profiles/ happen afterwards on the development machine.
set -uo pipefail
MODE="${1:-bench}"
P="${2:-r1_final}"
O=gpurun_out
mkdir -p "$O"
run() { echo "== $*" >&2; timeout 900 "$@"; }
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
case "$MODE" in
bench)
    run python bench.py --impl reference --steps 2 --warmup 1 > "$O/${P}_bench_reference.json" 2> "$O/${P}_bench_reference.log"
    run python bench.py > "$O/${P}_bench.json" 2> "$O/${P}_bench.log" || exit 1
    for cfg in c2q50 c2q95 c2nr c4 c5; do
        run python bench.py --config "$cfg" --no-cpu-baseline --e2e-steps 2 > "$O/${P}_bench_${cfg}.json" 2> "$O/${P}_bench_${cfg}.log"
    done ;;
list)
    run $CMD > /dev/null 2>&1 || exit 1
    run ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$O/${P}_launches.csv" $CMD > "$O/${P}_ncu_launches.log" 2>&1 ;;
full)
    run $CMD > /dev/null 2>&1 || exit 1
    run ncu --set full --clock-control none --import-source on -k regex:hjd_k_ -c 3 -f -o "$O/${P}_full" $CMD > "$O/${P}_ncu_full.log" 2>&1 ;;
*)  echo "unknown mode $MODE" >&2; exit 2 ;;
esac
ls -la "$O" >&2
