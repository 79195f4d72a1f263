#!/usr/bin/env bash
# tools/refresh_profiles.sh -- run on the GPU box (under gpurun): regenerates everything that
# profiles/<prefix>_* is made from.
#   bash tools/refresh_profiles.sh bench|list|full|all [prefix]        (prefix default: r2_final)
# bench: both bench arms on the headline config + the other workloads; list: ncu launch lists of the bench
# command (config 2), of 256 restart-free twins and of config 4; full: ncu --set full of config 2's three
# kernels and of kernel 1b's.  One profiler pass per command, and only after the same command has exited 0
# without the profiler.  Outputs go to gpurun_out/ (scratch); tools/ncu_summary.py and the copy into
# profiles/ happen afterwards on the development machine.
set -uo pipefail
MODE="${1:-bench}"
P="${2:-r2_final}"
O=gpurun_out
mkdir -p "$O"
run() { echo "== $*" >&2; timeout 900 "$@"; }
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
NR="$CMD --config c2nr --images 256"
C4="$CMD --config c4"
do_bench() {
    run python bench.py --impl reference --steps 2 --warmup 1 > "$O/${P}_bench_reference.json" 2> "$O/${P}_bench_reference.log"
    run python bench.py > "$O/${P}_bench.json" 2> "$O/${P}_bench.log" || exit 1
    : > "$O/${P}_bench_other_workloads.jsonl"
    for cfg in c2q50 c2q95 c2nr c4 c5; do
        run python bench.py --config "$cfg" --no-cpu-baseline --no-extras --e2e-steps 2 >> "$O/${P}_bench_other_workloads.jsonl" 2>> "$O/${P}_bench_other.log"
    done
}
do_list() {
    run $CMD > /dev/null 2>&1 || exit 1
    run ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$O/${P}_launches.csv" $CMD > "$O/${P}_ncu_launches.log" 2>&1
    run $NR > /dev/null 2>&1 && run ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$O/${P}_launches_restart_free.csv" $NR > "$O/${P}_ncu_launches_nr.log" 2>&1
    run $C4 > /dev/null 2>&1 && run ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$O/${P}_launches_c4.csv" $C4 > "$O/${P}_ncu_launches_c4.log" 2>&1
}
do_full() {
    run $CMD > /dev/null 2>&1 || exit 1
    run ncu --set full --clock-control none --import-source on -k regex:hjd_k_ -c 3 -f -o "$O/${P}_full" $CMD > "$O/${P}_ncu_full.log" 2>&1
    run ncu --set full --clock-control none --import-source on -k regex:'ss_|destuff' -c 6 -f -o "$O/${P}_ss_full" $NR > "$O/${P}_ncu_ss_full.log" 2>&1
}
case "$MODE" in
bench) do_bench ;;
list)  do_list ;;
full)  do_full ;;
all)   do_bench; do_list; do_full ;;
*)  echo "unknown mode $MODE" >&2; exit 2 ;;
esac
ls -la "$O" >&2
