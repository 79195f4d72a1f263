"""Summarise an ncu --set full report into profiles/ (text summary + per-kernel DRAM traffic JSON).

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_summary.py /tmp/raw.csv profiles/r1_final "<command that was profiled>"
"""
import csv
import json
import re
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block', 'launch__grid_size', 'launch__block_size',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct']


def to_bytes(x, unit):
    v = float(x.replace(',', ''))
    return v * (1e9 if unit.startswith('G') else 1e6 if unit.startswith('M') else 1e3 if unit.startswith('K') else 1)


def main():
    raw, prefix, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    out = [f"# {cmd}", "# ncu --set full --clock-control none, B200.  One launch of each kernel = the whole batch;",
           "# traffic = dram__bytes_read.sum + dram__bytes_write.sum per launch.  Durations under ncu are cold-cache and",
           "# serialised: compare SHARES with bench.py's CUDA-event stage_ms, not absolutes."]
    traffic = {}
    for r in rows[2:]:
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                out.append(f"{w:78s} {r[i][:70]:>70s} {units[i]}")
        stalls = []
        for i, h in enumerate(hdr):
            if 'smsp__average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio'):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v > 0.3:
                    stalls.append((round(v, 2), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
        out.append(f"warp stall reasons per issue: {sorted(stalls, reverse=True)}")
        name = r[hdr.index('Kernel Name')]
        key = ('scan' if 'marker' in name else 'entropy' if 'entropy' in name else
               'idct' if ('idct' in name or 'mcu_rgb' in name) else 'color' if 'color' in name else
               re.sub(r'\W+', '_', name.split('(')[0].replace('hjd_k_', ''))[:24])
        ir, iw = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
        traffic.setdefault(key, int(to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw])))
        out.append('---')
    open(prefix + "_ncu_summary.txt", "w").write("\n".join(out) + "\n")
    json.dump({"_source": f"{prefix}_ncu_summary.txt: ncu --set full --clock-control none on `{cmd}`, "
                          "dram__bytes_read.sum + dram__bytes_write.sum per launch",
               "images": 1024, "dram_bytes_per_launch": traffic}, open(prefix + "_traffic.json", "w"), indent=1)
    print("\n".join(out))


if __name__ == "__main__":
    main()
