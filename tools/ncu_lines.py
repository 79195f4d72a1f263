"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line.

    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:NAME > src.csv
    python tools/ncu_lines.py src.csv [top_n]
"""
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    hdr = rows[h]
    ie = hdr.index("Instructions Executed")
    sm = hdr.index("# Samples")
    av = hdr.index("Avg. Threads Executed")
    exc = hdr.index("L1 Wavefronts Shared Excessive") if "L1 Wavefronts Shared Excessive" in hdr else None
    out, tot, tots = [], 0, 0
    for r in rows[h + 1:]:
        if not r or not r[0].strip().isdigit():
            continue
        try:
            n = int(r[ie]); s = int(r[sm])
        except ValueError:
            continue
        if n == 0 and s == 0:
            continue
        tot += n; tots += s
        out.append((n, s, r[av], r[exc] if exc is not None else "-", int(r[0]), r[1].strip()[:100]))
    print(f"total warp instructions {tot}, samples {tots}")
    for n, s, a, e, ln, text in sorted(out, reverse=True)[:top]:
        print(f"{n:>12d} {100 * n / max(tot, 1):5.1f}%  samp {100 * s / max(tots, 1):5.1f}%  thr {a:>5s}  smem_exc {e:>9s}  L{ln}: {text}")


if __name__ == "__main__":
    main()
