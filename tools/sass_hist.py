"""Per-kernel SASS opcode histogram of libhjd.so (cuobjdump -sass): evidence of what the shipped binary is made of.

    python tools/sass_hist.py [path/to/libhjd.so] > profiles/r2_sass_histogram.txt

Columns: total instructions, then the opcodes that carry the design (packed FP32x2 FMAs, 256-bit loads,
fused truncate+saturate+pack conversions, dot-product de-quantisation, warp votes / reductions, shared-memory
traffic) and the tcgen05 opcodes of the tensor-core fused kernel (UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit,
LDTM = tcgen05.ld, UTCATOMSWS = TMEM allocation, SYNCS = mbarrier); HMMA / IMMA (mma.sync) and TMA tensor loads are absent.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["FFMA2", "FFMA", "FMUL2", "FADD2", "FADD", "FMUL", "LDG.E.256", "LDG.E.128", "LDG", "STG.E.128", "STG", "LDS", "STS",
         "F2IP", "F2I", "I2F", "I2FP", "IDP", "PRMT", "SHF", "LOP3", "VOTE", "REDUX", "SHFL", "BAR", "ATOMS", "LDC",
         "HFMA2", "HADD2", "HMUL2", "UTMALDG", "UTCHMMA", "UTCBAR", "LDTM", "SYNCS", "UTCATOMSWS", "HMMA", "IMMA", "LDL", "STL"]


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "hls_jpeg_decoder_b200", "libhjd.so")
    out = subprocess.check_output(["cuobjdump", "-sass", so], text=True)
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    kern = None
    hist = collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = subprocess.check_output(["c++filt", m.group(1)], text=True).strip().split("(")[0]
            hist[kern] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and kern:
            op = m.group(1)
            hist[kern]["_total"] += 1
            for w in WATCH:
                if w in ("LDG.E.256", "LDG.E.128", "STG.E.128"):      # width anywhere among the modifiers
                    base, width = w.split(".")[0], w.split(".")[-1]
                    if op.startswith(base + ".") and ("." + width) in op:
                        hist[kern][w] += 1
                elif op == w or op.startswith(w + "."):
                    hist[kern][w] += 1
    print(f"# {os.path.relpath(so, ROOT)}: cuobjdump -sass, architectures {arch}")
    print("# exact opcode or opcode-prefix matches; 'LDG' includes LDG.E.128/256, 'FFMA' excludes FFMA2")
    cols = ["_total"] + WATCH
    print(f"{'kernel':46s} " + " ".join(f"{c[:9]:>9s}" for c in cols))
    for k, h in hist.items():
        if h["_total"]:
            print(f"{k[:46]:46s} " + " ".join(f"{h[c]:9d}" for c in cols))


if __name__ == "__main__":
    main()
