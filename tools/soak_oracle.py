"""Oracle pin soak (CPU, needs /root/reference built into oracle/_ref by oracle/build_ref.sh): the plain-C
restatement (oracle/jpeg_oracle.c) against the real reference code on random JPEGs, bit for bit
(coefficients, planes, RGB).

    python tools/soak_oracle.py [N] [seed]
"""
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def check(args):
    i, seed = args
    from oracle import port, refbind
    from tools.soak_parity import make
    jpg = make(i, seed)
    a = port.decode(jpg)
    b = refbind.decode(jpg, mode=1)
    ok = np.array_equal(a["coef"], b["coef"]) and np.array_equal(a["rgb"], b["rgb"])
    if "planes" in a and "planes" in b:
        ok = ok and all(np.array_equal(x, y) for x, y in zip(a["planes"], b["planes"]))
    return i, ok, a["rgb"].shape[0] * a["rgb"].shape[1]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 11
    bad, px = 0, 0
    with ProcessPoolExecutor(max_workers=os.cpu_count()) as ex:
        for i, ok, p in ex.map(check, [(i, seed) for i in range(n)], chunksize=16):
            px += p
            if not ok:
                bad += 1
                print("MISMATCH port vs reference, image", i)
    print(f"oracle soak: {n} images, {px / 1e6:.1f} MP, seed {seed}: {bad} mismatches")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
