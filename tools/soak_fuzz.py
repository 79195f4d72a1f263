"""Robustness soak (B200 only; not part of the test suite): randomly damaged JPEGs, good files in
between.  Every call must return, the good files must decode exactly as they do alone, and the
outcome (status and output bytes) must not depend on what the slabs held before.

    python tools/soak_fuzz.py [rounds] [seed] [hjd_batch_create flags, e.g. 128 = HJD_FLAG_TENSOR_CORE_IDCT]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def damage(jpg: bytes, rng) -> bytes:
    b = bytearray(jpg)
    for _ in range(int(rng.integers(1, 10))):
        if len(b) < 8:
            break
        pos = int(rng.integers(2, len(b)))
        mode = int(rng.integers(0, 7))
        if mode == 0:
            b[pos] = int(rng.integers(0, 256))
        elif mode == 1:
            b[pos] = 0xFF
        elif mode == 2 and pos + 1 < len(b):
            b[pos], b[pos + 1] = 0xFF, int(rng.integers(0xD0, 0xDA))
        elif mode == 3:
            del b[pos:pos + int(rng.integers(1, 400))]
        elif mode == 4:
            b[pos:pos] = bytes(rng.integers(0, 256, size=int(rng.integers(1, 64)), dtype=np.uint8))
        elif mode == 5:
            del b[pos:]                                     # truncation
        else:
            n = int(rng.integers(1, 200))
            b[pos:pos + n] = bytes(n)                       # a run of zero bytes
    return bytes(b)


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    import hls_jpeg_decoder_b200 as hjd
    from tools.soak_parity import make
    rng = np.random.default_rng(seed)
    good = make(0, seed, 640, 480)
    with hjd.BatchDecoder(0, flags) as d:
        d.upload([good])
        d.decode()
        assert d.status()[0] == 0
        want = d.rgb(0).copy()
        n_bad = n_rejected = n_warn = 0
        for r in range(rounds):
            base = [make(1 + r * 64 + k, seed, 700, 500) for k in range(64)]
            files = [good]
            for f in base:
                files.append(damage(f, rng))
            files.append(good)
            outs = []
            for rep in range(2):
                if rep:                                     # different leftovers in the slabs
                    d.upload(base[:8]); d.decode(); d.sync()
                d.upload(files)
                d.decode()
                st = d.status().copy()
                outs.append((st, d.rgb_slab().copy()))
                if st[0] != 0 or st[-1] != 0 or not np.array_equal(d.rgb(0), want) or not np.array_equal(d.rgb(len(files) - 1), want):
                    print(f"round {r}: a good neighbour was disturbed: {st[0]} {st[-1]}")
                    n_bad += 1
            if not np.array_equal(outs[0][0], outs[1][0]) or not np.array_equal(outs[0][1], outs[1][1]):
                print(f"round {r}: outcome depends on previous slab contents (status equal: {np.array_equal(outs[0][0], outs[1][0])})")
                n_bad += 1
            n_rejected += int((outs[0][0] < 0).sum())
            n_warn += int((outs[0][0] > 0).sum())
    print(f"fuzz soak: {rounds} rounds x 64 damaged files, seed {seed}, flags {flags}: {n_rejected} rejected by the parser, "
          f"{n_warn} decoded with warnings, {n_bad} failures")
    sys.exit(1 if n_bad else 0)


if __name__ == "__main__":
    main()
