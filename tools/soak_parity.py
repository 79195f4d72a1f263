"""Parity soak (B200 only; not part of the test suite): N random JPEGs of mixed geometry, sampling,
quality, content, Huffman optimisation and restart interval, decoded through the C ABI and compared
bit for bit (coefficients and RGB) with the oracle port.

    python tools/soak_parity.py [N] [seed] [max_width max_height]     (default sizes: up to 200 x 160)
"""
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def make(i, seed, max_w=200, max_h=160):
    from tools.gen_jpegs import encode_jpeg, synth_rgb
    rng = np.random.default_rng(seed * 100003 + i)
    w, h = int(rng.integers(8, max_w)), int(rng.integers(8, max_h))
    kind = int(rng.integers(0, 5))
    if kind == 0:
        rgb = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    elif kind == 1:
        rgb = synth_rgb(w, h, int(rng.integers(0, 1 << 30)), noise_sigma=float(rng.uniform(0, 20)))
    elif kind == 2:                                                     # hard edges, saturated colours
        rgb = np.zeros((h, w, 3), np.uint8)
        for _ in range(int(rng.integers(1, 12))):
            x0, y0 = int(rng.integers(0, w)), int(rng.integers(0, h))
            rgb[y0:y0 + int(rng.integers(1, h)), x0:x0 + int(rng.integers(1, w))] = rng.integers(0, 2, 3) * 255
    elif kind == 3:                                                     # smooth gradients
        yy, xx = np.mgrid[0:h, 0:w]
        rgb = np.stack([(xx * rng.uniform(0, 3) + yy * rng.uniform(0, 3) + c * 40) % 256 for c in range(3)], -1).astype(np.uint8)
    else:
        rgb = np.full((h, w, 3), rng.integers(0, 256, 3), np.uint8)
    sub = ["4:4:4", "4:2:2", "4:2:0"][int(rng.integers(0, 3))]
    gray = bool(rng.integers(0, 6) == 0)
    q = int(rng.choice([1, 5, 10, 25, 50, 75, 85, 90, 95, 98, 100]))
    ri = int(rng.choice([0, 0, 1, 2, 3, 8, 17]))
    opt = bool(rng.integers(0, 2))
    try:
        return encode_jpeg(rgb, q, sub, ri, gray=gray, optimize=opt)
    except OSError:                                                     # PIL refuses a few corner combinations
        return encode_jpeg(rgb, q, sub, 0, gray=gray, optimize=False)


def make_job(args):
    return make(*args)


def oracle(jpg):
    from oracle import port
    o = port.decode(jpg)
    return o["coef"], o["rgb"]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    max_w = int(sys.argv[3]) if len(sys.argv) > 4 else 200
    max_h = int(sys.argv[4]) if len(sys.argv) > 4 else 160
    import hls_jpeg_decoder_b200 as hjd
    with ProcessPoolExecutor(max_workers=os.cpu_count()) as ex:
        files = list(ex.map(make_job, [(i, seed, max_w, max_h) for i in range(n)], chunksize=4 if max_w > 400 else 32))
        want = list(ex.map(oracle, files, chunksize=8))
    bad = 0
    variants = (("default", 0), ("planes", hjd.FLAG_KEEP_PLANES),
                ("no-selfsync", hjd.FLAG_NO_SELFSYNC), ("tensor-core", hjd.FLAG_TENSOR_CORE_IDCT))
    for name, flags in variants:
        with hjd.BatchDecoder(0, flags) as d:
            d.upload(files)
            d.decode()
            st = d.status()
            coef = d.coefficients()
            for i in range(n):
                wc, wr = want[i]
                ok = st[i] == 0 and np.array_equal(d.image_coefficients(i, coef), wc) and np.array_equal(d.rgb(i), wr)
                if not ok:
                    bad += 1
                    diff = int(np.abs(d.rgb(i).astype(int) - wr.astype(int)).max()) if d.rgb(i).shape == wr.shape else -1
                    print(f"MISMATCH [{name}] image {i}: status {st[i]}, max |rgb diff| {diff}, {len(files[i])} bytes")
    pixels = sum(int(w[1].shape[0]) * int(w[1].shape[1]) for w in want)
    print(f"soak: {n} images, {pixels / 1e6:.1f} MP, seed {seed}, {len(variants)} decoder variants: {bad} mismatches")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
