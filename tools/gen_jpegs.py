"""Seeded synthetic baseline-JPEG generator (PIL / libjpeg-turbo) for tests and bench.py.

There is no network and the reference ships a single image, so every workload of
BASELINE.json is synthesised here: photographic-like content (smooth
low-frequency fields + mid-frequency texture + hard-edged shapes + N(0, sigma) sensor
noise), encoded as baseline sequential JPEG with standard (non-optimised) Huffman tables.

    config 2/3 : 1920x1080, 4:2:0, q85, restart interval 8 MCUs  (seed 1234+i)
    config 4   : 8192x8192, 4:4:4, q85, no restart markers       (seed 4)
    config 5   : 256x256, even i grayscale / odd i 4:2:0, q75, Ri=8 (seed 5000+i)

CLI:  python tools/gen_jpegs.py --config 2 --count 4 --out /tmp/jpegs
"""
from __future__ import annotations

import argparse
import functools
import io
import os
from concurrent.futures import ProcessPoolExecutor

import numpy as np
from PIL import Image

SUBSAMPLING = {"4:4:4": 0, "4:2:2": 1, "4:2:0": 2, "4:4:0": "4:4:0"}


def synth_rgb(width: int, height: int, seed: int, noise_sigma: float = 6.0) -> np.ndarray:
    """Deterministic photographic-like RGB content, uint8 [h, w, 3]."""
    rng = np.random.default_rng(seed)

    def field(cell: int, amp: float) -> np.ndarray:
        gh, gw = height // cell + 3, width // cell + 3
        g = rng.uniform(-amp, amp, size=(gh, gw, 3)).astype(np.float32)
        chans = []
        for c in range(3):
            im = Image.fromarray(g[:, :, c], mode="F").resize((gw * cell, gh * cell), Image.BICUBIC)
            chans.append(np.asarray(im, dtype=np.float32)[cell:cell + height, cell:cell + width])
        return np.stack(chans, axis=-1)

    img = 128.0 + field(max(64, min(width, height) // 4), 70.0) + field(24, 26.0) + field(6, 9.0)
    # hard-edged shapes: rectangles with flat colours and a few diagonal stripes
    for _ in range(6 + (width * height) // 200000):
        x0, y0 = int(rng.integers(0, width)), int(rng.integers(0, height))
        w, h = int(rng.integers(4, max(5, width // 5))), int(rng.integers(4, max(5, height // 5)))
        col = rng.uniform(10, 245, size=3).astype(np.float32)
        alpha = float(rng.uniform(0.35, 1.0))
        sl = (slice(y0, min(height, y0 + h)), slice(x0, min(width, x0 + w)))
        img[sl] = (1 - alpha) * img[sl] + alpha * col
    if noise_sigma > 0:
        img += rng.standard_normal(size=img.shape, dtype=np.float32) * noise_sigma
    return np.clip(img, 0, 255).astype(np.uint8)


def encode_jpeg(rgb: np.ndarray, quality: int = 85, subsampling: str = "4:2:0",
                restart_blocks: int = 0, gray: bool = False, optimize: bool = False) -> bytes:
    """Baseline sequential JPEG; restart_blocks = restart interval in MCUs (0 = none)."""
    im = Image.fromarray(rgb)
    if gray:
        im = im.convert("L")
    buf = io.BytesIO()
    kw = dict(format="JPEG", quality=quality, optimize=optimize, progressive=False)
    if not gray:
        kw["subsampling"] = SUBSAMPLING[subsampling]
    if restart_blocks:
        kw["restart_marker_blocks"] = restart_blocks
    im.save(buf, **kw)
    return buf.getvalue()


@functools.lru_cache(maxsize=2)
def _base_scene(width: int, height: int, group: int) -> np.ndarray:
    # noise-free scene with a margin, shared by 64 consecutive images (float32 for later noise)
    return synth_rgb(width + 128, height + 128, 900000 + group, noise_sigma=0.0).astype(np.float32)


def synth_rgb_fast(width: int, height: int, i: int, noise_sigma: float = 6.0) -> np.ndarray:
    """Bench-scale content: image i = a per-image crop/flip of its group's scene + its own sensor
    noise (seed 1234+i).  ~5x cheaper than synth_rgb, still unique bytes per image."""
    rng = np.random.default_rng(1234 + i)
    base = _base_scene(width, height, i // 64)
    dx, dy = int(rng.integers(0, 129)), int(rng.integers(0, 129))
    img = base[dy:dy + height, dx:dx + width]
    if rng.integers(0, 2):
        img = img[:, ::-1]
    img = img + rng.standard_normal(size=img.shape, dtype=np.float32) * noise_sigma
    return np.clip(img, 0, 255).astype(np.uint8)


def make_c2(i: int, restart: bool = True, width: int = 1920, height: int = 1080, quality: int = 85) -> bytes:
    """Config 2/3 image i (and, with restart=False, its restart-free twin: same pixels)."""
    return encode_jpeg(synth_rgb_fast(width, height, i), quality, "4:2:0", 8 if restart else 0)


def make_c4(size: int = 8192, seed: int = 4) -> bytes:
    return encode_jpeg(synth_rgb(size, size, seed), 85, "4:4:4", 0)


def make_c5(i: int, restart: bool = True) -> bytes:
    rgb = synth_rgb(256, 256, 5000 + i)
    return encode_jpeg(rgb, 75, "4:2:0", 8 if restart else 0, gray=(i % 2 == 0))


def make_c2_restart_free(i: int) -> bytes:
    """The restart-free twin of config-2 image i (same pixels, no DRI): what most real-world files look like."""
    return make_c2(i, restart=False)


def make_c2_q50(i: int) -> bytes:
    """Quality sensitivity of config 2: same pixels at q50 (about 0.4x the scan bytes)."""
    return make_c2(i, quality=50)


def make_c2_q95(i: int) -> bytes:
    """Quality sensitivity of config 2: same pixels at q95 (about 2.4x the scan bytes)."""
    return make_c2(i, quality=95)


def _job(args):
    kind, i = args
    return {"c2": make_c2, "c2nr": make_c2_restart_free, "c2q50": make_c2_q50, "c2q95": make_c2_q95,
            "c5": make_c5}[kind](i)


def make_batch(kind: str, count: int, workers: int | None = None) -> list[bytes]:
    """count images of config `kind` ('c2' or 'c5'), generated on a process pool."""
    workers = workers or min(os.cpu_count() or 1, 64)
    if workers <= 1 or count < 4:
        return [_job((kind, i)) for i in range(count)]
    with ProcessPoolExecutor(max_workers=workers) as ex:
        return list(ex.map(_job, [(kind, i) for i in range(count)], chunksize=max(1, count // (workers * 4))))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2, choices=[2, 4, 5])
    ap.add_argument("--count", type=int, default=4)
    ap.add_argument("--out", default="/tmp/hjd_jpegs")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    if a.config == 4:
        files = [make_c4()]
    else:
        files = make_batch("c2" if a.config == 2 else "c5", a.count)
    for i, f in enumerate(files):
        with open(os.path.join(a.out, f"c{a.config}_{i:05d}.jpg"), "wb") as fh:
            fh.write(f)
    print(f"wrote {len(files)} files, mean {sum(map(len, files)) / len(files):.0f} B -> {a.out}")


if __name__ == "__main__":
    main()
