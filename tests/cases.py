"""Seeded JPEG test cases shared by the CPU and GPU tests (PIL / libjpeg-turbo encoder)."""
from __future__ import annotations

import io

import numpy as np
from PIL import Image

from tools.gen_jpegs import encode_jpeg, synth_rgb


def noise_rgb(w, h, seed):
    return np.random.default_rng(seed).integers(0, 256, size=(h, w, 3), dtype=np.uint8)


def flat_rgb(w, h, value):
    return np.full((h, w, 3), value, dtype=np.uint8)


def gradient_rgb(w, h):
    x = np.linspace(0, 255, w, dtype=np.float32)[None, :, None]
    y = np.linspace(0, 255, h, dtype=np.float32)[:, None, None]
    img = np.concatenate([np.broadcast_to(x, (h, w, 1)), np.broadcast_to(y, (h, w, 1)),
                          np.broadcast_to((x + y) / 2, (h, w, 1))], axis=2)
    return img.astype(np.uint8)


def encode_jpeg_cv(rgb, quality=85, sampling="440", restart=0):
    """OpenCV / libjpeg-turbo encoder: the only one here that writes 4:4:0 (luma 1x2) files."""
    import cv2
    params = [cv2.IMWRITE_JPEG_QUALITY, quality, cv2.IMWRITE_JPEG_SAMPLING_FACTOR,
              getattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR_" + sampling)]
    if restart:
        params += [cv2.IMWRITE_JPEG_RST_INTERVAL, restart]
    ok, buf = cv2.imencode(".jpg", np.ascontiguousarray(rgb[:, :, ::-1]), params)
    assert ok
    return buf.tobytes()


def small_cases():
    """name -> jpeg bytes; every sampling mode, odd sizes, restart intervals, table kinds, qualities."""
    c = {}
    c["444_64x48_q85"] = encode_jpeg(synth_rgb(64, 48, 1), 85, "4:4:4")
    c["422_64x48_q85"] = encode_jpeg(synth_rgb(64, 48, 2), 85, "4:2:2")
    c["420_64x48_q85"] = encode_jpeg(synth_rgb(64, 48, 3), 85, "4:2:0")
    c["420_37x53_q75"] = encode_jpeg(synth_rgb(37, 53, 4), 75, "4:2:0")
    c["444_1x1"] = encode_jpeg(synth_rgb(1, 1, 5), 90, "4:4:4")
    c["420_1x1"] = encode_jpeg(synth_rgb(1, 1, 6), 90, "4:2:0")
    c["420_17x16_ri1"] = encode_jpeg(synth_rgb(17, 16, 7), 85, "4:2:0", restart_blocks=1)
    c["420_100x70_ri2"] = encode_jpeg(synth_rgb(100, 70, 8), 85, "4:2:0", restart_blocks=2)
    c["422_100x70_ri3"] = encode_jpeg(synth_rgb(100, 70, 9), 60, "4:2:2", restart_blocks=3)
    c["444_100x70_ri8"] = encode_jpeg(synth_rgb(100, 70, 10), 85, "4:4:4", restart_blocks=8)
    c["420_100x70_ri1000"] = encode_jpeg(synth_rgb(100, 70, 11), 85, "4:2:0", restart_blocks=1000)
    c["440_100x70_q85"] = encode_jpeg_cv(synth_rgb(100, 70, 21), 85, "440")
    c["440_61x35_ri3"] = encode_jpeg_cv(synth_rgb(61, 35, 22), 70, "440", restart=3)
    c["gray_64x64"] = encode_jpeg(synth_rgb(64, 64, 12), 75, gray=True)
    c["gray_33x9_ri4"] = encode_jpeg(synth_rgb(33, 9, 13), 75, gray=True, restart_blocks=4)
    c["420_opt_96x96"] = encode_jpeg(synth_rgb(96, 96, 14), 85, "4:2:0", optimize=True)
    c["444_opt_ri5"] = encode_jpeg(synth_rgb(80, 40, 15), 92, "4:4:4", restart_blocks=5, optimize=True)
    c["420_noise_q100"] = encode_jpeg(noise_rgb(64, 64, 16), 100, "4:2:0")
    c["444_noise_q100_ri2"] = encode_jpeg(noise_rgb(48, 40, 17), 100, "4:4:4", restart_blocks=2)
    c["420_noise_q10"] = encode_jpeg(noise_rgb(64, 64, 18), 10, "4:2:0")
    c["420_flat128"] = encode_jpeg(flat_rgb(64, 64, 128), 85, "4:2:0")
    c["420_flat_dark_ri1"] = encode_jpeg(flat_rgb(40, 24, 17), 85, "4:2:0", restart_blocks=1)
    c["444_gradient_q95"] = encode_jpeg(gradient_rgb(128, 96), 95, "4:4:4")
    c["420_gradient_q50"] = encode_jpeg(gradient_rgb(128, 96), 50, "4:2:0")
    c["420_white_black"] = encode_jpeg(np.kron(np.indices((8, 8)).sum(0) % 2, np.ones((8, 8)))[..., None]
                                       .repeat(3, 2).astype(np.uint8) * 255, 90, "4:2:0")
    return c


def progressive_jpeg():
    buf = io.BytesIO()
    Image.fromarray(synth_rgb(64, 64, 99)).save(buf, format="JPEG", quality=80, progressive=True)
    return buf.getvalue()


def cmyk_jpeg():
    buf = io.BytesIO()
    Image.fromarray(synth_rgb(32, 32, 98)).convert("CMYK").save(buf, format="JPEG", quality=80)
    return buf.getvalue()
