"""ctypes binding of oracle/_ref/libhjdref_*.so -- TEST INFRASTRUCTURE ONLY.

The .so files are the UNMODIFIED reference (harutel/hls-jpeg-decoder) compiled by
oracle/build_ref.sh around oracle/ref_harness.cpp.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module; the product
package never does.
"""
from __future__ import annotations

import ctypes
import os
import struct

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
_LIBS: dict[str, ctypes.CDLL] = {}


def available(variant: str = "hd") -> bool:
    return os.path.exists(os.path.join(REF_DIR, f"libhjdref_{variant}.so"))


def lenna_path() -> str | None:
    p = os.path.join(REF_DIR, "data", "Lenna.jpg")
    return p if os.path.exists(p) else None


def _lib(variant: str) -> ctypes.CDLL:
    lib = _LIBS.get(variant)
    if lib is None:
        lib = ctypes.CDLL(os.path.join(REF_DIR, f"libhjdref_{variant}.so"))
        lib.hjdref_decode.restype = ctypes.c_int
        lib.hjdref_decode.argtypes = [
            ctypes.c_char_p, ctypes.c_int, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.POINTER(ctypes.c_uint), ctypes.POINTER(ctypes.c_uint), ctypes.POINTER(ctypes.c_uint),
        ]
        lib.hjdref_write_bmp24.restype = None
        lib.hjdref_write_bmp24.argtypes = [ctypes.c_char_p, ctypes.c_uint, ctypes.c_uint, ctypes.c_void_p]
        for f in ("hjdref_max_width", "hjdref_max_height", "hjdref_stream_size"):
            getattr(lib, f).restype = ctypes.c_int
        _LIBS[variant] = lib
    return lib


def sniff(jpg: bytes):
    """(width, height, ncomp, hF, vF, Ri) by a minimal marker walk (for buffer sizing only)."""
    i, w, h, nc, hf, vf, ri = 2, 0, 0, 0, 1, 1, 0
    while i + 4 <= len(jpg):
        if jpg[i] != 0xFF:
            break
        while jpg[i] == 0xFF:
            i += 1
        m = jpg[i]
        i += 1
        if m in (0xD8, 0xD9, 0x01) or 0xD0 <= m <= 0xD7:
            continue
        (ln,) = struct.unpack(">H", jpg[i:i + 2])
        if m in (0xC0, 0xC1, 0xC2):
            h, w = struct.unpack(">HH", jpg[i + 3:i + 7])
            nc = jpg[i + 7]
            hf, vf = jpg[i + 9] >> 4, jpg[i + 9] & 15
        elif m == 0xDD:
            (ri,) = struct.unpack(">H", jpg[i + 2:i + 4])
        elif m == 0xDA:
            break
        i += ln
    if nc == 1:
        hf = vf = 1
    return w, h, nc, hf, vf, ri


def pick_variant(w: int, h: int, size: int) -> str:
    if w <= 512 and h <= 512 and size <= 105 * 1000:
        return "std"
    if w <= 1920 and h <= 1088 and size <= 2000 * 1000:
        return "hd"
    return "big"


def geometry(jpg: bytes):
    w, h, nc, hf, vf, _ = sniff(jpg)
    mx, my = -(-w // (8 * hf)), -(-h // (8 * vf))
    bpm = hf * vf + 2 if nc == 3 else 1
    return dict(width=w, height=h, ncomp=nc, hf=hf, vf=vf, mcus_x=mx, mcus_y=my,
                blocks=mx * my * bpm, ypw=mx * 8 * hf, yph=my * 8 * vf, cpw=mx * 8, cph=my * 8)


def decode(jpg: bytes, mode: int = 1, variant: str | None = None, want_planes: bool = True):
    """Run the reference.  Returns dict(rc, width, height, coef[int16 nblocks x 64] | None,
    planes (Y, Cb, Cr) | None, rgb[h, w, 3], stream_index)."""
    g = geometry(jpg)
    variant = variant or pick_variant(g["width"], g["height"], len(jpg))
    lib = _lib(variant)
    coef = np.zeros((g["blocks"], 64), dtype=np.int16) if mode == 1 else None
    ysz, csz = g["ypw"] * g["yph"], g["cpw"] * g["cph"]
    planes = np.zeros(ysz + 2 * csz, dtype=np.uint8) if (mode == 1 and want_planes) else None
    rgb = np.zeros((g["height"], g["width"], 3), dtype=np.uint8)
    w, h, si = ctypes.c_uint(), ctypes.c_uint(), ctypes.c_uint()
    rc = lib.hjdref_decode(jpg, len(jpg), mode,
                           coef.ctypes.data if coef is not None else None,
                           planes.ctypes.data if planes is not None else None,
                           rgb.ctypes.data, ctypes.byref(w), ctypes.byref(h), ctypes.byref(si))
    out = dict(rc=rc, width=w.value, height=h.value, coef=coef, rgb=rgb, stream_index=si.value, planes=None)
    if planes is not None:
        out["planes"] = (planes[:ysz].reshape(g["yph"], g["ypw"]),
                         planes[ysz:ysz + csz].reshape(g["cph"], g["cpw"]),
                         planes[ysz + csz:].reshape(g["cph"], g["cpw"]))
    return out


def write_bmp24(path: str, rgb: np.ndarray) -> None:
    h, w, _ = rgb.shape
    rgb = np.ascontiguousarray(rgb)
    _lib(pick_variant(w, h, 0)).hjdref_write_bmp24(path.encode(), w, h, rgb.ctypes.data)
