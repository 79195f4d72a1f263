"""ctypes binding of oracle/liboracle.so (oracle/jpeg_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Same result layout as oracle.refbind.decode so tests can compare the two directly.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "jpeg_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"] + (["-B"] if force else []),
                              stdout=subprocess.DEVNULL)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.hjdo_decode.restype = ctypes.c_int
        L.hjdo_decode.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p,
                                  ctypes.c_void_p, ctypes.c_void_p,
                                  ctypes.POINTER(ctypes.c_uint), ctypes.POINTER(ctypes.c_uint)]
        L.hjdo_get_image_size.restype = ctypes.c_int
        L.hjdo_get_image_size.argtypes = [ctypes.c_char_p, ctypes.c_size_t] + \
            [ctypes.POINTER(ctypes.c_uint)] * 2 + [ctypes.POINTER(ctypes.c_int)] * 4
        L.hjdo_decode_single_block.restype = None
        L.hjdo_decode_single_block.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.hjdo_ycc_to_rgb.restype = None
        L.hjdo_ycc_to_rgb.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.hjdo_idct_tables.restype = None
        L.hjdo_idct_tables.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.hjdo_bmp24_size.restype = ctypes.c_size_t
        L.hjdo_bmp24_size.argtypes = [ctypes.c_uint, ctypes.c_uint]
        L.hjdo_bmp24_encode.restype = None
        L.hjdo_bmp24_encode.argtypes = [ctypes.c_uint, ctypes.c_uint, ctypes.c_void_p, ctypes.c_void_p]
        L.hjdo_set_dc16.restype = None
        L.hjdo_set_dc16.argtypes = [ctypes.c_int]
        _lib = L
    return _lib


def set_dc16(on: bool) -> None:
    """Deviation switch (default off = the reference, loadjpg.cpp:562): search DC codes of up to 16 bits."""
    lib().hjdo_set_dc16(1 if on else 0)


def info(jpg: bytes):
    w, h = ctypes.c_uint(), ctypes.c_uint()
    nc, hf, vf, ri = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = lib().hjdo_get_image_size(jpg, len(jpg), ctypes.byref(w), ctypes.byref(h), ctypes.byref(nc),
                                   ctypes.byref(hf), ctypes.byref(vf), ctypes.byref(ri))
    if rc != 0:
        return dict(rc=rc)
    mx, my = -(-w.value // (8 * hf.value)), -(-h.value // (8 * vf.value))
    bpm = hf.value * vf.value + 2 if nc.value == 3 else 1
    return dict(rc=0, width=w.value, height=h.value, ncomp=nc.value, hf=hf.value, vf=vf.value,
                ri=ri.value, mcus_x=mx, mcus_y=my, blocks=mx * my * bpm,
                ypw=mx * 8 * hf.value, yph=my * 8 * vf.value, cpw=mx * 8, cph=my * 8)


def decode(jpg: bytes, entropy_only: bool = False, want_planes: bool = True, want_coef: bool = True):
    g = info(jpg)
    if g["rc"] != 0:
        return dict(rc=g["rc"], coef=None, planes=None, rgb=None)
    coef = np.zeros((g["blocks"], 64), dtype=np.int16) if want_coef else None
    ysz, csz = g["ypw"] * g["yph"], g["cpw"] * g["cph"]
    planes = np.zeros(ysz + 2 * csz, dtype=np.uint8) if (want_planes and not entropy_only) else None
    rgb = np.zeros((g["height"], g["width"], 3), dtype=np.uint8) if not entropy_only else None
    w, h = ctypes.c_uint(), ctypes.c_uint()
    rc = lib().hjdo_decode(jpg, len(jpg), 1 if entropy_only else 0,
                           coef.ctypes.data if coef is not None else None,
                           planes.ctypes.data if planes is not None else None,
                           rgb.ctypes.data if rgb is not None else None,
                           ctypes.byref(w), ctypes.byref(h))
    out = dict(rc=rc, width=w.value, height=h.value, coef=coef, rgb=rgb, planes=None, geometry=g)
    if planes is not None:
        out["planes"] = (planes[:ysz].reshape(g["yph"], g["ypw"]),
                         planes[ysz:ysz + csz].reshape(g["cph"], g["cpw"]),
                         planes[ysz + csz:].reshape(g["cph"], g["cpw"]))
    return out


def decode_single_block(coef64: np.ndarray, q64: np.ndarray) -> np.ndarray:
    coef64 = np.ascontiguousarray(coef64, dtype=np.int16)
    q64 = np.ascontiguousarray(q64, dtype=np.float32)
    out = np.zeros((8, 8), dtype=np.uint8)
    lib().hjdo_decode_single_block(coef64.ctypes.data, q64.ctypes.data, out.ctypes.data, 8)
    return out


def idct_tables():
    c = np.zeros((8, 8), dtype=np.float32)
    cc = np.zeros((8, 8), dtype=np.float32)
    lib().hjdo_idct_tables(c.ctypes.data, cc.ctypes.data)
    return c, cc


def bmp24_bytes(rgb: np.ndarray) -> bytes:
    h, w, _ = rgb.shape
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    n = lib().hjdo_bmp24_size(w, h)
    out = np.zeros(n, dtype=np.uint8)
    lib().hjdo_bmp24_encode(w, h, rgb.ctypes.data, out.ctypes.data)
    return out.tobytes()
