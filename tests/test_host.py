"""CPU tests of the product's host side: the C-ABI library builds, loads, exports every symbol
include/hjd.h declares, parses headers, writes BMPs, and refuses to decode without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from tests import cases


def declared_symbols(header_path):
    text = open(header_path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hjd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(hjd):
    L = ctypes.CDLL(hjd.LIB_PATH)
    names = declared_symbols(hjd.INCLUDE_PATH)
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert hjd.lib().hjd_version() == 100


def test_reference_cxx_names_are_exported(hjd):
    """csrc/ref_shim.cpp: the reference's own (C++-mangled) entry points."""
    import subprocess
    out = subprocess.check_output(["nm", "-DC", hjd.LIB_PATH], text=True)
    assert "ConvertJpgFile(char*, char*)" in out
    assert "DecodeJpgFileData(unsigned char const*, int, unsigned char**, unsigned int*, unsigned int*)" in out
    assert "WriteBMP24(char const*, unsigned int, unsigned int, unsigned char*)" in out


def test_image_size_from_header(hjd):
    for name, jpg in cases.small_cases().items():
        w, h = hjd.JpegGetImageSize(jpg)
        m = re.search(r"_(\d+)x(\d+)", name)
        if m:
            assert (w, h) == (int(m.group(1)), int(m.group(2))), name
    with pytest.raises(hjd.HjdError):
        hjd.JpegGetImageSize(cases.progressive_jpeg())
    with pytest.raises(hjd.HjdError):
        hjd.JpegGetImageSize(b"\xff\xd8\xff")
    with pytest.raises(hjd.HjdError):
        hjd.JpegGetImageSize(cases.cmyk_jpeg())


def test_truncated_headers_never_crash(hjd):
    jpg = cases.small_cases()["420_100x70_ri2"]
    for cut in range(0, 700, 7):
        try:
            hjd.JpegGetImageSize(jpg[:cut])
        except hjd.HjdError:
            pass


def test_bmp_bytes_match_oracle(hjd, port, tmp_path):
    for (w, h) in [(1, 1), (2, 3), (13, 7), (16, 16), (5, 9)]:
        rgb = cases.noise_rgb(w, h, w * 100 + h)
        want = port.bmp24_bytes(rgb)
        assert hjd.encode_bmp24(rgb) == want
        p = str(tmp_path / f"o_{w}x{h}.bmp")
        hjd.WriteBMP24(p, w, h, rgb)
        assert open(p, "rb").read() == want
        assert len(want) == w * h * 3 + h * ((4 - (w * 3) % 4) % 4) + 54


def test_idct_constants_match_oracle(hjd, port):
    c, cc = hjd.idct_tables()
    oc, occ = port.idct_tables()
    assert np.array_equal(c, oc) and np.array_equal(cc, occ)
    assert (c[:, 0] == 1.0).all()          # the kernels rely on cos(0) == 1 exactly


def test_no_gpu_means_loud_failure(hjd):
    if hjd.lib().hjd_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(hjd.HjdError, match="no CUDA device"):
        hjd.BatchDecoder(0)
    with pytest.raises(hjd.HjdError):
        hjd.DecodeJpgFileData(cases.small_cases()["444_1x1"])
    assert hjd.ConvertJpgFile("/nonexistent.jpg", "/tmp/x.bmp") == 0


def test_product_does_not_touch_the_oracle():
    """The shipped package must not import, include, link or execute anything under oracle/."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "hls_jpeg_decoder_b200")
    bad = re.compile(r"^\s*(from|import)\s+oracle\b|#\s*include[^\n]*oracle|liboracle|jpeg_oracle|hjdo_|hjdref_|CDLL\([^)]*oracle",
                     re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not bad.search(text), (dirpath, f, bad.search(text).group(0))
