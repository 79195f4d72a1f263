"""CPU tests: the oracle (oracle/jpeg_oracle.c) against the golden vectors produced by the real
reference, and against the reference itself where oracle/_ref is built."""
import glob
import hashlib
import json
import os

import numpy as np
import pytest

from tests import cases

GOLDEN_DIR = os.path.join(os.path.dirname(__file__), "golden")
GOLDEN = json.load(open(os.path.join(GOLDEN_DIR, "golden.json")))["cases"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def golden_files():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.jpg")))


def planes_sha(planes, ncomp):
    return sha(np.concatenate([p.ravel() for p in planes][: (3 if ncomp == 3 else 1)]))


@pytest.mark.parametrize("name", golden_files())
def test_port_matches_golden(port, name):
    jpg = open(os.path.join(GOLDEN_DIR, name + ".jpg"), "rb").read()
    g = GOLDEN[name]
    assert hashlib.sha256(jpg).hexdigest() == g["jpeg_sha256"]
    r = port.decode(jpg)
    assert r["rc"] == 0
    assert (r["width"], r["height"]) == (g["width"], g["height"])
    assert r["coef"].shape[0] == g["n_blocks"]
    assert sha(r["coef"]) == g["coef_sha256"]          # bit-exact coefficients
    assert planes_sha(r["planes"], g["ncomp"]) == g["planes_sha256"]
    assert sha(r["rgb"]) == g["rgb_sha256"]            # bit-exact RGB (tolerance allowed by north_star: <= 1)


def test_fixture_generator_is_deterministic():
    """The committed fixtures are what tests/cases.py generates today (same PIL/libjpeg-turbo)."""
    regenerated = cases.small_cases()
    mismatched = [n for n, j in regenerated.items() if hashlib.sha256(j).hexdigest() != GOLDEN[n]["jpeg_sha256"]]
    if mismatched:
        pytest.skip(f"encoder differs from the one that made the fixtures: {mismatched[:3]}")


def test_port_matches_reference_lenna(port, refbind):
    path = refbind.lenna_path()
    if not path:
        pytest.skip("Lenna.jpg not available")
    jpg = open(path, "rb").read()
    g = GOLDEN["__lenna__"]
    # SURVEY.md section 4 known answers
    assert g["coef_sha256"] == "46c20f75d72e2525a21b3a0559c4fe098143cd9aed7468ac8c5ca78ec653a418"
    r = refbind.decode(jpg, mode=0)                    # unmodified JpegDecodeHW
    p = port.decode(jpg)
    assert r["rc"] == 0 and p["rc"] == 0
    assert np.array_equal(r["rgb"], p["rgb"])
    assert sha(p["coef"]) == g["coef_sha256"]
    assert sha(p["rgb"]) == g["rgb_sha256"]
    t = refbind.decode(jpg, mode=1)
    assert t["stream_index"] == 104033
    assert np.array_equal(t["coef"], p["coef"])
    for a, b in zip(t["planes"], p["planes"]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("seed", range(6))
def test_port_matches_reference_random(port, refbind, seed):
    """Seeded random geometry / quality / sampling / restart interval, port vs the real reference."""
    rng = np.random.default_rng(1000 + seed)
    w, h = int(rng.integers(1, 200)), int(rng.integers(1, 200))
    sub = ["4:4:4", "4:2:2", "4:2:0"][int(rng.integers(0, 3))]
    q = int(rng.integers(5, 101))
    ri = int(rng.choice([0, 1, 3, 8, 50]))
    gray = bool(rng.integers(0, 4) == 0)
    from tools.gen_jpegs import encode_jpeg, synth_rgb
    jpg = encode_jpeg(synth_rgb(w, h, seed, noise_sigma=float(rng.uniform(0, 30))), q, sub, ri, gray=gray,
                      optimize=bool(rng.integers(0, 2)))
    r = refbind.decode(jpg, mode=1)
    p = port.decode(jpg)
    assert r["rc"] == 0 and p["rc"] == 0
    assert np.array_equal(r["coef"], p["coef"])
    assert np.array_equal(r["rgb"], p["rgb"])


def test_restart_twin_route_a(port, refbind):
    """SURVEY.md 8c route A: the UNMODIFIED reference on the restart-free twin is the truth for the
    restart-marker file (same pixels, same tables => same quantised coefficients)."""
    from tools.gen_jpegs import encode_jpeg, synth_rgb
    rgb = synth_rgb(160, 120, 77)
    with_rst = encode_jpeg(rgb, 85, "4:2:0", 8)
    twin = encode_jpeg(rgb, 85, "4:2:0", 0)
    truth = refbind.decode(twin, mode=0)
    p = port.decode(with_rst)
    assert truth["rc"] == 0 and p["rc"] == 0
    assert np.array_equal(truth["rgb"], p["rgb"])


def test_bmp_matches_reference_writer(port, refbind, tmp_path):
    rgb = cases.noise_rgb(13, 7, 5)                    # width*3 % 4 != 0: exercises row padding
    path = str(tmp_path / "ref.bmp")
    refbind.write_bmp24(path, rgb)
    assert open(path, "rb").read() == port.bmp24_bytes(rgb)


def test_port_rejects_unsupported(port):
    assert port.decode(cases.progressive_jpeg())["rc"] != 0
    assert port.decode(b"not a jpeg at all")["rc"] != 0


def test_single_block_known_answers(port):
    """DecodeSingleBlock quirks: flat blocks whose DC*q is a multiple of 8 come out one
    level low because fl(C(0)*C(0)) = 0.49999997."""
    q = np.ones(64, dtype=np.float32)
    coef = np.zeros(64, dtype=np.int16)
    coef[0] = 8 * 10                                   # 0.125 * 80 = 10 -> 9 + 128
    assert (port.decode_single_block(coef, q) == 137).all()
    coef[0] = -80
    assert (port.decode_single_block(coef, q) == 128 - 9).all()
    coef[0] = 2000                                     # clamps
    assert (port.decode_single_block(coef, q) == 255).all()
    c, cc = port.idct_tables()
    assert cc[0, 0] == np.float32(0.49999997) and c[0, 0] == 1.0
