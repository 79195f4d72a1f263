"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI
(libhjd.so via ctypes); the oracle is only the checker.

Bars (BASELINE.json north_star): entropy-decoded coefficients bit-exact; RGB max |diff| <= 1 per
channel.  This implementation aims higher -- planes and RGB identical -- and the tests assert
that, reporting mismatch counts when it does not hold.
"""
import glob
import hashlib
import json
import os

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu

GOLDEN_DIR = os.path.join(os.path.dirname(__file__), "golden")
GOLDEN = json.load(open(os.path.join(GOLDEN_DIR, "golden.json")))["cases"]
RGB_TOLERANCE = 1      # north_star: max |diff| <= 1 per channel versus the reference's float IDCT


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def golden_files():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.jpg")))


@pytest.fixture(scope="module", params=["planes", "fused_mcu", "fused_mcu_tensor_core", "fused_mcu_cuda_core"])
def dec(hjd, request):
    """fused_mcu: the default product path (kernels 2+3 fused per MCU, planes in shared memory only; the IDCT's fast tier on
    the tensor cores for large colour images, on the CUDA cores for small ones); fused_mcu_tensor_core / fused_mcu_cuda_core:
    one of the two for every image (HJD_FLAG_TENSOR_CORE_IDCT, csrc/mcu_tc.cuh / HJD_FLAG_CUDA_CORE_IDCT, csrc/kernels.cu);
    planes: HJD_FLAG_KEEP_PLANES, unfused kernels 2 and 3 with the Y/Cb/Cr planes in HBM (parity tap)."""
    d = hjd.BatchDecoder(0, {"fused_mcu": 0, "fused_mcu_cuda_core": hjd.FLAG_CUDA_CORE_IDCT, "fused_mcu_tensor_core": hjd.FLAG_TENSOR_CORE_IDCT,
                             "planes": hjd.FLAG_KEEP_PLANES}[request.param])
    d.keeps_planes = request.param == "planes"
    yield d
    d.close()


def compare_image(dec, i, oracle, all_coef, slab, name):
    inf = dec.info(i)
    coef = dec.image_coefficients(i, all_coef)
    assert coef.shape == oracle["coef"].shape, name
    bad_blocks = int((coef != oracle["coef"]).any(axis=1).sum())
    assert bad_blocks == 0, f"{name}: {bad_blocks} of {coef.shape[0]} coefficient blocks differ"
    if slab is not None:
        planes = dec.planes(i, slab)
        for pname, a, b in zip("Y Cb Cr".split(), planes, oracle["planes"]):
            if a is None:
                continue
            diff = int((a != b).sum())
            assert diff == 0, f"{name}: plane {pname}: {diff} of {a.size} samples differ"
    rgb = dec.rgb(i)
    assert rgb.shape == oracle["rgb"].shape
    d = np.abs(rgb.astype(np.int16) - oracle["rgb"].astype(np.int16))
    mism = int((d != 0).sum())
    assert d.max(initial=0) <= RGB_TOLERANCE, f"{name}: max RGB diff {d.max()} ({mism} samples differ)"
    assert mism == 0, f"{name}: {mism} RGB samples differ (max {d.max()})"
    return inf


def test_golden_batch(dec, port):
    """All committed fixtures in ONE batch (mixed geometry, sampling, tables, restart intervals)."""
    names = golden_files()
    files = [open(os.path.join(GOLDEN_DIR, n + ".jpg"), "rb").read() for n in names]
    dec.upload(files)
    dec.decode()
    st = dec.status()
    assert (st == 0).all(), dict(zip(names, st.tolist()))
    all_coef, slab = dec.coefficients(), (dec.plane_slab() if dec.keeps_planes else None)
    for i, (n, f) in enumerate(zip(names, files)):
        g = GOLDEN[n]
        coef = dec.image_coefficients(i, all_coef)
        assert sha(coef) == g["coef_sha256"], n        # golden made by the real reference
        assert sha(dec.rgb(i)) == g["rgb_sha256"], n
        compare_image(dec, i, port.decode(f), all_coef, slab, n)


def test_lenna_config1(hjd, dec, port, tmp_path):
    """Config 1: data/Lenna.jpg -> BMP, coefficient sha256 from SURVEY.md section 4."""
    from oracle import refbind
    path = refbind.lenna_path()
    if not path:
        pytest.skip("oracle/_ref/data/Lenna.jpg not present")
    jpg = open(path, "rb").read()
    dec.upload([jpg])
    dec.decode()
    assert dec.status()[0] == 0
    coef = dec.coefficients()
    assert sha(coef) == "46c20f75d72e2525a21b3a0559c4fe098143cd9aed7468ac8c5ca78ec653a418"
    o = port.decode(jpg)
    compare_image(dec, 0, o, coef, dec.plane_slab() if dec.keeps_planes else None, "lenna")
    assert sha(dec.rgb(0)) == GOLDEN["__lenna__"]["rgb_sha256"]
    # the drop-in entry points
    out = str(tmp_path / "out.bmp")
    assert hjd.ConvertJpgFile(path, out) == 1
    bmp = open(out, "rb").read()
    assert len(bmp) == 786486
    assert bmp == port.bmp24_bytes(o["rgb"])
    rgb, w, h = hjd.DecodeJpgFileData(jpg)
    assert (w, h) == (512, 512) and np.array_equal(rgb, o["rgb"])


@pytest.mark.parametrize("seed", range(8))
def test_random_images(dec, port, seed):
    from tools.gen_jpegs import encode_jpeg, synth_rgb
    rng = np.random.default_rng(2000 + seed)
    files, names = [], []
    for k in range(6):
        w, h = int(rng.integers(1, 400)), int(rng.integers(1, 300))
        sub = ["4:4:4", "4:2:2", "4:2:0"][int(rng.integers(0, 3))]
        q = int(rng.integers(5, 101))
        ri = int(rng.choice([0, 0, 1, 2, 7, 8, 64]))
        gray = bool(rng.integers(0, 5) == 0)
        opt = bool(rng.integers(0, 2))
        sigma = float(rng.choice([0, 2, 6, 25]))
        files.append(encode_jpeg(synth_rgb(w, h, 100 * seed + k, noise_sigma=sigma), q, sub, ri, gray=gray, optimize=opt))
        names.append(f"s{seed}k{k}_{w}x{h}_{sub}_q{q}_ri{ri}_g{int(gray)}_o{int(opt)}")
    dec.upload(files)
    dec.decode()
    assert (dec.status() == 0).all()
    all_coef, slab = dec.coefficients(), (dec.plane_slab() if dec.keeps_planes else None)
    for i, (n, f) in enumerate(zip(names, files)):
        compare_image(dec, i, port.decode(f), all_coef, slab, n)


def test_config2_1080p_restart8(dec, port):
    """Config 2 geometry: 1920x1080 4:2:0 q85 Ri=8 (two images), full pixel parity."""
    from tools.gen_jpegs import make_c2
    files = [make_c2(0), make_c2(1)]
    dec.upload(files)
    dec.decode()
    assert (dec.status() == 0).all()
    inf = dec.info(0)
    assert (inf.mcus_x, inf.mcus_y, inf.n_intervals, inf.n_blocks) == (120, 68, 1020, 48960)
    all_coef, slab = dec.coefficients(), (dec.plane_slab() if dec.keeps_planes else None)
    for i, f in enumerate(files):
        compare_image(dec, i, port.decode(f), all_coef, slab, f"c2_{i}")


def test_restart_twin_property(dec):
    """Size-independent property: a restart-marker file and its restart-free twin (same pixels)
    decode to identical coefficients and pixels."""
    from tools.gen_jpegs import make_c2
    a, b = make_c2(3, restart=True, width=640, height=360), make_c2(3, restart=False, width=640, height=360)
    dec.upload([a, b])
    dec.decode()
    assert (dec.status() == 0).all()
    c = dec.coefficients()
    assert np.array_equal(dec.image_coefficients(0, c), dec.image_coefficients(1, c))
    assert np.array_equal(dec.rgb(0), dec.rgb(1))


def test_batch_results_independent_of_batch_composition(dec):
    """Per-image outputs do not depend on what else is in the batch, nor on the order."""
    files = list(cases.small_cases().values())[:10]
    dec.upload(files)
    dec.decode()
    a = [dec.rgb(i).copy() for i in range(len(files))]
    dec.upload(files[::-1])
    dec.decode()
    b = [dec.rgb(i).copy() for i in range(len(files))][::-1]
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    dec.upload([files[4]])
    dec.decode()
    assert np.array_equal(dec.rgb(0), a[4])


def test_bad_inputs_do_not_poison_the_batch(dec, port):
    good = cases.small_cases()["420_64x48_q85"]
    trunc = cases.small_cases()["420_100x70_ri2"]
    files = [good, b"garbage", cases.progressive_jpeg(), good[:200], trunc[:len(trunc) // 2], cases.cmyk_jpeg(), good]
    dec.upload(files)
    dec.decode()
    st = dec.status()
    assert st[0] == 0 and st[6] == 0
    assert st[1] < 0 and st[2] < 0 and st[3] < 0 and st[5] < 0
    assert st[4] > 0                                    # decodes what is there, flags the rest
    o = port.decode(good)
    assert np.array_equal(dec.rgb(0), o["rgb"]) and np.array_equal(dec.rgb(6), o["rgb"])


def test_corrupt_entropy_data_terminates(dec):
    """Random bytes as entropy data: every loop is bounded (the reference can spin forever,
    loadjpg.cpp:700-829) and the call returns with a status."""
    jpg = bytearray(cases.small_cases()["420_100x70_ri2"])
    rng = np.random.default_rng(5)
    start = len(jpg) // 2
    jpg[start:-2] = rng.integers(0, 255, size=len(jpg) - 2 - start, dtype=np.uint8).tobytes()
    dec.upload([bytes(jpg)])
    dec.decode()
    dec.status()


def test_host_scan_flag_matches_gpu_scan(hjd, dec):
    files = [cases.small_cases()[k] for k in ("420_100x70_ri2", "444_100x70_ri8", "gray_33x9_ri4", "420_64x48_q85")]
    dec.upload(files)
    dec.decode()
    a = [dec.rgb(i).copy() for i in range(len(files))]
    with hjd.BatchDecoder(0, hjd.FLAG_HOST_SCAN) as d2:
        d2.upload(files)
        d2.decode()
        assert (d2.status() == 0).all()
        for i in range(len(files)):
            assert np.array_equal(a[i], d2.rgb(i))


def test_decode_host_end_to_end(hjd, port):
    files = [cases.small_cases()[k] for k in ("420_100x70_ri2", "444_64x48_q85", "gray_64x64", "420_37x53_q75")]
    arena = hjd.PinnedArena(files)
    need = hjd.rgb_slab_bytes(arena)
    out = np.zeros(need, dtype=np.uint8)
    with hjd.BatchDecoder(0) as d:
        offs, st = d.decode_host(arena, out.ctypes.data, need, chunk_images=3)
    assert (st == 0).all()
    for f, off in zip(files, offs):
        o = port.decode(f)
        n = o["rgb"].size
        assert np.array_equal(out[int(off):int(off) + n].reshape(o["rgb"].shape), o["rgb"])
    arena.close()


@pytest.mark.parametrize("flags", [0, "tensor_core"])
def test_chunked_overlapped_execution_matches_serial(hjd, port, flags):
    """Multi-stream chunked execution (forced with a tiny blocks-per-chunk target) gives the same
    bytes as the serial path, resident and host-buffer variants -- also with the tensor-core kernel on every
    chunk (persistent CTAs that hold all of an SM's TMEM and shared memory, launched from three streams)."""
    files = list(cases.small_cases().values())
    with hjd.BatchDecoder(0, hjd.FLAG_TENSOR_CORE_IDCT if flags else 0) as d:
        d.set_overlap(0)
        d.upload(files)
        d.decode()
        want = [d.rgb(i).copy() for i in range(len(files))]
        want_coef = d.coefficients().copy()
        d.set_overlap(40)                      # ~40 blocks per chunk -> 8 chunks on 3 streams
        d.upload(files)
        d.decode()
        assert (d.status() == 0).all()
        assert np.array_equal(d.coefficients(), want_coef)
        for i in range(len(files)):
            assert np.array_equal(d.rgb(i), want[i])
        arena = hjd.PinnedArena(files)
        need = hjd.rgb_slab_bytes(arena)
        out = np.zeros(need, dtype=np.uint8)
        offs, st = d.decode_host(arena, out.ctypes.data, need)
        assert (st == 0).all()
        for i, off in enumerate(offs):
            n = want[i].size
            assert np.array_equal(out[int(off):int(off) + n].reshape(want[i].shape), want[i])
        arena.close()


def test_selfsync_matches_single_thread_path(hjd, port):
    """Restart-free scans: kernel 1b (speculative self-synchronising decode) against kernel 1a run
    with one thread per scan (HJD_FLAG_NO_SELFSYNC) and against the oracle."""
    from tools.gen_jpegs import encode_jpeg, synth_rgb
    files = [encode_jpeg(synth_rgb(320, 200, 41), 85, "4:4:4"),
             encode_jpeg(synth_rgb(333, 217, 42), 60, "4:2:0"),
             encode_jpeg(synth_rgb(256, 256, 43), 95, "4:2:2", optimize=True),
             encode_jpeg(synth_rgb(200, 300, 44), 75, gray=True),
             encode_jpeg(cases.noise_rgb(160, 160, 45), 100, "4:4:4"),
             encode_jpeg(cases.flat_rgb(512, 512, 90), 85, "4:2:0"),
             encode_jpeg(synth_rgb(300, 200, 46), 85, "4:2:0", restart_blocks=8)]     # one restart image in the mix
    with hjd.BatchDecoder(0) as d1, hjd.BatchDecoder(0, hjd.FLAG_NO_SELFSYNC) as d2:
        for d in (d1, d2):
            d.upload(files)
            d.decode()
            assert (d.status() == 0).all(), d.status()
        assert d1.selfsync_rounds >= 2 and d2.selfsync_rounds == 0
        c1, c2 = d1.coefficients(), d2.coefficients()
        assert np.array_equal(c1, c2)
        for i, f in enumerate(files):
            o = port.decode(f)
            assert np.array_equal(d1.image_coefficients(i, c1), o["coef"]), i
            assert np.array_equal(d1.rgb(i), o["rgb"]), i


def test_selfsync_large_scan(hjd, port):
    """Config 4 geometry scaled down: one 2048x2048 4:4:4 restart-free image (196,608 blocks),
    coefficients bit-exact against the oracle's entropy stage, RGB against the full oracle."""
    from tools.gen_jpegs import make_c4
    jpg = make_c4(2048, seed=4)
    with hjd.BatchDecoder(0) as d:
        d.upload([jpg])
        d.decode()
        assert d.status()[0] == 0
        inf = d.info(0)
        assert (inf.restart_interval, inf.n_intervals, inf.n_blocks) == (0, 0, 196608)
        o = port.decode(jpg)
        assert np.array_equal(d.coefficients(), o["coef"])
        assert np.array_equal(d.rgb(0), o["rgb"])
        assert d.selfsync_rounds <= 6, d.selfsync_rounds


def test_selfsync_ranges_and_twins(hjd, port):
    """Restart-free 1080p twins of config 2 (the stream shape the fix-up rounds were tuned on): every
    range length of the synchronisation rounds gives the coefficients of the restart twins (decoded by
    kernel 1a), and image 0 matches the oracle."""
    from tools.gen_jpegs import make_c2
    n = 6
    twins = [make_c2(i, restart=False) for i in range(n)]
    with hjd.BatchDecoder(0) as d:
        d.upload([make_c2(i, restart=True) for i in range(n)])
        d.decode()
        assert (d.status() == 0).all()
        want = d.coefficients().copy()
        want_rgb = d.rgb_slab().copy()
        for rng in (0, 32, 64, 128, 256):
            d.set_selfsync_range(rng)
            d.upload(twins)
            d.decode()
            assert (d.status() == 0).all(), (rng, d.status())
            assert d.selfsync_rounds >= 2
            assert np.array_equal(d.coefficients(), want), rng
            assert np.array_equal(d.rgb_slab(), want_rgb), rng
        o = port.decode(twins[0])
        assert np.array_equal(d.image_coefficients(0), o["coef"])
        assert np.array_equal(d.rgb(0), o["rgb"])
        with pytest.raises(Exception):
            d.set_selfsync_range(48)


@pytest.mark.timeout(600)
def test_selfsync_damaged_scans_are_deterministic(hjd):
    """Truncated and corrupted restart-free scans through kernel 1b: the call returns with a per-image
    warning, good neighbours are untouched, and the output depends on the input alone (blocks the
    scan never reached are zero, whatever the slab held before)."""
    from tools.gen_jpegs import encode_jpeg, synth_rgb
    rng = np.random.default_rng(7)
    good = encode_jpeg(synth_rgb(640, 480, 51), 85, "4:2:0")
    other = encode_jpeg(cases.noise_rgb(640, 480, 52), 95, "4:4:4")
    eoi = good[-2:]
    bad = [good[:len(good) * 2 // 5] + eoi,                    # truncated at 40 %
           good[:len(good) - 700] + eoi]                       # truncated inside the last sub-sequences
    for k in range(6):                                         # random damage inside the scan
        b = bytearray(good)
        for _ in range(int(rng.integers(1, 6))):
            pos = int(rng.integers(len(b) // 3, len(b) - 2))
            mode = int(rng.integers(0, 3))
            if mode == 0:
                b[pos] = int(rng.integers(0, 256))
            elif mode == 1:
                b[pos:pos + 2] = b"\xff\x00"
            else:
                del b[pos:pos + int(rng.integers(1, 300))]
        bad.append(bytes(b))
    with hjd.BatchDecoder(0) as d:
        d.upload([good])
        d.decode()
        want = d.rgb(0).copy()
        outs = []
        for rep in range(2):
            d.upload([other, other, other])                    # different content in the slabs in between
            d.decode()
            files = [good] + bad + [good]
            d.upload(files)
            d.decode()
            st = d.status()
            assert st[0] == 0 and st[-1] == 0, st
            assert st[1] > 0 and (st[1] & 4), st                # HJD_IMG_WARN_OVERRUN on the truncated one
            assert np.array_equal(d.rgb(0), want) and np.array_equal(d.rgb(len(files) - 1), want)
            outs.append((st.copy(), d.coefficients().copy()))
        assert np.array_equal(outs[0][0], outs[1][0])
        assert np.array_equal(outs[0][1], outs[1][1])
        inf = d.info(1)
        tail = d.image_coefficients(1, outs[1][1])[inf.n_blocks * 3 // 5:]
        assert not tail.any()


def test_selfsync_padding_after_the_last_block_is_not_an_error(hjd, port):
    """Found by tools/soak_parity.py (2 of 40,000 images): when the last block ends a few bytes before
    the end of the last sub-sequence, the thread of that sub-sequence is entered in the padding after
    the image; skipping it must not raise HJD_IMG_WARN_BAD_CODE."""
    from tools.soak_parity import make
    files = [make(10674, 3), make(25358, 3)]
    with hjd.BatchDecoder(0) as d:
        d.upload(files)
        d.decode()
        assert (d.status() == 0).all(), d.status()
        for i, f in enumerate(files):
            o = port.decode(f)
            assert np.array_equal(d.image_coefficients(i), o["coef"]), i
            assert np.array_equal(d.rgb(i), o["rgb"]), i


def test_long_restart_scans_use_the_sliced_marker_scan(hjd):
    """Scans above 1 MB are cut into 64 KB slices by kernel 0 (count, then number the markers): the
    interval table must be what the host scan finds, i.e. the same RGB, for a clean image, for one
    with a marker removed and in a batch with small images on both sides."""
    from tools.gen_jpegs import encode_jpeg, synth_rgb
    big = encode_jpeg(synth_rgb(3072, 2304, 81), 92, "4:2:0", 8)
    inf_len = len(big)
    assert inf_len > (1 << 20) + 4096
    small = cases.small_cases()["420_100x70_ri2"]
    broken = bytearray(big)
    k = broken.find(b"\xff\xd3", len(broken) // 2)
    assert k > 0
    del broken[k:k + 2]                                        # one RSTn gone
    files = [small, big, small, bytes(broken), small]
    with hjd.BatchDecoder(0) as d1, hjd.BatchDecoder(0, hjd.FLAG_HOST_SCAN) as d2:
        for d in (d1, d2):
            d.upload(files)
            d.decode()
        s1, s2 = d1.status(), d2.status()
        assert s1[0] == 0 and s1[1] == 0 and s1[2] == 0 and s1[4] == 0, s1
        assert s1[3] > 0 and (s1[3] & 8), s1                    # HJD_IMG_WARN_RESTART
        assert d1.info(1).scan_bytes > (1 << 20)
        for i in (0, 1, 2, 4):
            assert np.array_equal(d1.rgb(i), d2.rgb(i)), i
        assert np.array_equal(d1.coefficients()[:d1.info(2).block_base + d1.info(2).n_blocks],
                              d2.coefficients()[:d2.info(2).block_base + d2.info(2).n_blocks])


def _patch_dqt(jpg: bytes, value: int) -> bytes:
    """Overwrite every 8-bit quantisation table entry with `value`."""
    b = bytearray(jpg)
    i = 2
    while i + 4 < len(b):
        assert b[i] == 0xFF
        m = b[i + 1]
        ln = (b[i + 2] << 8) | b[i + 3]
        if m == 0xDB:
            p = i + 4
            while p < i + 2 + ln:
                assert b[p] >> 4 == 0
                b[p + 1:p + 65] = bytes([value]) * 64
                p += 65
        if m == 0xDA:
            break
        i += 2 + ln
    return bytes(b)


def test_absurd_dequantisation_wraps_like_the_reference(hjd, port):
    """Large coefficients (a q100 noise image) against quantisation tables patched to 255 and to 64:
    coef*q overflows the reference's short (loadjpg.cpp:150), the IDCT sums reach 1e6 and wrap again
    (loadjpg.cpp:136-137).  Every sample then goes through the exact path; the result must still be
    the reference's, bit for bit."""
    from tools.gen_jpegs import encode_jpeg
    base = [encode_jpeg(cases.noise_rgb(96, 80, 71), 100, "4:4:4"),
            encode_jpeg(cases.noise_rgb(112, 64, 72), 100, "4:2:0", restart_blocks=4)]
    files = [_patch_dqt(f, v) for f in base for v in (255, 64, 17)]
    # the tensor-core kernel takes its exact tier for all 64 samples of such blocks (quantised coefficients beyond +-511,
    # products beyond +-2047, A >= 4000): the preconditions of its fast tier, csrc/mcu_tc.cuh
    with hjd.BatchDecoder(0, hjd.FLAG_KEEP_PLANES) as d1, hjd.BatchDecoder(0) as d2, \
            hjd.BatchDecoder(0, hjd.FLAG_TENSOR_CORE_IDCT) as d3:
        for d in (d1, d2, d3):
            d.upload(files)
            d.decode()
            assert (d.status() == 0).all(), d.status()
        for i, f in enumerate(files):
            o = port.decode(f)
            assert np.array_equal(d1.image_coefficients(i), o["coef"]), i
            for a, b in zip(d1.planes(i), o["planes"]):
                assert np.array_equal(a, b), i
            assert np.array_equal(d1.rgb(i), o["rgb"]), i
            assert np.array_equal(d2.rgb(i), o["rgb"]), i
            assert np.array_equal(d3.rgb(i), o["rgb"]), i


def _patch_component_ids(jpg: bytes, ids):
    """Rewrite the component identifiers in SOF0 and SOS (the reference indexes arrays with them,
    openjpg.cpp:212-213,343-345, so it only works for 1,2,3)."""
    b = bytearray(jpg)
    i = 2
    while i + 4 < len(b):
        assert b[i] == 0xFF
        m = b[i + 1]
        ln = (b[i + 2] << 8) | b[i + 3]
        if m == 0xC0:
            for c in range(b[i + 9]):
                b[i + 10 + 3 * c] = ids[c]
        if m == 0xDA:
            for c in range(b[i + 4]):
                b[i + 5 + 2 * c] = ids[c]
            break
        i += 2 + ln
    return bytes(b)


def _move_dqt_after_sof(jpg: bytes):
    """Reorder segments so that the DQT segments follow SOF0 (the reference copies the tables at
    SOF time, openjpg.cpp:347-350, and would de-quantise with zeros)."""
    segs, i = [], 2
    while True:
        m = jpg[i + 1]
        ln = (jpg[i + 2] << 8) | jpg[i + 3]
        if m == 0xDA:
            tail = jpg[i:]
            break
        segs.append((m, jpg[i:i + 2 + ln]))
        i += 2 + ln
    dqt = [s for m, s in segs if m == 0xDB]
    rest = [(m, s) for m, s in segs if m != 0xDB]
    out = bytearray(jpg[:2])
    for m, s in rest:
        out += s
        if m == 0xC0:
            for q in dqt:
                out += q
    return bytes(out + tail)


def test_inputs_the_reference_cannot_decode(hjd, port):
    """component ids other than 1,2,3, DQT after SOF, extra APPn/COM segments and
    fill bytes before markers.  The oracle port handles them by construction (components by position,
    tables resolved at SOS); the GPU path must agree with it bit for bit."""
    base = cases.small_cases()
    a = _patch_component_ids(base["420_100x70_ri2"], [0, 1, 2])
    b = _patch_component_ids(base["444_64x48_q85"], [82, 71, 66])
    c = _move_dqt_after_sof(base["420_64x48_q85"])
    d0 = base["422_64x48_q85"]
    d = d0[:2] + b"\xff\xfe\x00\x07hello" + b"\xff\xff\xff\xe5\x00\x04ab" + d0[2:]      # COM, fill bytes + APP5
    files = [a, b, c, d]
    with hjd.BatchDecoder(0) as dec:
        dec.upload(files)
        dec.decode()
        assert (dec.status() == 0).all(), dec.status()
        coef = dec.coefficients()
        for i, f in enumerate(files):
            o = port.decode(f)
            assert o["rc"] == 0
            assert np.array_equal(dec.image_coefficients(i, coef), o["coef"]), i
            assert np.array_equal(dec.rgb(i), o["rgb"]), i
    # same pixels as the untouched files
    assert np.array_equal(port.decode(a)["rgb"], port.decode(base["420_100x70_ri2"])["rgb"])
    assert np.array_equal(port.decode(c)["rgb"], port.decode(base["420_64x48_q85"])["rgb"])


def test_convert_jpg_files_batch(hjd, port, tmp_path):
    """Batch-scale ConvertJpgFile: batched loader, one decode, parallel BMP writers; a missing file and
    a non-JPEG do not stop the rest."""
    names = ["420_100x70_ri2", "444_64x48_q85", "gray_64x64", "420_37x53_q75", "444_1x1"]
    ins, outs = [], []
    for n in names:
        p = tmp_path / (n + ".jpg")
        p.write_bytes(cases.small_cases()[n])
        ins.append(str(p))
        outs.append(str(tmp_path / (n + ".bmp")))
    bad = tmp_path / "bad.jpg"
    bad.write_bytes(b"this is not a jpeg")
    ins += [str(bad), str(tmp_path / "missing.jpg")]
    outs += [str(tmp_path / "bad.bmp"), str(tmp_path / "missing.bmp")]
    ok = hjd.ConvertJpgFiles(ins, outs, threads=3)
    assert ok == [1, 1, 1, 1, 1, 0, 0]
    for n, o in zip(names, outs):
        want = port.bmp24_bytes(port.decode(cases.small_cases()[n])["rgb"])
        assert open(o, "rb").read() == want, n
    assert not os.path.exists(outs[5]) and not os.path.exists(outs[6])


@pytest.mark.timeout(600)
def test_fuzzed_files_never_hang_or_crash(hjd):
    """Random byte corruption anywhere in valid files (headers, tables, entropy data, markers):
    every call returns, every loop is bounded, good neighbours in the batch are untouched."""
    rng = np.random.default_rng(99)
    base = cases.small_cases()
    names = ["420_100x70_ri2", "444_64x48_q85", "gray_33x9_ri4", "420_opt_96x96", "422_100x70_ri3", "444_opt_ri5",
             "440_61x35_ri3", "420_noise_q100"]
    good = base["420_64x48_q85"]
    with hjd.BatchDecoder(0) as d:
        d.upload([good])
        d.decode()
        want = d.rgb(0).copy()
        for rounds in range(12):
            files = [good]
            for n in names:
                b = bytearray(base[n])
                k = int(rng.integers(1, 12))
                for _ in range(k):
                    pos = int(rng.integers(2, len(b)))
                    mode = int(rng.integers(0, 4))
                    if mode == 0:
                        b[pos] = int(rng.integers(0, 256))
                    elif mode == 1:
                        b[pos] = 0xFF
                    elif mode == 2 and pos + 1 < len(b):
                        b[pos], b[pos + 1] = 0xFF, int(rng.integers(0xD0, 0xDA))
                    else:
                        del b[pos:pos + int(rng.integers(1, 40))]
                files.append(bytes(b))
            files.append(good)
            d.upload(files)
            d.decode()
            st = d.status()
            assert st[0] == 0 and st[-1] == 0
            assert np.array_equal(d.rgb(0), want) and np.array_equal(d.rgb(len(files) - 1), want)
