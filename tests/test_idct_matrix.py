"""CPU tests of the premises the tensor-core IDCT kernel (csrc/mcu_tc.cuh) rests on.

The kernel computes, per 8x8 block, D = V * [M_hi | M_lo]^T on the tensor cores (FP16 operands that hold integers, FP32
accumulation that is therefore exact) and h = D_hi + 2^-24 * D_lo.  Its claim: h is within 18 units of the reference's float
evaluation of 0.25 * sum (loadjpg.cpp:105-124), one unit being 2^-24 * A, A = sum |C(u)C(v) * v|; samples within 20 units of a
change of the truncated, clamped byte are re-evaluated in the reference's own order.  Here the tensor-core tier is
emulated in exact integer arithmetic (which is what the hardware computes: tools/exp_umma_idct.cu checks that on the GPU)
and compared with the reference's order of float operations in numpy float32.
"""
import numpy as np
import pytest

import hls_jpeg_decoder_b200 as hjd

ZZ = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
      35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63]


@pytest.fixture(scope="module")
def tables():
    cos, cc = hjd.idct_tables()          # cos[p][k], cc[u][v]; float32, the reference's libm values
    hi, lo = hjd.idct_matrix()
    return cos, cc, hi, lo


def test_matrix_tile_is_the_fixed_point_split_of_the_reference_constants(tables):
    cos, cc, hi, lo = tables
    assert np.array_equal(hi, np.rint(hi)) and np.array_equal(lo, np.rint(lo)), "both halves are integers"
    assert np.abs(hi).max() <= 2048 and np.abs(lo).max() <= 1024, "exact in FP16"
    worst = 0.0
    for k in range(64):
        nat = ZZ[k]
        u, v = nat & 7, nat >> 3                        # block[8v + u]: u pairs with x, v with y (loadjpg.cpp:112-121)
        for y in range(8):
            for x in range(8):
                c = float(cc[0][0]) if nat == 0 else (float(cc[0][1]) if (u == 0 or v == 0) else 1.0)
                m = 0.25 * c * float(cos[x][u]) * float(cos[y][v])
                worst = max(worst, abs(m - (hi[k, 8 * y + x] * 2048.0 + lo[k, 8 * y + x]) * 2.0 ** -24))
    assert worst <= 2.0 ** -25 * (1 + 1e-9), worst


def _reference_float_sums(v_nat, cos, cc):
    """0.25 * sum in the reference's order of float32 operations, for every (x, y) of every block.  v_nat: [n][64] int."""
    n = v_nat.shape[0]
    ccn = np.ones(64, dtype=np.float32)
    for nat in range(64):
        u, v = nat & 7, nat >> 3
        ccn[nat] = cc[0][0] if nat == 0 else (cc[0][1] if (u == 0 or v == 0) else np.float32(1.0))
    bp = (ccn[None, :] * v_nat.astype(np.float32)).astype(np.float32)         # (C(u)*C(v)) * block[u][v]
    out = np.zeros((n, 64), dtype=np.float32)
    for y in range(8):
        for x in range(8):
            s = np.zeros(n, dtype=np.float32)
            for u in range(8):
                for v in range(8):
                    t = (bp[:, 8 * v + u] * cos[x][u]).astype(np.float32)
                    t = (t * cos[y][v]).astype(np.float32)
                    s = (s + t).astype(np.float32)
            out[:, 8 * y + x] = np.float32(0.25) * s
    return out, np.abs(bp).sum(axis=1, dtype=np.float64)


def _blocks(rng, n):
    """De-quantised coefficient blocks in natural order: photographic, DC + few, dense near the kernel's limits."""
    v = np.zeros((n, 64), dtype=np.int64)
    for b in range(n):
        kind = b % 4
        if kind < 2:
            v[b, 0] = rng.integers(-1016, 1017)
            for k in range(1, 64):
                lim = 60 if k < 6 else 30 if k < 15 else 10 if k < 28 else 2
                if rng.integers(0, 100) < lim:
                    v[b, ZZ[k]] = rng.integers(1, 300 if k < 6 else 80 if k < 15 else 24) * rng.choice([-1, 1])
        elif kind == 2:
            v[b, 0] = 8 * rng.integers(-127, 128)
            for _ in range(2):
                v[b, ZZ[rng.integers(1, 11)]] = rng.integers(-8, 9)
        else:
            budget = 7900                                # sum |v| < 8192 <= A < 4000 * 2 is what the kernel admits
            for k in range(64):
                mag = int(rng.integers(0, 2048 if k < 2 else 257))
                mag = min(mag, budget)
                budget -= mag
                v[b, ZZ[k]] = mag * rng.choice([-1, 1])
    return v


def test_tensor_core_tier_stays_inside_its_error_budget(tables):
    cos, cc, hi, lo = tables
    rng = np.random.default_rng(20261018)
    v = _blocks(rng, 1200)
    ref, A = _reference_float_sums(v, cos, cc)
    vz = v[:, ZZ]                                            # zig-zag order, as the kernel holds them
    d_hi = vz.astype(np.float64) @ hi                        # exact: integers below 2^53
    d_lo = vz.astype(np.float64) @ lo
    assert np.abs(d_hi).max() < 2 ** 24 and np.abs(d_lo).max() < 2 ** 24, "every partial sum is an integer FP32 holds exactly"
    h = (d_lo.astype(np.float32) * np.float32(2.0 ** -24) + (d_hi * 2.0 ** -13).astype(np.float32)).astype(np.float32)   # the FMA
    unit = (A * 2.0 ** -24)[:, None]
    err = np.abs(h.astype(np.float64) - ref.astype(np.float64)) / np.maximum(unit, 1e-300)
    nz = (A > 0)
    assert err[nz].max() <= 18.0, f"error budget of mcu_tc.cuh exceeded: {err[nz].max()} units"
    # the truncation test: a sample is accepted when h - win and h + win give the same clamped byte; then that byte is the reference's
    win = (20.0 * unit * 1.0).astype(np.float32)
    def byte(t):
        return np.clip(np.trunc(t.astype(np.float64)) + 128, 0, 255)
    accepted = byte(h - win) == byte(h + win)
    assert (byte(h)[accepted] == byte(ref)[accepted]).all()
    flagged = 1.0 - accepted[nz].mean()
    assert flagged < 0.05, flagged
