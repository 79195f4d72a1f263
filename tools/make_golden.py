"""Generate tests/golden/: small JPEG fixtures + known answers produced by the REAL reference.

Runs only where /root/reference is mounted (uses oracle/_ref built by oracle/build_ref.sh).
For every case of tests/cases.small_cases() it stores the .jpg and the sha256 of the
reference's coefficients (int16 LE, MCU raster order, zig-zag), planes (Y|Cb|Cr, MCU-padded)
and RGB (top-down, R first).  Restart-free 3-component cases are additionally checked against
the UNMODIFIED JpegDecodeHW path (harness mode 0); restart cases against their restart-free
twin decoded by the unmodified path (SURVEY.md 8c route A) where the encoder allows it.

    python tools/make_golden.py
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refbind  # noqa: E402
from tests.cases import small_cases  # noqa: E402


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main() -> None:
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    meta = {"_generator": "tools/make_golden.py", "_oracle": "oracle/_ref (reference compiled as plain C++)", "cases": {}}
    cases = dict(small_cases())
    lenna = refbind.lenna_path()
    if lenna:
        cases["__lenna__"] = open(lenna, "rb").read()
    for name, jpg in cases.items():
        r = refbind.decode(jpg, mode=1)
        assert r["rc"] == 0, (name, r["rc"])
        w, h, nc, hf, vf, ri = refbind.sniff(jpg)
        entry = dict(width=w, height=h, ncomp=nc, hf=hf, vf=vf, restart_interval=ri, jpeg_bytes=len(jpg),
                     jpeg_sha256=hashlib.sha256(jpg).hexdigest(), n_blocks=int(r["coef"].shape[0]),
                     coef_sha256=sha(r["coef"]), rgb_sha256=sha(r["rgb"]),
                     planes_sha256=sha(np.concatenate([p.ravel() for p in r["planes"]][: (3 if nc == 3 else 1)])))
        if nc == 3 and ri == 0:
            r0 = refbind.decode(jpg, mode=0)
            assert r0["rc"] == 0 and np.array_equal(r0["rgb"], r["rgb"]), name
            entry["checked_against"] = "unmodified JpegDecodeHW"
        else:
            entry["checked_against"] = "harness (MCU-counted restarts / grayscale extension)"
        meta["cases"][name] = entry
        if name != "__lenna__":          # the reference's own fixture stays out of the tracked tree
            with open(os.path.join(out_dir, name + ".jpg"), "wb") as fh:
                fh.write(jpg)
    with open(os.path.join(out_dir, "golden.json"), "w") as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)
    print(f"wrote {len(cases)} cases to {out_dir}")


if __name__ == "__main__":
    main()
