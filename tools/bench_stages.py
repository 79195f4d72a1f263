"""Per-stage CUDA-event timings of the decode variants on the config-2 batch (B200 only).

    python tools/bench_stages.py            # every variant
    HJD_STAGES_ONLY=default python tools/bench_stages.py
    HJD_LIB_PATH=/path/to/tuning_build.so python tools/bench_stages.py   # kernels built with other -D knobs

Variants: default (kernels 2+3 fused per MCU), planes (HJD_FLAG_KEEP_PLANES: unfused kernels 2 and 3),
strip (HJD_FLAG_FUSED: strip-fused kernel), and the chunked 3-stream execution of the first two.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import hls_jpeg_decoder_b200 as hjd  # noqa: E402

VARIANTS = (("default", 0, 0), ("planes", hjd.FLAG_KEEP_PLANES, 0), 
            ("default-chunked", 0, 1500000), ("planes-chunked", hjd.FLAG_KEEP_PLANES, 1500000))


def main():
    files = bench.load_images(1024, 0, 1)
    arena = hjd.PinnedArena(files)
    only = os.environ.get("HJD_STAGES_ONLY")
    for name, flags, overlap in VARIANTS:
        if only and name != only:
            continue
        dec = hjd.BatchDecoder(0, flags)
        dec.set_overlap(overlap)
        dec.upload_arena(arena)
        dec.sync()
        for _ in range(3):
            dec.decode()
        dec.sync()
        acc = {}
        for _ in range(5):
            dec.decode()
            t = dec.timings()
            for k, v in t.items():
                acc[k] = acc.get(k, 0) + v / 5
        print(name, {k: round(v, 3) for k, v in acc.items()}, "MP/s", round(dec.pixels / 1e6 / (acc["total_ms"] / 1e3)))
        dec.close()


if __name__ == "__main__":
    main()
