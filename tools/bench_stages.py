import sys, json, time
sys.path.insert(0, '/root/repo')
import bench, hls_jpeg_decoder_b200 as hjd
files = bench.load_images(1024, 0, 1)
arena = hjd.PinnedArena(files)
import os
only = os.environ.get('HJD_STAGES_ONLY')
for name, flags, ov in (("fused-serial", hjd.FLAG_FUSED, 0), ("unfused-serial", 0, 0), ("unfused-overlap", 0, 1), ("fused-overlap", hjd.FLAG_FUSED, 1)):
    if only and name != only:
        continue
    dec = hjd.BatchDecoder(0, flags)
    dec.set_overlap(ov)
    dec.upload_arena(arena); dec.sync()
    for _ in range(3): dec.decode()
    dec.sync()
    acc = {}
    for _ in range(5):
        dec.decode(); t = dec.timings()
        for k, v in t.items(): acc[k] = acc.get(k, 0) + v / 5
    print(name, {k: round(v, 3) for k, v in acc.items()}, "MP/s", round(dec.pixels / 1e6 / (acc["total_ms"] / 1e3)))
    dec.close()
