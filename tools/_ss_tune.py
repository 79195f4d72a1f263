"""Builder tuning aid: entropy-stage time of restart-free batches against the range length of the
synchronisation kernel (hjd_batch_set_selfsync_range)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import hls_jpeg_decoder_b200 as hjd

for cfg, n in (("c2nr", 1024), ("c2nr", 256), ("c4", 1)):
    files = bench.load_images(n, 0, 1, cfg)
    arena = hjd.PinnedArena(files)
    dec = hjd.BatchDecoder(0)
    for rng in (0, 32, 64, 128, 256):
        dec.set_selfsync_range(rng)
        dec.upload_arena(arena)
        dec.sync()
        for _ in range(3):
            dec.decode()
        dec.sync()
        dec.mark(0)
        for _ in range(5):
            dec.decode()
        dec.mark(1)
        dec.sync()
        t = dec.timings()
        print(cfg, n, "range", rng, "ms/step %.3f entropy %.3f rounds %d" % (dec.elapsed_ms(0, 1) / 5, t["entropy_ms"], dec.selfsync_rounds), flush=True)
    dec.close()
    arena.close()
