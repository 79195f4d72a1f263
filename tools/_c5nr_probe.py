import sys, time
sys.path.insert(0, '.')
from concurrent.futures import ProcessPoolExecutor
import hls_jpeg_decoder_b200 as hjd
from tools import gen_jpegs
def job(i): return gen_jpegs.make_c5(i, restart=False)
with ProcessPoolExecutor(16) as ex: files = list(ex.map(job, range(2048), chunksize=32))
files = files * 4
arena = hjd.PinnedArena(files)
dec = hjd.BatchDecoder(0)
dec.set_overlap(0)
dec.upload_arena(arena); dec.sync()
for _ in range(3): dec.decode()
dec.sync()
acc = {}
for _ in range(5):
    dec.decode(); t = dec.timings()
    for k, v in t.items(): acc[k] = acc.get(k, 0) + v / 5
print({k: round(v, 3) for k, v in acc.items()}, 'MP/s', round(dec.pixels / 1e6 / (acc['total_ms'] / 1e3)), 'rounds', dec.selfsync_rounds, 'scan bytes/img', dec.scan_bytes // len(files))
