"""Raw host-link probe: N processes, one per GPU, each copying a config-2 step's bytes
(0.37 GB up, 6.37 GB down) between pinned host memory and HBM concurrently, with NO kernels -- the ceiling
of the end-to-end (`e2e`) number at N GPUs on this box.

    python tools/d2h_probe.py --gpus 8 [--no-numa] [--reps 3]

Prints one JSON line: aggregate and per-rank GB/s.  Uses hjd_link_probe / hjd_host_alloc_near of libhjd.so.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

H2D = 371274352
D2H = 6370099200


def worker(rank, n, numa, reps, start, q):
    import hls_jpeg_decoder_b200 as hjd
    L = hjd.lib()
    a = L.hjd_host_alloc_near(rank, H2D) if numa else L.hjd_host_alloc(H2D)
    b = L.hjd_host_alloc_near(rank, D2H) if numa else L.hjd_host_alloc(D2H)
    hjd.link_probe(rank, a, H2D, b, D2H, 1)            # context + first touch
    start.wait()
    up, down = hjd.link_probe(rank, a, H2D, b, D2H, reps)
    q.put((rank, up, down, L.hjd_device_numa_node(rank)))
    L.hjd_host_free(a)
    L.hjd_host_free(b)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--no-numa", action="store_true")
    a = ap.parse_args()
    ctx = mp.get_context("spawn")
    q, start = ctx.Queue(), ctx.Barrier(a.gpus)
    ps = [ctx.Process(target=worker, args=(r, a.gpus, not a.no_numa, a.reps, start, q)) for r in range(a.gpus)]
    for p in ps:
        p.start()
    res = sorted(q.get() for _ in ps)
    for p in ps:
        p.join()
    worst = max(max(u, d) for _, u, d, _ in res) / 1e3
    print(json.dumps({"gpus": a.gpus, "numa_near_buffers": not a.no_numa,
                      "aggregate_d2h_GB_per_s": round(a.gpus * D2H / worst / 1e9, 2),
                      "aggregate_h2d_GB_per_s": round(a.gpus * H2D / worst / 1e9, 2),
                      "config2_step_ms_floor": round(worst * 1e3, 2),
                      "e2e_MP_per_s_ceiling": round(a.gpus * 1024 * 1920 * 1080 / 1e6 / worst, 1),
                      "per_rank": [{"rank": r, "h2d_ms": round(u, 2), "d2h_ms": round(d, 2),
                                    "d2h_GB_per_s": round(D2H / d / 1e6, 2), "numa_node": node} for r, u, d, node in res]}))


if __name__ == "__main__":
    main()
