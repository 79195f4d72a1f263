"""CPU tests of the multi-GPU host logic: shard ranges, and a world_size-2 gloo run in which two
ranks shard a batch, "decode" their ranges with the oracle (the CUDA library needs a GPU) and
agree on a checksum of checksums."""
import hashlib
import os
import subprocess
import sys
import textwrap

import numpy as np

from hls_jpeg_decoder_b200.sharding import shard_range


def test_shard_ranges_cover_and_balance():
    rng = np.random.default_rng(0)
    for n in (0, 1, 2, 7, 64, 1000):
        sizes = rng.integers(1000, 500000, size=n).tolist()
        for world in (1, 2, 3, 4, 8):
            rs = [shard_range(sizes, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            for a, b in zip(rs, rs[1:]):
                assert a[1] == b[0] and a[0] <= a[1]
            if n >= 8 * world:
                loads = [sum(sizes[a:b]) for a, b in rs]
                assert max(loads) <= 1.25 * (sum(sizes) / world) + 500000


def test_two_rank_gloo_sharded_decode(tmp_path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "rank.py"
    script.write_text(textwrap.dedent(f"""
        import hashlib, os, sys
        sys.path.insert(0, {root!r})
        import torch, torch.distributed as dist
        from hls_jpeg_decoder_b200.sharding import shard_range
        from oracle import port
        from tests import cases
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        files = list(cases.small_cases().values())
        lo, hi = shard_range([len(f) for f in files], rank, world)
        digests = [hashlib.sha256(port.decode(f)["rgb"].tobytes()).digest() for f in files[lo:hi]]
        gathered = [None] * world
        dist.all_gather_object(gathered, (lo, hi, digests))
        if rank == 0:
            gathered.sort()
            assert gathered[0][0] == 0 and gathered[-1][1] == len(files)
            allsum = hashlib.sha256(b"".join(d for _, _, ds in gathered for d in ds)).hexdigest()
            want = hashlib.sha256(b"".join(hashlib.sha256(port.decode(f)["rgb"].tobytes()).digest() for f in files)).hexdigest()
            assert allsum == want
            print("OK", allsum)
        dist.barrier()
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "OK" in out.stdout


def test_library_shard_rule_equals_python_rule():
    """hjd_shard_range (what hjd_multi_decode_host cuts a batch with) == sharding.shard_range (what bench.py's
    ranks use): same ranges for every rank, including empty batches and more ranks than images."""
    import hls_jpeg_decoder_b200 as hjd
    rng = np.random.default_rng(3)
    for n in (0, 1, 2, 5, 33, 1024):
        sizes = rng.integers(1, 900000, size=n).tolist()
        for world in (1, 2, 3, 4, 8, 16):
            for r in range(world):
                assert hjd.shard_range_c(sizes, r, world) == shard_range(sizes, r, world), (n, world, r)
