"""World-size-2 `gloo` test of the multi-GPU plumbing on CPU (the data path has no collective: each
rank scores its own contiguous stream range; torch.distributed only merges results and timing). The
per-shard scorer here is the oracle, standing in for one GPU's rp_batch."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rustpotter_b200.sharding import gather_detections, reduce_step_stats, shard_range  # noqa: E402


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 4096, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi)) if n < 100 else [(lo, hi)]
            if n < 100:
                assert seen == list(range(n))
            else:
                assert seen[0][0] == 0 and seen[-1][1] == n and all(a[1] == b[0] for a, b in zip(seen, seen[1:]))
                sizes = [hi - lo for lo, hi in seen]
                assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from tests.helpers import make_wakeword, splice, synth_audio
    n_streams = 7
    rpw, utts = make_wakeword(O, d=16, lengths=(40, 44, 48), seed=5)
    audio = synth_audio(n_streams, 60 * 480, seed=9)          # identical on every rank (seeded)
    for b in (0, 3, 6):
        splice(audio[b], utts[b % 3], 30 + b)
    lo, hi = shard_range(n_streams, rank, world)
    windows, counts, dets = O.run_streams(O.default_config(min_scores=2), [rpw], audio[lo:hi], n_threads=1, max_det=4)
    local = [(s, 0, d) for s in range(hi - lo) for d in dets[s]]
    merged = gather_detections(dist, local, lo)
    t, u = reduce_step_stats(dist, torch.device("cpu"), 10.0 + rank, float(windows))
    if rank == 0:
        w1, c1, d1 = O.run_streams(O.default_config(min_scores=2), [rpw], audio, n_threads=1, max_det=4)
        want = [(s, float(d["score"])) for s in range(n_streams) for d in d1[s]]
        got = [(s, float(d["score"])) for s, _, d in merged]
        q.put((got == want and len(want) >= 2, t, u, w1))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shards_merge_to_single_process_result():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok, t, u, w1 = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
    assert t == 11.0          # max over ranks
    assert u == float(w1)     # windows summed over shards == single-process count
