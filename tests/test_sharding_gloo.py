"""world_size-2 `gloo` test of the multi-process path bench.py uses for N > 1 (one process per GPU, no collective on the
data path): each rank takes the contiguous shard the planner assigns, processes it independently, and only the timing
reduction (max over ranks) and a gather for checking cross the process boundary.  On CPU the per-shard work is done by
the oracle (test infrastructure standing in for the device); the assembled result must equal the unsharded answer."""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist

    import coracle
    import wgpu_sigops_b200 as w

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = w.load()
    bounds = (ctypes.c_size_t * (world + 1))()
    used = ctypes.c_int()
    assert lib.sigops_plan_shards(n, world, bounds, ctypes.byref(used)) == 0 and used.value == world
    lo, hi = bounds[rank], bounds[rank + 1]
    sigs, msgs, pks = coracle.gen_ecdsa(0, n, seed=5, threads=2)  # every rank can regenerate the whole batch
    out, st = coracle.ecrecover(0, sigs[lo:hi], msgs[lo:hi], threads=2)  # this rank's shard only
    # max-over-ranks of a per-rank "time" (bench.py's reduction)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # gather the shards on rank 0 for checking (not part of the data path)
    parts = [None] * world
    dist.all_gather_object(parts, (int(lo), int(hi), out.tobytes(), st.tobytes()))
    if rank == 0:
        full = np.zeros((n, 64), dtype=np.uint8)
        cover = np.zeros(n, dtype=np.int32)
        for plo, phi, ob, sb in parts:
            full[plo:phi] = np.frombuffer(ob, dtype=np.uint8).reshape(-1, 64)
            cover[plo:phi] += 1
            assert not np.frombuffer(sb, dtype=np.uint8).any()
        q.put((bool((cover == 1).all()), bool((full == pks).all()), float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_batch_matches_unsharded():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n, world, port = 8192 + 3, 2, _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    covered, equal, tmax = q.get(timeout=10)
    assert covered and equal and tmax == 2.0
