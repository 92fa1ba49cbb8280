"""world_size-2 gloo test of the sharded-search exchange step: each rank owns a contiguous
row block, emits local top-k keys, ONE all-gather, merge -> identical to the global top-k.
(The CUDA local search is replaced by the oracle here; the exchange + merge logic is the
same code path shape as ShardedIndex.search / merge_keys.)"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    from oracle import oracle, synth
    from mdir_b200.search import ShardedIndex, make_keys_host, keys_to_host, merge_keys_host
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, d, nq, k = 777, 32, 6, 25
        db = synth.descriptors(n, d, 31, clusters=12)
        q, _ = synth.planted_queries(db, nq, 32)
        lo, hi = ShardedIndex.shard_bounds(n, world, rank)
        sc_local = np.round(oracle.scores(db[lo:hi].T, q.T) * 40).astype(np.float32) / 40
        idx, val = oracle.topk_from_scores(sc_local, k)
        keys = make_keys_host(val.T, idx.T + lo)                                   # (nq, k) uint64
        local = torch.from_numpy(keys.view(np.int64).copy())
        gathered = torch.empty((world * nq, k), dtype=torch.int64)
        dist.all_gather_into_tensor(gathered, local)
        merged = merge_keys_host(gathered.view(world, nq, k).numpy().view(np.uint64), k)
        msc, midx = keys_to_host(merged)
        sc_all = np.round(oracle.scores(db.T, q.T) * 40).astype(np.float32) / 40
        gidx, gval = oracle.topk_from_scores(sc_all, k)
        ok = bool(np.array_equal(midx, gidx.T) and np.array_equal(msc, gval.T))
        flag = torch.tensor([1 if ok else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            ret.put(int(flag.item()))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_merge():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert ret.get(timeout=10) == 1
