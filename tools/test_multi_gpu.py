#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun on a box with >= 2 GPUs):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/test_multi_gpu.py

Every rank holds a row shard; the sharded search / alpha-QE / DBA results must equal the
single-GPU results on the whole database bit for bit (indices) -- SURVEY.md section 4,
"distributed" row.  Rank 0 prints one PASS/FAIL line per check and exits non-zero on failure."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import synth  # noqa: E402
import mdir_b200  # noqa: E402
from mdir_b200 import qe  # noqa: E402
from mdir_b200.search import GraphedSearch, ShardedIndex  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    ok_all = True

    def check(name, cond):
        nonlocal ok_all
        flag = torch.tensor([1 if cond else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        good = bool(flag.item())
        ok_all &= good
        if rank == 0:
            print("[%s] %s" % ("PASS" if good else "FAIL", name), flush=True)

    n, d, nq, k = 60013, 128, 70, 100
    db = synth.descriptors(n, d, 81, clusters=150)
    q, src = synth.planted_queries(db, nq, 82)
    lo, hi = ShardedIndex.shard_bounds(n, world, rank)
    sharded = ShardedIndex(db[lo:hi], idx_base=lo, device=dev)
    check("peer-memory mailboxes mapped on every rank", sharded._mb is not None)
    single = mdir_b200.Index(db, device=dev)
    for prec in ("bf16", "fp32"):
        s1, i1 = single.search(q, k, precision=prec)
        s2, i2 = sharded.search(q, k, precision=prec)
        check("sharded search == single (%s, fused NVLink exchange + merge)" % prec, torch.equal(i1, i2) and torch.equal(s1, s2))
    for rep in range(5):                                   # sequence numbers / parity buffers across back-to-back steps
        qq, _ = synth.planted_queries(db, nq, 90 + rep)
        s1, i1 = single.search(qq, k)
        s2, i2 = sharded.search(qq, k)
        check("  repeat %d" % rep, torch.equal(i1, i2) and torch.equal(s1, s2))
    big_q, _ = synth.planted_queries(db, 300, 99)          # more than one exchange block of 128 queries
    s1, i1 = single.search(big_q, 10)
    s2, i2 = sharded.search(big_q, 10)
    check("300 queries (3 exchange blocks)", torch.equal(i1, i2) and torch.equal(s1, s2))
    check("exchange status clean", sharded.exchange_status() == 0)
    ShardedIndex.p2p = False                               # the NCCL all-gather route stays available
    nccl = ShardedIndex.from_local(sharded.local)
    ShardedIndex.p2p = True
    s1, i1 = single.search(q, k)
    s2, i2 = nccl.search(q, k)
    check("sharded search == single (ncclAllGather + merge kernel)", nccl._mb is None and torch.equal(i1, i2) and torch.equal(s1, s2))
    gs = GraphedSearch(sharded, nq, k)
    s3, i3 = gs(torch.from_numpy(q).pin_memory())
    torch.cuda.synchronize()
    s1, i1 = single.search(q, k)
    check("CUDA-graph sharded search == single", torch.equal(i1, i3) and torch.equal(s1, s3) and not gs.check_overflow())
    check("planted neighbour first", bool(np.array_equal(i3.cpu().numpy()[:, 0], src)))
    # deferred exchange: replay t returns the merged result of replay t-1, drain() the last one
    gd = GraphedSearch(sharded, nq, k, deferred=True)
    batches = [torch.from_numpy(synth.planted_queries(db, nq, 120 + b)[0]).to(dev) for b in range(5)]
    refs = [single.search(b, k) for b in batches]
    good = True
    for b in range(5):
        s_prev, i_prev = gd(batches[b])
        torch.cuda.synchronize()
        if b >= 1:
            good &= torch.equal(i_prev, refs[b - 1][1]) and torch.equal(s_prev, refs[b - 1][0])
    s_last, i_last = gd.drain()
    torch.cuda.synchronize()
    good &= torch.equal(i_last, refs[4][1]) and torch.equal(s_last, refs[4][0])
    check("deferred-exchange graph: step t returns step t-1, drain returns the last", good)
    # overlapped exchange: the graph holds the local step, the push + merge of the same ticket runs on a second stream
    go = GraphedSearch(sharded, nq, k, overlap=True)
    good = go.overlap
    for b in range(5):
        s_o, i_o = go(batches[b])
        torch.cuda.synchronize()
        good &= torch.equal(i_o, refs[b][1]) and torch.equal(s_o, refs[b][0]) and not go.check_overflow()
    check("overlap-mode graph (local step graph + exchange kernel) == single", good)
    # split graphs need the one-launch scan route on every shard: a database of 40,000 rows per rank
    n_big = 40000 * world
    db_big = synth.descriptors(n_big, 64, 181, clusters=200)
    lo_b, hi_b = ShardedIndex.shard_bounds(n_big, world, rank)
    sh_big = ShardedIndex(db_big[lo_b:hi_b], idx_base=lo_b, device=dev)
    single_big = mdir_b200.Index(db_big, device=dev)
    gsplit = GraphedSearch(sh_big, nq, k, overlap=True, split=True)
    good = gsplit.split
    for b in range(5):
        qb = torch.from_numpy(synth.planted_queries(db_big, nq, 220 + b)[0]).to(dev)
        s_ref, i_ref = single_big.search(qb, k)
        s_o, i_o = gsplit(qb)
        torch.cuda.synchronize()
        good &= torch.equal(i_o, i_ref) and torch.equal(s_o, s_ref) and not gsplit.check_overflow()
    check("split graphs (scan | finalize + exchange, private candidate workspace; split=%s) == single" % gsplit.split, good)
    pipe_s = mdir_b200.SearchPipeline(sh_big, nq, k)
    hq = [torch.from_numpy(synth.planted_queries(db_big, nq, 220 + b)[0]).pin_memory() for b in range(7)]
    outs_s = [(s_.copy(), i_.copy()) for s_, i_ in pipe_s.map(hq)]
    good = all(g_.split for g_ in pipe_s.graphs)
    for b in range(7):
        s_ref, i_ref = single_big.search(hq[b], k)
        good &= bool(np.array_equal(outs_s[b][1], i_ref.cpu().numpy()) and np.array_equal(outs_s[b][0], s_ref.cpu().numpy()))
    check("SearchPipeline with split graphs == single, in order", good)
    sh_big.close()
    del single_big, gsplit, pipe_s
    hb = [b.cpu().pin_memory() for b in batches]
    pipe_o = mdir_b200.SearchPipeline(sharded, nq, k)
    outs = [(s.copy(), i.copy()) for s, i in pipe_o.map(hb * 3)]
    good = pipe_o.overlap and not pipe_o.deferred and len(outs) == 15
    for b in range(15):
        good &= bool(np.array_equal(outs[b][1], refs[b % 5][1].cpu().numpy()) and np.array_equal(outs[b][0], refs[b % 5][0].cpu().numpy()))
    t_a = pipe_o.submit(hb[0])
    sa, ia = pipe_o.result(t_a)
    good &= bool(np.array_equal(ia, refs[0][1].cpu().numpy()))
    check("SearchPipeline over the sharded index (exchange overlapped with the next scan) == single, in order", good)
    pipe = mdir_b200.SearchPipeline(sharded, nq, k, deferred=True)
    outs = [(s.copy(), i.copy()) for s, i in pipe.map(hb)]
    good = pipe.deferred and not pipe.overlap and len(outs) == 5
    for b in range(5):
        good &= bool(np.array_equal(outs[b][1], refs[b][1].cpu().numpy()) and np.array_equal(outs[b][0], refs[b][0].cpu().numpy()))
    t_a = pipe.submit(hb[0])
    sa, ia = pipe.result(t_a)                              # latest ticket collected right away -> drain path
    good &= bool(np.array_equal(ia, refs[0][1].cpu().numpy()))
    outs2 = [(s.copy(), i.copy()) for s, i in pipe.map(hb[1:3])]
    good &= bool(np.array_equal(outs2[0][1], refs[1][1].cpu().numpy()) and np.array_equal(outs2[1][1], refs[2][1].cpu().numpy()))
    check("SearchPipeline over the sharded index (deferred exchange) == single, in order", good)
    s2, i2 = sharded.search(q, k)
    s1, i1 = single.search(q, k)
    check("sync search after deferred traffic still exact", torch.equal(i1, i2) and torch.equal(s1, s2) and sharded.exchange_status() == 0)
    # ---- a flagged step on ONE rank must be visible on EVERY rank: the last rank's shard is adversarial
    # (scores grow with the row index, so its candidate segments overflow inside the graph, where nothing recovers)
    n_sh, d2 = 40000, 64
    rs = np.random.RandomState(5)
    base = rs.randn(d2).astype(np.float32)
    base /= np.linalg.norm(base)
    parts = [synth.descriptors(n_sh, d2, 300 + r) for r in range(world - 1)]
    parts.append((np.linspace(0.1, 1.0, n_sh, dtype=np.float32)[:, None] * base[None, :] + rs.randn(n_sh, d2).astype(np.float32) * 0.01).astype(np.float32))
    db_adv = np.concatenate(parts)
    q_adv = np.stack([base, -base, base + 0.1 * rs.randn(d2).astype(np.float32)]).astype(np.float32)
    q_easy = synth.descriptors(3, d2, 77)
    lo3, hi3 = rank * n_sh, (rank + 1) * n_sh
    sh3 = ShardedIndex(db_adv[lo3:hi3], idx_base=lo3, device=dev)
    single3 = mdir_b200.Index(db_adv, device=dev)
    ref_adv, ref_easy = single3.search(q_adv, k), single3.search(q_easy, k)
    g3 = GraphedSearch(sharded if False else sh3, 3, k)
    g3(torch.from_numpy(q_adv).to(dev))
    torch.cuda.synchronize()
    local_flag = bool(g3.ovf.any().item())
    check("overflow inside the graph on the last rank only (%s here)" % local_flag, local_flag == (rank == world - 1) or local_flag)
    check("... raises the merged step's status word on EVERY rank", g3.check_overflow())
    s_r, i_r = sh3.search_collective_recovery(q_adv, k)
    check("collective recovery == single", torch.equal(i_r, ref_adv[1]) and torch.equal(s_r, ref_adv[0]))
    pipe3 = mdir_b200.SearchPipeline(sh3, 3, k)
    hb3 = [torch.from_numpy(x).pin_memory() for x in (q_adv, q_easy, q_adv)]
    outs3 = [(s_.copy(), i_.copy()) for s_, i_ in pipe3.map(hb3)]
    good = pipe3.n_recovered >= 2
    for (s_, i_), ref in zip(outs3, (ref_adv, ref_easy, ref_adv)):
        good &= bool(np.array_equal(i_, ref[1].cpu().numpy()) and np.array_equal(s_, ref[0].cpu().numpy()))
    check("SearchPipeline redoes flagged steps collectively (%d redone) and stays exact" % pipe3.n_recovered, good)
    sh3.close()
    q1 = qe.expand_queries(single, q, 3.0, 10)
    q2 = qe.expand_queries(sharded, q, 3.0, 10)
    check("sharded alpha-QE expansion == single (1e-6)", bool(torch.allclose(q1, q2, rtol=0, atol=1e-6)))
    small_n = 3000
    dbs = synth.descriptors(small_n, 64, 83, clusters=20)
    lo2, hi2 = ShardedIndex.shard_bounds(small_n, world, rank)
    sh2 = ShardedIndex(dbs[lo2:hi2], idx_base=lo2, device=dev)
    aug_sh = qe.dba_sharded(sh2, 3.0, 5)
    aug_single = qe.dba(mdir_b200.Index(dbs, device=dev), 3.0, 5)
    check("sharded DBA rows == single (1e-6)", bool(torch.allclose(aug_sh.local.db32, aug_single.db32[lo2:hi2], rtol=0, atol=1e-6)))
    torch.cuda.synchronize()
    dist.barrier()
    sh2.close()
    sharded.close()
    check("mailboxes closed", sharded._mb is None)
    if rank == 0:
        print("MULTI-GPU %s" % ("PASSED" if ok_all else "FAILED"), flush=True)
    sys.stdout.flush()
    os._exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
