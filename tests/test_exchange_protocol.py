"""Model check of the mailbox protocol of csrc/shard_merge.cu (host-only, no GPU): W ranks run the kernel's
micro-operations (push data to every peer, publish flags, wait for the world's flags of the step being merged, read that
step's data) under a random interleaving.  Safety = every read sees the data of exactly the step it merges and nobody
deadlocks.  Four buffers are safe in the synchronous and in the deferred mode; the model also shows why fewer are not
enough for the deferred mode (the check is able to fail)."""
import random

import pytest


def simulate(world, n_steps, n_bufs, deferred, seed):
    rnd = random.Random(seed)
    data = [[[0] * world for _ in range(n_bufs)] for _ in range(world)]      # data[owner][buf][src] = seq tag
    flag = [[[0] * world for _ in range(n_bufs)] for _ in range(world)]
    pc = [(1, 0) for _ in range(world)]                                       # (seq, micro-op index) per rank

    def ops_of(seq):
        mseq = seq - 1 if deferred else seq
        ops = [("push", p) for p in range(world)] + [("flag", p) for p in range(world)]
        if mseq >= 1:
            ops += [("wait", None)] + [("read", s) for s in range(world)]
        return ops, mseq

    while True:
        runnable = []
        for r in range(world):
            seq, i = pc[r]
            if seq > n_steps:
                continue
            ops, mseq = ops_of(seq)
            kind, arg = ops[i]
            if kind == "wait" and not all(flag[r][mseq % n_bufs][s] == mseq for s in range(world)):
                continue
            runnable.append(r)
        if not runnable:
            return "ok" if all(seq > n_steps for seq, _ in pc) else "deadlock"
        r = rnd.choice(runnable)
        seq, i = pc[r]
        ops, mseq = ops_of(seq)
        kind, arg = ops[i]
        if kind == "push":
            data[arg][seq % n_bufs][r] = seq
        elif kind == "flag":
            flag[arg][seq % n_bufs][r] = seq
        elif kind == "read" and data[r][mseq % n_bufs][arg] != mseq:
            return "stale read: rank %d step %d merges %d but source %d's slot holds %d" % (r, seq, mseq, arg, data[r][mseq % n_bufs][arg])
        i += 1
        pc[r] = (seq + 1, 0) if i == len(ops) else (seq, i)


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("deferred", [False, True])
def test_four_buffers_are_safe(world, deferred):
    for seed in range(300):
        assert simulate(world, 12, 4, deferred, seed) == "ok", (world, deferred, seed)


def test_the_model_can_fail_with_too_few_buffers():
    # deferred merging with two buffers: a fast peer overwrites (or re-flags) a slot before its owner merged it
    outcomes = {simulate(3, 12, 2, True, seed) for seed in range(300)}
    assert any(o != "ok" for o in outcomes)
