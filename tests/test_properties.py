"""Property tests (hypothesis) of the host-side ordering / merge logic: the 64-bit key order is exactly the
stable descending argsort, and merging per-shard top-k lists is independent of how the rows were sharded."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import oracle
from mdir_b200.search import ShardedIndex, make_keys_host, keys_to_host, merge_keys_host, merge_keys_by_rank_host, default_shortlist

scores_st = st.lists(st.one_of(st.floats(-4, 4, width=32), st.sampled_from([0.0, -0.0, 1.0, 1.0, -1.0, float("inf"), float("-inf")])),
                     min_size=1, max_size=200)


@settings(max_examples=60, deadline=None)
@given(scores_st)
def test_key_order_is_stable_descending_argsort(vals):
    s = np.asarray(vals, dtype=np.float32)
    keys = make_keys_host(s, np.arange(s.shape[0]))
    assert len(set(keys.tolist())) == s.shape[0]                         # keys are unique
    assert np.array_equal(np.argsort(keys, kind="stable"), oracle.ranks_from_scores(s[:, None])[:, 0])
    sc, idx = keys_to_host(keys)
    assert np.array_equal(idx, np.arange(s.shape[0])) and np.array_equal(sc, np.where(s == 0, np.float32(0), s))


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 400), st.integers(1, 30), st.integers(0, 2 ** 31 - 1))
def test_shard_merge_is_partition_independent(n, k, seed):
    rs = np.random.RandomState(seed)
    sc = (np.round(rs.randn(n, 3) * 4) / 4).astype(np.float32)            # heavy ties
    ref_i, ref_v = oracle.topk_from_scores(sc, min(k, n))
    for world in (1, 2, 4, 8):
        parts = []
        for r in range(world):
            lo, hi = ShardedIndex.shard_bounds(n, world, r)
            pad = np.full((3, k), np.uint64(0xffffffffffffffff))
            if hi > lo:
                idx, val = oracle.topk_from_scores(sc[lo:hi], min(k, hi - lo))
                keys = make_keys_host(val.T, idx.T + lo)
                pad[:, :keys.shape[1]] = keys
            parts.append(pad)
        merged = merge_keys_host(np.stack(parts), k)
        # the lower_bound-rank merge of csrc/shard_merge.cu (numpy restatement) gives the same keys as the sort
        assert np.array_equal(merge_keys_by_rank_host(np.stack(parts), k), merged)
        msc, midx = keys_to_host(merged)
        kk = min(k, n)
        assert np.array_equal(midx[:, :kk], ref_i.T) and np.array_equal(msc[:, :kk], ref_v.T)
        assert np.all(midx[:, kk:] == -1)


def test_default_shortlist_monotone():
    prev = 0
    for k in range(1, 2000, 7):
        s = default_shortlist(k)
        assert s >= 1.25 * k and s % 64 == 0 and s >= prev          # whole rounds of the 2 x 32 re-scoring warps
        prev = s


# ---- histogram-sort route of the full ranking (csrc/ranks.cu: hs_plan / hs_scatter / hs_bucket_sort) -------------------
def _hs_rank_restatement(s, cells=4096, t_rows=2048, fine=4096, hi=None, scale=None):
    """numpy restatement (fp32 arithmetic, one operation at a time like the kernels' __fsub_rn / __fmul_rn) of how the
    histogram sort orders one query's scores: cell -> bucket (= floor(first rank of the cell / t_rows)) -> fine bin
    inside the bucket's cell range -> (score key, row) composite.  Returns the permutation it produces."""
    s = np.asarray(s, dtype=np.float32)
    n = s.shape[0]
    keys = make_keys_host(s, np.arange(n))                                # (desc key << 32) | row: the order to reproduce
    fin = s[np.isfinite(s)]
    if hi is None:
        m = np.float32(fin.mean()) if fin.size else np.float32(0)
        sd = np.float32(fin.std()) if fin.size else np.float32(0)
        hi, scale = np.float32(m + np.float32(4) * sd), (np.float32(cells) / (np.float32(8) * sd) if sd > 0 else np.float32(0))
    canon = np.where(s == 0, np.float32(0), s)                            # the key's canonical score: -0 -> +0
    with np.errstate(invalid="ignore", over="ignore"):
        v = ((np.float32(hi) - canon).astype(np.float32) * np.float32(scale)).astype(np.float32)
    nan = np.isnan(v)
    vt = np.where(nan, 0, np.clip(v, -2.0 ** 31, 2.0 ** 31 - 128)).astype(np.int64)      # cvt.rzi saturates
    cell = np.where(nan, cells - 1, np.clip(vt, 0, cells - 1))
    counts = np.bincount(cell, minlength=cells)
    excl = np.concatenate([[0], np.cumsum(counts)[:-1]])
    bucket_of_cell = excl // t_rows
    bucket = bucket_of_cell[cell]
    vc = np.where(nan, np.float32(cells - 1), np.clip(v, np.float32(0), np.float32(cells - 1))).astype(np.float32)
    fbin = np.zeros(n, dtype=np.int64)
    for b in np.unique(bucket):
        sel = bucket == b
        c0, c1 = cell[sel].min(), cell[sel].max()
        fs = np.float32(fine) / np.float32(c1 - c0 + 1)
        f = ((vc[sel] - np.float32(c0)).astype(np.float32) * fs).astype(np.float32)
        fbin[sel] = np.clip(f.astype(np.int64), 0, fine - 1)
    # the kernels place by (bucket, fine bin) and settle the rest by comparing composites
    return np.lexsort((keys, fbin, bucket)), np.argsort(keys, kind="stable"), bucket, fbin, keys


@settings(max_examples=60, deadline=None)
@given(st.lists(st.one_of(st.floats(-2, 2, width=32), st.floats(width=32, allow_nan=True, allow_infinity=True),
                          st.sampled_from([0.0, -0.0, 0.25, 0.25, 1e-30, -1e-30, float("inf"), float("-inf"), float("nan")])),
                min_size=2, max_size=400), st.integers(0, 3))
def test_hist_sort_bucket_and_bin_are_monotone_in_the_key(vals, mode):
    s = np.asarray(vals, dtype=np.float32)
    kw = {}
    if mode == 1:                                   # a range far off the data: everything lands in the end cells
        kw = {"hi": np.float32(1e3), "scale": np.float32(1e-2)}
    elif mode == 2:                                 # tiny buckets and few fine bins: many boundaries
        kw = {"cells": 64, "t_rows": 8, "fine": 16}
    elif mode == 3:                                 # huge scale: saturating conversions
        kw = {"hi": np.float32(0.5), "scale": np.float32(3e38)}
    fin = s[np.isfinite(s)]
    if mode in (0, 2) and not (fin.size and np.float32(fin.std()) > 0):
        return          # no finite spread: scale = 0, the plan kernel flags the query (hs_plan_kernel) and the sample sort takes it
    got, ref, bucket, fbin, keys = _hs_rank_restatement(s, **kw)
    assert np.array_equal(got, ref)                                        # == the stable descending argsort
    order = ref
    assert np.all(np.diff(bucket[order]) >= 0)                             # bucket is monotone along the key order
    same_bucket = np.diff(bucket[order]) == 0
    assert np.all(np.diff(fbin[order])[same_bucket] >= 0)                  # and so is the fine bin inside a bucket


def test_hist_sort_restatement_on_similarity_like_scores():
    rs = np.random.RandomState(3)
    s = (rs.randn(100000) * 0.044).astype(np.float32)
    s[rs.randint(0, s.size, 50)] = rs.rand(50).astype(np.float32)           # planted neighbours
    s[:6] = [np.inf, -np.inf, np.nan, 0.0, -0.0, np.nan]
    got, ref, bucket, _, _ = _hs_rank_restatement(s)
    assert np.array_equal(got, ref) and np.array_equal(ref, oracle.ranks_from_scores(s[:, None])[:, 0])
    assert np.bincount(bucket).max() <= 4096                                # what one CTA sorts: no query of this shape is flagged


# ---- sample-sort route (csrc/ranks.cu: ss_splitters / ss_scatter): the score -> bucket-range table -----------------------
def _ss_bucket_restatement(s, B, k_tab=1024, oversample=32):
    """numpy fp32 restatement of how the sample sort buckets one query's scores: splitters = order statistics of a
    systematic sample of (score key, row) composites; a table over k_tab cells linear in the score narrows the binary
    search to [lower[cell], lower[cell + 1]].  Returns (bucket by the narrowed search, bucket by the full search)."""
    s = np.asarray(s, dtype=np.float32)
    n = s.shape[0]
    keys = make_keys_host(s, np.arange(n))
    m = min(B * oversample, n)
    samp = np.sort(keys[(np.arange(m, dtype=np.int64) * n) // m])
    spl = np.array([samp[(b * m) // B] for b in range(1, B)], dtype=np.uint64)               # B - 1 splitters
    canon = np.where(s == 0, np.float32(0), s)
    sc_of = lambda kk: keys_to_host(np.asarray([kk], dtype=np.uint64))[0][0] if (int(kk) >> 32) != 0xffffffff else np.float32(np.nan)
    hi, lo = np.float32(sc_of(samp[0])), np.float32(sc_of(samp[-1]))
    with np.errstate(invalid="ignore", over="ignore"):
        rng = np.float32(hi - lo)
        scale = np.float32(k_tab) / np.float32(rng * np.float32(1.0001)) if (rng > 0 and np.isfinite(rng)) else np.float32(0)

        def cell(x):
            x = np.asarray(x, dtype=np.float32)
            if scale == 0:
                return np.zeros(x.shape, dtype=np.int64)
            v = ((hi - x).astype(np.float32) * scale).astype(np.float32)
            nan = np.isnan(v)
            vt = np.where(nan, 0, np.clip(v, -2.0 ** 31, 2.0 ** 31 - 128)).astype(np.int64)
            return np.where(nan, k_tab - 1, np.clip(vt, 0, k_tab - 1))

        spl_sc = np.array([sc_of(k_) for k_ in spl], dtype=np.float32)
        hist = np.bincount(cell(spl_sc) + 1, minlength=k_tab + 2)
        lower = np.cumsum(hist)                                                                # lower[j] = splitters in cells < j
        c = cell(canon)
    full = np.searchsorted(spl, keys, side="right")
    narrowed = np.empty(n, dtype=np.int64)
    for i in range(n):
        a, b = int(lower[c[i]]), int(lower[c[i] + 1])
        narrowed[i] = a + np.searchsorted(spl[a:b], keys[i], side="right")
    return narrowed, full


@settings(max_examples=60, deadline=None)
@given(st.lists(st.one_of(st.floats(-2, 2, width=32), st.floats(width=32, allow_nan=True, allow_infinity=True),
                          st.sampled_from([0.0, -0.0, 0.25, 0.25, 0.25, float("inf"), float("-inf"), float("nan")])),
                min_size=8, max_size=300), st.integers(2, 9))
def test_sample_sort_table_never_narrows_the_search_wrongly(vals, B):
    narrowed, full = _ss_bucket_restatement(vals, B)
    assert np.array_equal(narrowed, full)


def test_sample_sort_table_zero_spread_sample_with_infinite_rows():
    """The regression behind ss_cell's `scale == 0 -> cell 0`: a constant row whose +-inf entries the systematic sample
    misses.  Without it (hi - inf) * 0 = NaN sent those rows through the last cell to the last bucket."""
    s = np.full(1500, 0.25, dtype=np.float32)
    s[[3, 700, 1499]] = np.inf
    s[[5, 9]] = -np.inf
    s[11] = np.nan
    narrowed, full = _ss_bucket_restatement(s, 4)
    assert np.array_equal(narrowed, full) and narrowed[3] == 0 and narrowed[5] == 3 and narrowed[11] == 3
