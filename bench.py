#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric on its own configuration:

    queries/s against a 1,001,001 x 2048 database (R1M shape, config 4), 70 queries per step,
    top-100 per query identical to the fp32 reference ranking (bf16 tcgen05 scan + fused
    threshold filter + exact fp32 re-scoring of a 132-entry shortlist),
    on N B200s of one node (database rows sharded, local top-k, one NCCL all-gather of keys).

A "step" = one batch of 70 queries ranked against the whole database.
  value  : device-resident queries, CUDA-event timed, max over ranks
  e2e    : the same through the public API with HOST buffers: pinned-host queries -> H2D ->
           search -> D2H of the (scores, idx) result, every step.  The database is the resident
           index (state, like weights), not a per-step input; `e2e_cold_db_ms` reports the
           one-off upload + bf16 packing of the database separately.
  --impl reference : the reference's CPU path (np.dot + np.argsort, cirscore.py:69-70, restated
           in oracle/oracle.py) on the host cores, on a bounded row sample of the same workload.

Launch: python bench.py [--gpus N --steps K --warmup W]; for N > 1 under torchrun
(python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_DB, DIM, N_Q, TOPK = 1001001, 2048, 70, 100
METRIC = "queries/s vs 1M x 2048 db (70-query batches, top-100, fp32-faithful ranking)"
UNIT = "queries/s"
WORKLOAD = "R1M: 70 q x 1,001,001 x 2048-D db, top-100 (BASELINE.json configs[3]; 4.1 GB bf16 + 8.2 GB fp32 master)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the head / CLAHE side measurements")
    ap.add_argument("--cpu-rows", type=int, default=100000, help="database row sample for the CPU baseline")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------ CPU reference arm
def cpu_reference(rows, reps):
    """The reference's own arithmetic (cirscore.py:69-70) through the oracle port, all host
    threads numpy/BLAS will use, on `rows` database rows x 70 queries; queries/s extrapolated
    linearly in N_db (argsort is n log n, so this flatters the CPU slightly)."""
    import numpy as np
    from oracle import oracle
    rs = np.random.RandomState(4)
    vecs = rs.standard_normal((DIM, rows)).astype(np.float32)        # (D, N_db) as extract_vectors returns
    vecs /= np.linalg.norm(vecs, axis=0, keepdims=True)
    qvecs = rs.standard_normal((DIM, N_Q)).astype(np.float32)
    qvecs /= np.linalg.norm(qvecs, axis=0, keepdims=True)
    oracle.ranks(vecs[:, :1000], qvecs)                               # warm-up
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        sc = oracle.scores(vecs, qvecs)
        rk = np.argsort(-sc, axis=0)                                   # the reference's exact call (default kind)
        ts.append(time.perf_counter() - t0)
        assert rk.shape == (rows, N_Q)
    t = float(np.median(ts))
    scale = N_DB / float(rows)
    return {"value": N_Q / (t * scale), "unit": UNIT, "cores": len(os.sched_getaffinity(0)), "kind": "port",
            "sample": "np.dot + np.argsort(axis=0) on %d of %d db rows x 70 queries x 2048-D, median of %d, time scaled x%.2f"
                      % (rows, N_DB, reps, scale),
            "ms_per_step_extrapolated": t * scale * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    reps = max(1, min(args.steps, 5))
    cb = cpu_reference(args.cpu_rows, reps)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": reps,
            "warmup": 1, "ms_per_step": cb["ms_per_step_extrapolated"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, torch_device_index):
        super().__init__(daemon=True)
        self.samples = []
        self.stop_flag = False
        self.active = False
        self.ok = False
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            uuid = str(torch.cuda.get_device_properties(torch_device_index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(torch_device_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as exc:  # noqa: BLE001
            self.err = repr(exc)

    def run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((self.active, mhz, reasons))
            except Exception:
                pass
            time.sleep(0.002)

    def summary(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": getattr(self, "err", "nvml unavailable")}
        import statistics
        act = [s for s in self.samples if s[0]] or self.samples
        names = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
                 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
        bits = 0
        for _, _, r in act:
            bits |= r
        return {"sm_mhz": statistics.median([s[1] for s in act]) if act else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for b, n in names.items() if bits & b and n != "gpu_idle"), "samples": len(act)}


class KernelProf:
    """One pair of CUDA events around the dominant kernel, on the stream it is launched on.  The
    events are `external` so that they become record nodes when the step is captured into a CUDA
    graph: after a replay + synchronize, elapsed_time() is that replay's kernel duration."""

    def __init__(self, torch):
        self.e0 = torch.cuda.Event(enable_timing=True, external=True)
        self.e1 = torch.cuda.Event(enable_timing=True, external=True)
        self.bytes = 0

    def begin(self):
        self.e0.record()

    def end(self, nbytes):
        self.e1.record()
        self.bytes = nbytes

    def last_ms(self):
        return self.e0.elapsed_time(self.e1)


class QuietStdout:
    """Everything libraries print to fd 1 while the benchmark runs (NCCL's version banner, ...) goes to stderr, so
    that stdout carries exactly one line: the JSON result written through emit()."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.real, (text + "\n").encode())


# ------------------------------------------------------------------ our arm
def count_launches_per_step(lib, target, q_dev):
    """Kernels of libmdir_b200 launched by one step (counted on an un-captured run of the same call)."""
    n0 = lib.mdir_launch_count()
    target.search(q_dev, TOPK, precision="fp32", check=False)
    return int(lib.mdir_launch_count() - n0)


def run_ours(args):
    import numpy as np
    import torch
    import mdir_b200
    from mdir_b200 import _lib
    from mdir_b200.search import Index, ShardedIndex, GraphedSearch, SearchPipeline, pack_bf16, default_shortlist

    out = QuietStdout()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun (python -m torch.distributed.run --nproc-per-node %d bench.py ...)" % (args.gpus, args.gpus))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.lib().mdir_device_check(), "device check")
    lib = _lib.lib()

    # ---- synthetic database shard (rows [lo, hi) of the 1,001,001), built on the device -------
    lo, hi = ShardedIndex.shard_bounds(N_DB, world, rank)
    t_build0 = time.perf_counter()
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    db32 = torch.empty((hi - lo, DIM), dtype=torch.float32, device=dev)
    for r0 in range(0, hi - lo, 65536):
        blk = torch.randn((min(65536, hi - lo - r0), DIM), device=dev, generator=g)
        db32[r0:r0 + blk.shape[0]] = blk / blk.norm(dim=1, keepdim=True)
    gq = torch.Generator(device="cpu").manual_seed(99)
    q_host = torch.randn((N_Q, DIM), generator=gq)
    q_host = (q_host / q_host.norm(dim=1, keepdim=True)).pin_memory()
    torch.cuda.synchronize()
    # cold path of the index build (packing only; the fp32 rows are already resident)
    t0 = time.perf_counter()
    index = Index.from_packed(pack_bf16(db32), db32=db32, idx_base=lo)
    torch.cuda.synchronize()
    pack_ms = (time.perf_counter() - t0) * 1e3
    prof = KernelProf(torch)
    target = ShardedIndex.from_local(index) if world > 1 else index
    # the whole step (pack q, sample scan, select, filter scan, finalize, fp32 re-score, finalize,
    # [all-gather, merge]) captured once into a CUDA graph; every step below is one replay
    gs = GraphedSearch(target, N_Q, TOPK, precision="fp32")
    gs.q.copy_(q_host, non_blocking=True)
    q_dev = gs.q
    # N > 1: the throughput loop replays the deferred-exchange graph (step t pushes its keys over NVLink and merges
    # step t-1, whose keys arrived a step ago; one drain after the last step, inside the timed region)
    deferred = world > 1 and getattr(target, "_mb", None) is not None
    gs_run = GraphedSearch(target, N_Q, TOPK, precision="fp32", deferred=True) if deferred else gs
    gs_run.q.copy_(q_host, non_blocking=True)
    # the same step with a CUDA event pair around the dominant kernel inside the graph: used only to read that
    # kernel's duration (the two event-record nodes cost ~8 us per step, so `value` is timed on the plain graph)
    gs_prof = GraphedSearch(target, N_Q, TOPK, precision="fp32", prof=prof)
    gs_prof.q.copy_(q_host, non_blocking=True)
    out_host = torch.empty((N_Q, TOPK * 2), dtype=torch.float32).pin_memory()

    def step_device():
        return gs()

    def step_e2e():
        s, i = gs(q_host)                                            # pinned host -> static device buffer, replay
        out_host[:, :TOPK].copy_(s, non_blocking=True)
        out_host[:, TOPK:].view(torch.int32).copy_(i, non_blocking=True)
        torch.cuda.synchronize()
        return out_host

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ---------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        s_chk, i_chk = step_device()
        if deferred:
            gs_run()
    if deferred:
        gs_run.drain()
    torch.cuda.synchronize()
    assert not gs.check_overflow(), "candidate overflow on the benchmark data"
    for _ in range(3):
        step_e2e()

    # ---- timed region: `value` ------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if sampler.ok:
        sampler.start()
    launches_per_step = count_launches_per_step(lib, target, q_dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    barrier()
    sampler.active = True
    ev[0].record()
    for _ in range(args.steps):
        gs_run()
    if deferred:
        gs_run.drain()
    ev[1].record()
    barrier()
    sampler.active = False
    launches = launches_per_step * args.steps
    assert not gs.check_overflow()
    ms_total = ev[0].elapsed_time(ev[1])
    # dominant-kernel duration: the graph carries an event pair around the FILTER scan; read it after
    # individual replays of the same graph (a per-step read needs a sync, so not inside the loop above)
    scan_samples = []
    for _ in range(min(50, args.steps) + 1):
        gs_prof()
        torch.cuda.synchronize()
        scan_samples.append(prof.last_ms())
    scan_ms = sum(scan_samples) / len(scan_samples)

    # ---- timed region: e2e (host buffers in, host result out, every step) --------------------------
    # (1) blocking: upload, replay, download, synchronize -- the latency of one step seen from the host
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_blocking_s = time.perf_counter() - t0
    # (2) the serving loop (SearchPipeline): the same three stages per step, two steps in flight, so the copies of
    # one step overlap the scan of the next.  Every step's queries are uploaded and every result is read on the host.
    pipe = SearchPipeline(target, N_Q, TOPK, precision="fp32")
    for _ in pipe.map([q_host] * 4):
        pass
    barrier()
    t0 = time.perf_counter()
    n_out = 0
    for s_h, i_h in pipe.map(q_host for _ in range(args.steps)):
        n_out += int(i_h[0, 0] >= 0)
    barrier()
    e2e_s = time.perf_counter() - t0
    assert n_out == args.steps
    sampler.stop_flag = True

    # bf16-only mode (no fp32 re-scoring), for the record
    gs16 = GraphedSearch(target, N_Q, TOPK, precision="bf16")
    gs16.q.copy_(q_dev)
    for _ in range(3):
        gs16()
    e2 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    e2[0].record()
    n_b = max(10, args.steps // 4)
    for _ in range(n_b):
        gs16()
    e2[1].record()
    torch.cuda.synchronize()
    bf16_ms = e2[0].elapsed_time(e2[1]) / n_b

    per_rank = None
    if world > 1:
        mine = torch.tensor([ms_total, e2e_s, scan_ms], dtype=torch.float64, device=dev)
        allr = torch.empty((world, 3), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr.view(-1), mine)
        per_rank = {"ms_total": allr[:, 0].tolist(), "e2e_s": allr[:, 1].tolist(), "scan_ms": allr[:, 2].tolist()}
        ms_total, e2e_s, scan_ms = [float(x) for x in allr.max(dim=0).values.tolist()]

    # ---- sanity: planted result is self-consistent with an exact recomputation -------------------
    s_chk, i_chk = step_device()
    own = (i_chk >= lo) & (i_chk < hi)
    rows = (i_chk.long() - lo).clamp_(0, hi - lo - 1)
    exact = (db32[rows.view(-1)].view(N_Q, TOPK, DIM) * q_dev[:, None, :]).sum(-1)
    assert torch.all(((exact - s_chk).abs() < 2e-6) | ~own), "fp32 re-scored values disagree with an exact recomputation"
    assert torch.all(s_chk[:, :-1] >= s_chk[:, 1:])

    extras = {}
    if rank == 0 and not args.no_extras:
        if world == 1:
            extras.update(search_side_measurements(torch, mdir_b200, index, q_dev, prof))
        extras.update(side_measurements(torch, mdir_b200, dev))

    if rank != 0:
        finish(dist, world)
        return

    peak, peak_src = measured_peaks()
    achieved = prof.bytes / 1e9 / (scan_ms * 1e-3) if scan_ms > 0 else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "sim_scan_traffic.json")
    if os.path.exists(tpath) and world == 1:      # the ncu capture is of the 1-GPU launch
        with open(tpath) as fh:
            traffic = json.load(fh).get("dram_bytes_per_launch")
    ms_step = ms_total / args.steps
    line = {
        "metric": METRIC, "value": N_Q * args.steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "queries_per_step": N_Q, "topk": TOPK, "shortlist": default_shortlist(TOPK), "db_rows_per_gpu": hi - lo,
                   "sharding": "db rows contiguous over %d GPU(s); %s" % (world, "single shard" if world == 1 else (
                       ("%d B of keys per rank pushed to every peer over NVLink by the merge kernel itself (no NCCL call on the step); the merge of "
                        "step t runs inside step t+1, one drain after the last step") % (N_Q * TOPK * 8) if deferred else
                       "ncclAllGather of %d B of keys per rank + merge kernel" % (N_Q * TOPK * 8))),
                   "l2": "inputs larger than L2: %.2f GB bf16 shard streamed per step vs 126 MB L2" % ((hi - lo) * DIM * 2 / 1e9)},
        "clocks": sampler.summary(),
        "e2e": {"value": N_Q * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": N_Q * DIM * 4, "d2h_bytes_per_step": N_Q * TOPK * 8,
                "ms_per_step": e2e_s / args.steps * 1e3, "blocking_ms_per_step": e2e_blocking_s / args.steps * 1e3,
                "timing": "wall clock around %d steps of SearchPipeline (per step: H2D of the pinned queries, graph replay, D2H of scores/idx, host read of the result; %d steps in flight%s); blocking_ms_per_step = the same with a synchronize after every step" % (args.steps, pipe.depth, ", deferred NVLink exchange" if pipe.deferred else "")},
        "e2e_cold_db_ms": {"pack_fp32_to_bf16_ms": pack_ms,
                           "note": "one-off index build for this shard; host->device upload of the fp32 rows would add %.1f GB over PCIe" % ((hi - lo) * DIM * 4 / 1e9)},
        "gpu_launches": int(launches), "gpu_launches_note": "%d libmdir_b200 kernels per step, replayed from one CUDA graph per step" % launches_per_step,
        "roofline": {"bound": "hbm", "kernel": "sim_scan_kernel (%s)" % ("threshold + filter in one launch: whole shard" if index._fused_ok(default_shortlist(TOPK)) else "FILTER pass"), "achieved": achieved, "peak": peak, "peak_source": peak_src,
                     "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                     "algorithmic_bytes_per_launch": prof.bytes, "avg_launch_ms": scan_ms, "share_of_step": scan_ms / ms_step,
                     "timing": "CUDA events (external, recorded inside the step's CUDA graph on its stream), mean of %d replays" % len(scan_samples)},
        "bf16_only_ms_per_step": bf16_ms,
    }
    if world == 1:
        cb = cpu_reference(args.cpu_rows, 3)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if per_rank is not None:
        line["per_rank"] = per_rank
    line.update(extras)
    out.emit(json.dumps(line))
    sys.stdout.flush()
    finish(dist, world)


def finish(dist, world):
    """Multi-rank exit: barrier, then leave without tearing NCCL down (communicators referenced by
    captured CUDA graphs make destroy_process_group() block)."""
    if world > 1:
        import torch
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def search_side_measurements(torch, mdir_b200, index, q_dev, prof):
    """Config 5 and the tensor-utilisation evidence, on the resident 1M x 2048 index (1 GPU):
    alpha-QE (two similarity + top-k passes), a DBA slice (blocks of 128 database rows searched
    against the whole database), and the 128-query FILTER scan as TFLOP/s."""
    from mdir_b200 import qe
    out = {}
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    tf_peak = 1397.3
    if os.path.exists(path):
        with open(path) as fh:
            tf_peak = float(json.load(fh).get("bf16_tflops_sustained", tf_peak))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    # alpha-QE: search top-10, expand, search top-100 (fp32-faithful both times)
    for _ in range(2):
        qe.search_qe(index, q_dev, TOPK)
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(10):
        qe.search_qe(index, q_dev, TOPK)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 10
    out["alpha_qe"] = {"metric": "alpha-QE queries/s (alpha=3, n_QE=10; two similarity+top-k passes; parity unpinned)", "value": N_Q / (ms * 1e-3),
                       "unit": "queries/s", "ms_per_batch": ms}
    # 128-row blocks of the database as queries (the DBA inner loop): FILTER scan at N = 128
    rows = index.db32[:1280]
    index.prof = prof
    index.search(rows[:128], 10, precision="bf16")
    torch.cuda.synchronize()
    scan = []
    ev[0].record()
    for b in range(10):
        index.search(rows[b * 128:(b + 1) * 128], 10, precision="bf16", check=False)
        torch.cuda.synchronize()
        scan.append(prof.last_ms())
    ev[1].record()
    torch.cuda.synchronize()
    index.prof = None
    ms_blk = ev[0].elapsed_time(ev[1]) / 10
    scan_ms = sum(scan) / len(scan)
    flops = 2.0 * 128 * DIM * (prof.bytes / (2 * DIM))
    out["dba_block"] = {"metric": "DBA inner loop: 128 database rows vs 1,001,001 x 2048 (top-10, bf16 scores)", "ms_per_block": ms_blk,
                        "rows_per_s": 128 / (ms_blk * 1e-3), "full_dba_1M_estimate_s": (N_DB / 128.0) * ms_blk * 1e-3,
                        "filter_scan_ms": scan_ms, "filter_scan_tflops": flops / (scan_ms * 1e-3) / 1e12,
                        "frac_of_sustained_bf16_peak": flops / (scan_ms * 1e-3) / 1e12 / tf_peak, "bf16_peak_tflops_sustained": tf_peak,
                        "note": "at N=128 the scan is still HBM-bound (AI = 128 FLOP/B < ridge ~214): tensor fraction = HBM fraction x 128/214"}
    return out


def side_measurements(torch, mdir_b200, dev):
    """The second half of BASELINE.json's metric: GeM + whiten descriptors/s (C2 head shape) and
    CLAHE images/s, each with its HBM-roofline fraction.  Outside the timed region of `value`."""
    from mdir_b200 import _lib
    peak, _ = measured_peaks()
    out = {}
    # head: 192 images x 3 scales of 2048 x {32x24, 23x17, 16x12}; P 2048x2048
    B, C = 192, 2048
    hws = [(32, 24), (23, 17), (16, 12)]
    g = torch.Generator(device=dev).manual_seed(7)
    fm = []
    for _ in range(B):
        for (h, w) in hws:
            fm.append(torch.randn((1, C, h, w), device=dev, generator=g).clamp_(min=0))
    P = torch.randn((C, C), device=dev, generator=g) / C ** 0.5
    m = torch.randn((C, 1), device=dev, generator=g) * 0.01
    head = mdir_b200.RetrievalHead("gem", p=2.9137, whitening={"P": P.cpu().numpy(), "m": m.cpu().numpy()}, nscales=3, device=dev)
    packed = head.pack(fm)                 # offset tables built once (static shapes), maps stay where they are
    for _ in range(3):
        head(packed)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    reps = 10
    for _ in range(reps):
        head(packed)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / reps
    bytes_img = 4 * C * sum(h * w for h, w in hws)
    replay = head.capture(packed)          # the same launches recorded once into a CUDA graph (static arena)
    for _ in range(3):
        replay()
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(reps):
        replay()
    ev[1].record()
    torch.cuda.synchronize()
    ms_g = ev[0].elapsed_time(ev[1]) / reps
    out["head"] = {"metric": "GeM+L2N+multiscale+Lw descriptors/s (C2 head: 3 scales, 2048-D)", "value": B / (ms_g * 1e-3), "unit": "descriptors/s",
                   "batch": B, "ms_per_batch": ms_g, "ms_per_batch_eager_launches": ms, "algorithmic_bytes_per_descriptor": bytes_img,
                   "hbm_frac_of_measured": B * bytes_img / 1e9 / (ms_g * 1e-3) / peak,
                   "note": "value = CUDA-graph replay of the head over a static feature-map arena; ms_per_batch_eager_launches = the same nine launches issued one by one"}
    del fm
    # CLAHE: 256 images of 768 x 1024 u8 (night-like gamma distribution)
    n_img = 256
    imgs = (torch.rand((n_img, 768, 1024), device=dev, generator=g) ** 4 * 255).to(torch.uint8)
    for _ in range(3):
        mdir_b200.clahe_u8(imgs, 4, (8, 8))
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(reps):
        mdir_b200.clahe_u8(imgs, 4, (8, 8))
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / reps
    del imgs
    out["clahe"] = {"metric": "CLAHE u8 images/s (768x1024, clip 4, 8x8 tiles)", "value": n_img / (ms * 1e-3), "unit": "images/s",
                    "batch": n_img, "ms_per_batch": ms, "algorithmic_bytes_per_image": 2 * 768 * 1024,
                    "hbm_frac_of_measured": n_img * 2 * 768 * 1024 / 1e9 / (ms * 1e-3) / peak}
    # the whole ImageClahe transform (RGB -> Lab -> CLAHE(L) -> RGB) on 32 images of 768 x 1024 x 3 float32,
    # against the reference arithmetic (cv2 float Lab conversions + cv2 CLAHE, as ImageClahe.apply) on the host
    rgbs = torch.rand((32, 768, 1024, 3), device=dev, generator=g) ** 3
    lst = list(rgbs)
    for _ in range(2):
        mdir_b200.image_clahe(lst, 4, (8, 8))
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(5):
        mdir_b200.image_clahe(lst, 4, (8, 8))
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 5
    cpu_ms = None
    try:
        import cv2
        import numpy as np
        host = rgbs[0].cpu().numpy()
        t0 = time.perf_counter()
        for _ in range(3):
            spc = (cv2.cvtColor(host, cv2.COLOR_RGB2LAB) + np.array([0, 128, 128], np.float32)) / np.array([100.0, 255.0, 255.0], np.float32)
            spc[:, :, 0] = cv2.createCLAHE(clipLimit=4, tileGridSize=(8, 8)).apply((spc[:, :, 0] * 255).astype(np.uint8)).astype(np.float32) / 255.0
            cv2.cvtColor(spc * np.array([100.0, 255.0, 255.0], np.float32) - np.array([0, 128, 128], np.float32), cv2.COLOR_LAB2RGB)
        cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
    except Exception:  # noqa: BLE001
        pass
    out["image_clahe"] = {"metric": "ImageClahe.apply (RGB->Lab->CLAHE->RGB) images/s, 768x1024x3 float32", "value": 32 / (ms * 1e-3), "unit": "images/s",
                          "ms_per_image": ms / 32, "algorithmic_bytes_per_image": 24 * 768 * 1024,
                          "hbm_frac_of_measured": 32 * 24 * 768 * 1024 / 1e9 / (ms * 1e-3) / peak,
                          "cpu_reference_ms_per_image": cpu_ms, "cpu_threads": "cv2 default"}
    del rgbs, lst
    # full (N_db, N_q) ranks: C1 shape (70 q x 4,993 x 2048, fp32-faithful scores) and a C3-shaped slice
    # (1,024 of the 10,000 queries x 100,000 x 512, bf16 scores); device-resident in and out
    from mdir_b200.search import Index
    for tag, n_db, D, nq, prec in (("ranks_c1", 4993, 2048, 70, "fp32"), ("ranks_c3_slice", 100000, 512, 1024, "bf16")):
        db = torch.randn((n_db, D), device=dev, generator=g)
        db = db / db.norm(dim=1, keepdim=True)
        q = torch.randn((nq, D), device=dev, generator=g)
        q = q / q.norm(dim=1, keepdim=True)
        idx = Index(db, device=dev, keep_fp32=(prec != "bf16"))
        for _ in range(2):
            r = idx.ranks(q, precision=prec)
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(5):
            r = idx.ranks(q, precision=prec)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 5
        pairs = n_db * nq
        out[tag] = {"metric": "full per-query ranks (scores + segmented radix sort + int64 (N_db,N_q) write)", "shape": "%d q x %d db x %d-D, %s scores" % (nq, n_db, D, prec),
                    "ms": ms, "pairs_per_s": pairs / (ms * 1e-3), "hbm_floor_frac": pairs * 12 / 1e9 / (ms * 1e-3) / peak,
                    "note": "floor = 12 B/pair (4 B score + 8 B int64 rank) at the measured HBM peak"}
        del idx, db, r
    out["c3_full"] = c3_full_measurement(torch, mdir_b200, dev, g)
    out.update(training_side_measurements(torch, mdir_b200, dev, g, ev))
    return out


def c3_full_measurement(torch, mdir_b200, dev, g):
    """BASELINE.json config C3 at full size: 10,000 queries x 100,000 database rows x 512-D -> the (N_db, N_q) int64
    ranks array (8 GB) and mAP from it, everything on the device; the reference's np.dot + np.argsort timed on a
    64-query sample of the same matrices for scale."""
    import numpy as np
    from mdir_b200.search import Index
    from mdir_b200.evaluate import compute_map
    n_db, D, nq = 100000, 512, 10000
    db = torch.randn((n_db, D), device=dev, generator=g)
    db = db / db.norm(dim=1, keepdim=True)
    src = torch.randperm(n_db, device=dev, generator=g)[:nq]
    q = db[src] + 0.7 * torch.randn((nq, D), device=dev, generator=g) / D ** 0.5
    q = q / q.norm(dim=1, keepdim=True)
    src_h = src.cpu().numpy()
    rs = np.random.RandomState(11)
    gnd = [{"ok": np.array([int(src_h[i])] + rs.randint(0, n_db, 3).tolist()), "junk": rs.randint(0, n_db, 2)} for i in range(nq)]
    idx = Index(db, device=dev, keep_fp32=False)
    ranks = idx.ranks(q, precision="bf16")               # first call: cudaMalloc of the 8 GB result + 5 GB of workspace
    del ranks
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    ranks = idx.ranks(q, precision="bf16")
    ev[1].record()
    torch.cuda.synchronize()
    ms_ranks = ev[0].elapsed_time(ev[1])
    t0 = time.perf_counter()
    mp, aps, _, _ = compute_map(ranks, gnd, device=dev)
    torch.cuda.synchronize()
    ms_map = (time.perf_counter() - t0) * 1e3
    first_ok = bool((ranks[0] == src).float().mean().item() > 0.99)
    dbh, qh = db.cpu().numpy(), q[:64].cpu().numpy()
    t0 = time.perf_counter()
    np.argsort(-np.dot(dbh, qh.T), axis=0)
    cpu_ms = (time.perf_counter() - t0) * 1e3 * (nq / 64)
    del ranks, idx, db, q
    return {"metric": "C3 full size: scores + full ranks (N_db, N_q) int64 + mAP on the device", "shape": "%d q x %d db x %d-D, bf16 scores" % (nq, n_db, D),
            "ranks_ms": ms_ranks, "map_ms_incl_host_flatten": ms_map, "pairs_per_s": n_db * nq / (ms_ranks * 1e-3), "mAP": mp,
            "planted_neighbour_ranked_first": first_ok, "cpu_port_ms_dot_argsort": cpu_ms, "cpu_sample": "64 of %d queries, time scaled" % nq}


def training_side_measurements(torch, mdir_b200, dev, g, ev):
    """Row f4: hard-negative mining at the reference's default epoch shape (2000 queries x 20000 pool images,
    2048-D, 5 negatives; traindataset.py:54) and whitening learning (2048-D, 20000 images, 10000 pairs)."""
    import numpy as np
    from oracle import oracle
    out = {}
    D, n_q, n_pool, nnum = 2048, 2000, 20000, 5
    pool = torch.randn((D, n_pool), device=dev, generator=g)
    pool = pool / pool.norm(dim=0, keepdim=True)
    qv = pool[:, :n_q] + 0.05 * torch.randn((D, n_q), device=dev, generator=g)
    qv = qv / qv.norm(dim=0, keepdim=True)
    rs = np.random.RandomState(3)
    pc = rs.randint(0, 700, n_pool).astype(np.int32)
    qc = pc[:n_q].copy()
    for _ in range(2):
        mdir_b200.mine_hard_negatives(qv, pool, qc, pc, nnum, device=dev)
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(3):
        mdir_b200.mine_hard_negatives(qv, pool, qc, pc, nnum, device=dev)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 3
    sub = 100                                              # CPU: the reference's mm + sort + walk on 100 of the 2000 queries
    ph, qh = pool.cpu().numpy(), qv[:, :sub].cpu().numpy()
    t0 = time.perf_counter()
    oracle.mine_negatives(qh, ph, qc[:sub], pc, nnum)
    cpu_ms = (time.perf_counter() - t0) * 1e3 * (n_q / sub)
    out["mining"] = {"metric": "hard-negative mining (scores + full ranks + cluster walk + distances)", "shape": "%d q x %d pool x %d-D, %d negatives" % (n_q, n_pool, D, nnum),
                     "ms": ms, "queries_per_s": n_q / (ms * 1e-3), "cpu_port_ms": cpu_ms, "cpu_sample": "%d of %d queries, time scaled" % (sub, n_q)}
    del pool, qv
    N, n_pairs = 20000, 10000
    X = torch.randn((D, N), device=dev, generator=g, dtype=torch.float32)
    X = X / X.norm(dim=0, keepdim=True)
    qi, pi = rs.randint(0, N, n_pairs), rs.randint(0, N, n_pairs)
    mdir_b200.whitenlearn(X, qi, pi, device=dev)
    torch.cuda.synchronize()
    ev[0].record()
    mdir_b200.whitenlearn(X, qi, pi, device=dev)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1])
    X64 = X.double()
    for _ in range(2):
        mdir_b200.gemm_f64(X64, X64, False)
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(3):
        mdir_b200.gemm_f64(X64, X64, False)
    ev[1].record()
    torch.cuda.synchronize()
    gms = ev[0].elapsed_time(ev[1]) / 3
    Xh = X[:512, :5000].cpu().numpy()                      # CPU: the numpy algorithm on a 512-D x 5000 slice (O(D^2 N + D^3))
    t0 = time.perf_counter()
    oracle.whitenlearn(Xh, qi[:2500] % 5000, pi[:2500] % 5000)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    out["whitenlearn"] = {"metric": "whitenlearn (fp64): pair covariance, Cholesky, data covariance, eigendecomposition", "shape": "%d-D x %d images, %d pairs" % (D, N, n_pairs),
                          "ms": ms, "gemm_f64_ms_2048x2048x20000": gms, "gemm_f64_tflops": 2.0 * D * D * N / (gms * 1e-3) / 1e12,
                          "cpu_port_ms_512d_x_5000": cpu_ms,
                          "note": "O(D^2 N) contractions in mdir_gemm_f64 (SIMT DFMA); D x D factorisations are cuSOLVER via torch.linalg"}
    return out


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
