#!/usr/bin/env python
"""
bench.py -- the hot-path benchmark (contract: one JSON line on stdout from rank 0).

Workload at N GPUs (BASELINE.json configs[1], the config the metric is quoted on):
  * database: 10 M random hg38-shaped intervals (24 chromosomes, start~U, len~U{1..2000}), one device index per rank
    holding the chromosomes LPT-assigned to that rank;
  * queries: N x 10 M intervals of the same law ("weak" scaling: 10 M queries per GPU), routed to the rank that owns
    their chromosome; no data-path collective; the per-chromosome hit counters are summed with one NCCL all-reduce.
  A "step" is one batched find() over the rank's queries: count pass, exclusive scan, fill pass -> ordered CSR hit
  lists (bit-identical to bx.intervals.intersection order; checked against the oracle on a sample every run).

  value  = queries/s with queries and results resident in HBM (CUDA events on the library stream, max over ranks)
  e2e    = queries/s through the host API with pinned HOST buffers: H2D of (chrom,start,end), find, D2H of
           (offsets, hits) inside the timed region
  roofline      = the dominant find kernel: algorithmic bytes / CUDA-event time vs the measured HBM peak
  cpu_baseline  = the reference's Cython IntervalTree (oracle/_ref) timed on this box's host cores (bounded sample)
  extra.bitset  = BinnedBitSet AND+count over 24 chromosome-length bitmaps (BASELINE configs[2]), GB/s and HBM fraction

--impl reference: the unmodified reference (oracle/_ref, compiled from /root/reference by oracle/Makefile) on all
host cores, sharded per chromosome with multiprocessing, same metric/config; rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from bx_python_b200 import synth  # noqa: E402
from bx_python_b200.dist import Comm, env_rank, lpt_assign  # noqa: E402

METRIC = "interval_overlap_queries_per_sec"
UNIT = "queries/s"
N_DB = 10_000_000
NQ_PER_GPU = 10_000_000


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------------------------------
def make_workload(world, n_db, nq_per_gpu):
    """-> per-chromosome database [(s,e)] and queries [(s,e)] (numpy int32), and the chromosome -> rank shards."""
    db = synth.genome_intervals(n_db, 2001)
    qq = synth.genome_intervals(nq_per_gpu * world, 2002)
    weights = [len(q[0]) + 0.25 * len(d[0]) for q, d in zip(qq, db)]
    shards = lpt_assign(weights, world)
    return db, qq, shards


def flatten(per_chrom, chroms):
    """Concatenate the selected chromosomes; tree ids are LOCAL (0..len(chroms)-1)."""
    tid = np.concatenate([np.full(len(per_chrom[c][0]), k, np.int32) for k, c in enumerate(chroms)])
    s = np.concatenate([per_chrom[c][0] for c in chroms])
    e = np.concatenate([per_chrom[c][1] for c in chroms])
    return tid, s, e


# ---------------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self, t0, t1):
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[1]) for r in rows if r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            for k, nme in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(rows)}


# ---------------------------------------------------------------------------------------------------------------------
# reference (CPU) arm
# ---------------------------------------------------------------------------------------------------------------------
def _ref_worker(args):
    """Build the reference IntervalTree for some chromosomes and time find() on a bounded query sample."""
    chroms, db, qs_list, repeats = args
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    from bx.intervals.intersection import IntervalTree  # the unmodified reference, compiled by oracle/Makefile
    t0 = time.perf_counter()
    trees = []
    for (s, e) in db:
        t = IntervalTree()
        ins = t.insert
        for a, b in zip(s.tolist(), e.tolist()):
            ins(a, b, None)
        trees.append(t)
    build = time.perf_counter() - t0
    times, hits, nq = [], 0, 0
    for _ in range(repeats):
        t0 = time.perf_counter()
        hits = nq = 0
        for t, (qs, qe) in zip(trees, qs_list):
            f = t.find
            for a, b in zip(qs.tolist(), qe.tolist()):
                hits += len(f(a, b))
            nq += len(qs)
        times.append(time.perf_counter() - t0)
    return build, times, hits, nq


def reference_rate(db, qq, sample_frac, workers, repeats=1, chroms=None):
    """queries/s of the compiled reference on `workers` processes; chromosome-sharded; find time only.
    Every chromosome keeps its FULL database (so hits/query match the GPU workload); queries are subsampled."""
    import multiprocessing as mp
    chroms = list(range(len(db))) if chroms is None else chroms
    nq = {c: max(1, int(len(qq[c][0]) * sample_frac)) for c in chroms}
    jobs = []
    if workers >= 2 * len(chroms):
        # more cores than chromosomes: several processes per chromosome, each with its own copy of that chromosome's
        # tree (build not timed) and a slice of its queries -- the way a user would spread the reference over cores
        parts = workers // len(chroms)
        for c in chroms:
            for k in range(parts):
                sl = slice(k * nq[c] // parts, (k + 1) * nq[c] // parts)
                jobs.append(([c], [db[c]], [(qq[c][0][:nq[c]][sl], qq[c][1][:nq[c]][sl])], repeats))
    else:
        shards = lpt_assign([len(db[c][0]) + nq[c] for c in chroms], min(workers, len(chroms)))
        for sh in shards:
            cs = [chroms[i] for i in sh]
            jobs.append((cs, [db[c] for c in cs], [(qq[c][0][:nq[c]], qq[c][1][:nq[c]]) for c in cs], repeats))
    if len(jobs) == 1:
        res = [_ref_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(len(jobs)) as pool:
            res = pool.map(_ref_worker, jobs)
    total_q = sum(r[3] for r in res)
    total_hits = sum(r[2] for r in res)
    per_step = [max(r[1][k] for r in res) for k in range(repeats)]      # slowest shard bounds each step
    return {"rates": [total_q / t for t in per_step], "times": per_step, "queries": total_q, "hits": total_hits,
            "build_s": max(r[0] for r in res), "workers": len(jobs)}


def usable_cores():
    """Cores this process may actually use: affinity mask and cgroup CPU quota, not just os.cpu_count()."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(int(txt[0]) / int(txt[1]))))
            else:
                q = int(txt[0])
                per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                if q > 0:
                    n = min(n, max(1, q // per))
        except (OSError, ValueError, IndexError):
            pass
    return n


def run_reference(args):
    rank, world, _ = env_rank()
    if rank != 0:
        return
    from oracle import oracle as orc
    if not orc.ref_available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (needs /root/reference at build time)"}))
        return
    # load the reference's extension module in THIS process too (the workers are forked from it), so the driver's record
    # of loaded native libraries shows which implementation this arm ran
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    import bx.intervals.intersection as _ref_ix
    log(f"reference module: {_ref_ix.__file__}")
    db, qq, _ = make_workload(max(1, args.gpus), args.n_db, args.nq)
    cores = usable_cores()
    # measured on the GPU box (cgroup quota 16 CPUs of a 2 x 32-core Xeon 8562Y+): 8 procs 3.8e6, 24 procs 6.5e6,
    # 48 procs 6.1e6 q/s -> mild oversubscription of the quota is the reference's best case; use it
    workers = max(1, min(24, int(cores * 1.5)))
    if args.ref_workers:
        workers = args.ref_workers
    frac = args.ref_sample
    r = reference_rate(db, qq, frac, workers, repeats=args.steps + args.warmup)
    times = r["times"][args.warmup:]
    val = r["queries"] * len(times) / sum(times)
    sample = (f"all 24 chromosomes at full database density ({args.n_db} intervals, built once in {r['build_s']:.1f} s, "
              f"not timed), first {frac:.3%} of each chromosome's queries ({r['queries']} queries, {r['hits']} hits) "
              f"per step; {r['workers']} processes on {cores} host cores (every process holds the full tree of its "
              f"chromosome and a slice of that chromosome's queries)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(args, max(1, args.gpus)),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": r["workers"], "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_cores": cores,
    }
    print(json.dumps(line))


def workload_config(args, world):
    return {"workload": "find: 10M hg38-shaped intervals (24 chromosomes) vs 10M queries per GPU, ordered CSR hit lists",
            "n_intervals": args.n_db, "n_queries": args.nq * world, "queries_per_gpu": args.nq,
            "sharding": "per-chromosome LPT, no data-path collective; NCCL all-reduce of the per-chromosome counters "
                        "(hit counts here; inside the device-timed step for configs[3]/[4], see roofline.c4_* / c5_*)",
            "other_configs": "roofline.bitset_and_* = configs[2] (24 x BinnedBitSet(250 Mbp) AND), roofline.c4_* = configs[3] "
                             "(bed_intersect 50M x 50M), roofline.c5_* = configs[4] (aggregate 100M scores / 5M windows): all "
                             "ranks, chromosomes sharded, parity against the compiled reference's full-size goldens in the same run",
            "l2_policy": "inputs larger than L2 (index 160 MB + queries 120 MB + results >300 MB per step vs 126 MB L2)"}


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import ctypes as C

    rank, world, local_rank = env_rank()
    os.environ.setdefault("BXB200_DEVICE", str(local_rank))
    # stdout carries exactly one JSON line: keep NCCL's "NCCL version ..." banner (printed to stdout at
    # NCCL_DEBUG=VERSION) out of it
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    from bx_python_b200 import _lib
    from bx_python_b200._lib import check, ptr
    from bx_python_b200.intervals import IntervalForest
    L = _lib.lib()
    numa = _lib.bind_to_gpu_numa_node()       # sysfs, or a copy-rate probe where sysfs reports no node; never raises
    comm = Comm("nccl")
    info = _lib.device_info()
    t_gen = time.perf_counter()
    db, qq, shards = make_workload(world, args.n_db, args.nq)
    mine = shards[rank]
    tid, s, e = flatten(db, mine)
    qt, qs, qe = flatten(qq, mine)
    # shuffle queries so they arrive in file order (not grouped by chromosome), as a BED file would deliver them
    perm = np.random.default_rng(7 + rank).permutation(len(qs))
    qt, qs, qe = qt[perm], qs[perm], qe[perm]
    nq = len(qs)
    if rank == 0:
        log(f"numa: {numa}")
        log(f"device {info['name']} x{world}; rank0 owns chroms {mine}: {len(s)} intervals, {nq} queries "
            f"(gen {time.perf_counter() - t_gen:.1f}s)")

    # ---- build (not part of the step; reported) ----------------------------------------------------------------------
    forest = IntervalForest(len(mine))
    timer = _lib.Timer()
    d_tid, d_s, d_e = _lib.DeviceBuffer(tid), _lib.DeviceBuffer(s), _lib.DeviceBuffer(e)
    build_ms = []
    for _ in range(3):
        timer.start()
        check(L.bxg_itree_build(forest.handle, d_tid.ptr, d_s.ptr, d_e.ptr, len(s), len(mine), _lib.DEVICE))
        timer.stop()
        build_ms.append(timer.elapsed_ms())
    forest.n = len(s)
    del d_tid, d_s, d_e

    # ---- value: queries + results resident in HBM ---------------------------------------------------------------------
    d_qt, d_qs, d_qe = _lib.DeviceBuffer(qt), _lib.DeviceBuffer(qs), _lib.DeviceBuffer(qe)
    total = C.c_int64()

    def step_dev():
        check(L.bxg_itree_find(forest.handle, d_qt.ptr, d_qs.ptr, d_qe.ptr, nq, _lib.DEVICE, C.byref(total)))

    for _ in range(args.warmup):
        step_dev()
    _lib.sync()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    comm.barrier()
    _lib.sync()
    check(L.bxg_launch_count_reset())
    t0 = time.perf_counter()
    timer.start()
    for _ in range(args.steps):
        step_dev()
    timer.stop()
    ms = timer.elapsed_ms()
    _lib.sync()
    launches = _lib.launch_count()
    # the timed region lasts only tens of milliseconds, shorter than nvidia-smi's sampling period: keep the same step
    # running (untimed) for ~1.2 s so that the clock / throttle record is taken under this very load
    t_load = time.perf_counter()
    while rank == 0 and time.perf_counter() - t_load < 1.2:
        for _ in range(20):
            step_dev()
        _lib.sync()
    t1 = time.perf_counter()
    comm.barrier()
    ms_max = float(comm.allreduce_max_f64(np.array([ms]))[0])
    hits_total = total.value
    q_all = int(comm.allreduce_sum_i64(np.array([nq]))[0])
    value = q_all * args.steps / (ms_max * 1e-3)

    # ---- A/B: the single-pass implementation (ticketed tiles + decoupled look-back) on the same inputs ----------------
    check(L.bxg_set_find_mode(1))
    for _ in range(2):
        step_dev()
    timer.start()
    for _ in range(args.steps):
        step_dev()
    timer.stop()
    single_pass_ms = timer.elapsed_ms() / args.steps
    # ... and the three passes with the count of chunk k+1 overlapping the fill of chunk k on a second stream
    check(L.bxg_set_find_mode(2))
    for _ in range(2):
        step_dev()
    timer.start()
    for _ in range(args.steps):
        step_dev()
    timer.stop()
    overlap_ms = timer.elapsed_ms() / args.steps
    check(L.bxg_set_find_mode(-1))
    step_dev()

    # ---- the same queries in sorted-BED order (chrom, start): what locality buys (not the headline: `value` is shuffled) --
    order = np.lexsort((qs, qt))
    s_qt, s_qs, s_qe = (_lib.DeviceBuffer(a[order]) for a in (qt, qs, qe))
    tot2 = C.c_int64()
    for _ in range(2):
        check(L.bxg_itree_find(forest.handle, s_qt.ptr, s_qs.ptr, s_qe.ptr, nq, _lib.DEVICE, C.byref(tot2)))
    timer.start()
    for _ in range(args.steps):
        check(L.bxg_itree_find(forest.handle, s_qt.ptr, s_qs.ptr, s_qe.ptr, nq, _lib.DEVICE, C.byref(tot2)))
    timer.stop()
    sorted_ms = timer.elapsed_ms() / args.steps
    assert tot2.value == hits_total
    del s_qt, s_qs, s_qe
    step_dev()

    # ---- per-kernel CUDA-event times for the roofline (separate pass so event overhead is not in `value`) --------------
    _lib.profile_enable(True)
    for _ in range(args.steps):
        step_dev()
    prof = _lib.profile_report()
    _lib.profile_enable(False)
    clock_summary = None
    if rank == 0:
        time.sleep(0.2)
        clocks.stop()
        clock_summary = clocks.summary(t0, t1)

    # ---- parity spot check against the oracle (every run; outside the timed regions) -----------------------------------
    parity = None
    if rank == 0:
        parity = spot_check(forest, tid, s, e, qt, qs, qe, total.value)

    # ---- per-chromosome hit counts, reduced over ranks with NCCL -------------------------------------------------------
    off = np.empty(nq + 1, np.int64)
    check(L.bxg_itree_fetch(forest.handle, ptr(off), None))
    cnt = np.diff(off)
    per_chrom = np.zeros(24, np.int64)
    np.add.at(per_chrom, np.asarray(mine)[qt], cnt)
    per_chrom = comm.allreduce_sum_i64(per_chrom)
    hits_all = int(per_chrom.sum())

    # ---- e2e: pinned host buffers in, host CSR out ---------------------------------------------------------------------
    h_qt, h_qs, h_qe = (_lib.PinnedArray(nq, np.int32) for _ in range(3))
    h_qt.array[:], h_qs.array[:], h_qe.array[:] = qt, qs, qe
    h_off = _lib.PinnedArray(nq + 1, np.int64)
    h_hits = _lib.PinnedArray(int(hits_total * 1.05) + 1024, np.int32)

    p_off, p_hits = C.c_void_p(), C.c_void_p()

    def step_host_serial():      # upload, find, download one after the other (round-1a path; kept as a comparison)
        check(L.bxg_itree_find(forest.handle, ptr(h_qt.array), ptr(h_qs.array), ptr(h_qe.array), nq, _lib.HOST,
                               C.byref(total)))
        check(L.bxg_itree_fetch(forest.handle, ptr(h_off.array), ptr(h_hits.array)))

    def step_host():             # the public host call: copies overlapped with the kernels
        check(L.bxg_itree_find_host(forest.handle, ptr(h_qt.array), ptr(h_qs.array), ptr(h_qe.array), nq,
                                    C.byref(p_off), C.byref(p_hits), C.byref(total)))

    def time_host(fn, steps):
        for _ in range(max(2, args.warmup // 2)):
            fn()
        comm.barrier()
        _lib.sync()
        tw = time.perf_counter()
        for _ in range(steps):
            fn()
        _lib.sync()
        dt = time.perf_counter() - tw
        return float(comm.allreduce_max_f64(np.array([dt]))[0])

    e2e_steps = max(1, min(args.steps, 5))
    e2e_serial = q_all * e2e_steps / time_host(step_host_serial, e2e_steps)
    assert np.array_equal(h_off.array, off), "host-path offsets differ from device-path offsets"
    serial_hits = h_hits.array[:hits_total].copy()
    e2e_i64 = q_all * e2e_steps / time_host(step_host, e2e_steps)
    r_off = np.frombuffer((C.c_int64 * (nq + 1)).from_address(p_off.value), np.int64)
    r_hits = np.frombuffer((C.c_int32 * total.value).from_address(p_hits.value), np.int32)
    assert np.array_equal(r_off, off) and np.array_equal(r_hits, serial_hits), "pipelined host path differs"
    # the same call with int32 CSR offsets (IntervalForest.find_batch(..., offsets32=True)): the return leg is PCIe-bound
    # and the offsets are 8 of its ~34 bytes per query.  Every rank's hits fit int32 here; if they did not, the int64
    # figure above would be the headline.
    off_bytes = 8
    e2e_value = e2e_i64
    if hits_total < 2**31:
        def step_host32():
            check(L.bxg_itree_find_host32(forest.handle, ptr(h_qt.array), ptr(h_qs.array), ptr(h_qe.array), nq,
                                          C.byref(p_off), C.byref(p_hits), C.byref(total)))
        e2e_value = q_all * e2e_steps / time_host(step_host32, e2e_steps)
        r_off32 = np.frombuffer((C.c_int32 * (nq + 1)).from_address(p_off.value), np.int32)
        r_hits = np.frombuffer((C.c_int32 * total.value).from_address(p_hits.value), np.int32)
        assert np.array_equal(r_off32, off) and np.array_equal(r_hits, serial_hits), "int32-offset host path differs"
        off_bytes = 4
    del serial_hits

    # ---- count-only e2e: len(find()) per line is all scripts/bed_count_overlapping.py:27-33 needs -- 16 bytes per query
    #      cross PCIe (12 up, 4 down) instead of ~42 ---------------------------------------------------------------------
    h_cnt = _lib.PinnedArray(nq, np.int32)

    def step_host_count():
        check(L.bxg_itree_count(forest.handle, ptr(h_qt.array), ptr(h_qs.array), ptr(h_qe.array), nq, _lib.HOST,
                                ptr(h_cnt.array), C.byref(total)))
    e2e_count = q_all * e2e_steps / time_host(step_host_count, e2e_steps)
    assert np.array_equal(h_cnt.array, cnt.astype(np.int32)) and total.value == hits_total, "count-only host path differs"

    # ---- what the box can copy when every rank copies at once (the ceiling of e2e) ---------------------------------------
    probe = copy_probe(comm)
    e2e_bytes = 12 * nq + off_bytes * (nq + 1) + 4 * hits_total
    e2e_gbs_rank = e2e_bytes / (nq / (e2e_value / world)) / 1e9            # bytes this rank moves / its step time

    # ---- strong scaling: the N=1 workload (10 M queries in total) split over the ranks -------------------------------------
    strong = None
    if world > 1:
        nq_s = max(1, nq // world)
        tot_s = C.c_int64()

        def step_dev_strong():
            check(L.bxg_itree_find(forest.handle, d_qt.ptr, d_qs.ptr, d_qe.ptr, nq_s, _lib.DEVICE, C.byref(tot_s)))

        def step_host_strong():
            check(L.bxg_itree_find_host32(forest.handle, ptr(h_qt.array), ptr(h_qs.array), ptr(h_qe.array), nq_s,
                                          C.byref(p_off), C.byref(p_hits), C.byref(tot_s)))
        q_s_all = int(comm.allreduce_sum_i64(np.array([nq_s]))[0])
        ms_s = DevTimer(comm)(step_dev_strong, args.steps, warm=3)
        strong = {"queries_total": q_s_all, "value": q_s_all / (ms_s * 1e-3), "ms_per_step": ms_s,
                  "e2e": q_s_all * e2e_steps / time_host(step_host_strong, e2e_steps)}
        step_dev()

    extra = {"build_ms": min(build_ms), "hits_per_step": hits_all, "hits_per_query": hits_all / q_all,
             "per_chrom_hits_checksum": int((per_chrom * np.arange(1, 25)).sum()), "parity_spot_check": parity,
             "e2e_serial_copies": e2e_serial, "e2e_int64_offsets": e2e_i64, "e2e_count_only": e2e_count,
             "single_pass_kernel_ms_per_step": single_pass_ms, "three_pass_overlapped_ms_per_step": overlap_ms,
             "sorted_queries_ms_per_step": sorted_ms,
             "copy_probe": probe, "strong_scaling": strong}
    del h_off, h_hits, h_cnt, h_qt, h_qs, h_qe, d_qt, d_qs, d_qe
    forest = None                                          # free the index before the bitmap / score legs

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"

    # ---- configs[2..4]: every rank takes part (chromosomes sharded, counters all-reduced inside the timed step) ----------
    legs = {}
    if not args.no_bitset:
        for name, fn in (("bitset", leg_bitset), ("bed_intersect", leg_bed_intersect), ("aggregate", leg_aggregate)):
            t_leg = time.perf_counter()
            legs[name] = fn(comm, peak, args)
            if rank == 0:
                legs[name]["wall_s_incl_setup_and_checks"] = round(time.perf_counter() - t_leg, 1)
                log(f"{name}: {legs[name].get('ms_per_step', legs[name].get('and_genome', {}).get('ms_per_pass'))} ms, "
                    f"parity_ok={legs[name]['parity_ok']} ({legs[name]['wall_s_incl_setup_and_checks']} s wall)")

    if rank != 0:
        comm.close()
        return

    # ---- roofline of the dominant find kernel ------------------------------------------------------------------------------
    n_items = len(s)
    alg = {
        # per launch, bytes that must move (DESIGN.md "algorithmic bytes"):
        # count pass: read (chrom,start,end) 12 B/query, write (count,lo,hi,mask) 20 B/query, read S,E once 8 B/item
        # (the PM array itself is no longer read: coarse lo)
        # (the library's profiler reports the launch macro's text: PROBE is the search variant, BXB200_FIND_PROBE)
        "k_find<false, PROBE>": 32 * nq + 8 * n_items,
        # fill pass: read lo,mask,offset 20 B/query, read I + write hit 8 B/hit (E is not read again: mask stash)
        "k_find<true, 0>": 20 * nq + 8 * hits_total,
        "k_fill_staged": 20 * nq + 8 * hits_total,         # same bytes; stores staged through shared memory
        # single-pass kernel: read (chrom,start,end) 12 B/query, write offset 8 B/query, S,E once 8 B/item,
        # read I + write hit 8 B/hit
        "k_find_fused<PROBE>": 20 * nq + 8 * n_items + 8 * hits_total,
    }
    kern = {}
    for name, (n_l, tot_ms) in prof.items():
        key = name.replace("(", "").replace(")", "")
        kern[key] = {"launches": n_l, "avg_ms": tot_ms / max(1, n_l)}
    dom = max((k for k in kern if k in alg), key=lambda k: kern[k]["avg_ms"] * kern[k]["launches"])
    dom_ms = kern[dom]["avg_ms"]
    achieved = alg[dom] / (dom_ms * 1e-3) / 1e9
    tj = {}
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except (OSError, ValueError):
        pass
    traffic = tj.get(dom, tj.get(dom.replace("PROBE", os.environ.get("BXB200_FIND_PROBE", "4"))))
    roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg[dom], "avg_launch_ms": dom_ms,
                "note": "find is not HBM-bound (SURVEY 8d): every query is a dependent chain of random 32-byte sector reads over "
                        "an ~85 MB index; frac is reported, not targeted.  The HBM-bound half of the metric is bitset_and_* below; "
                        "c4_* / c5_* are BASELINE configs[3] / [4] (all ranks, NCCL reduce inside the timed step)"}
    step_ms = sum(v["avg_ms"] * v["launches"] for v in kern.values()) / args.steps
    extra["kernels"] = {k: {"avg_ms": round(v["avg_ms"], 4), "per_step": v["launches"] / args.steps,
                            "share": round(v["avg_ms"] * v["launches"] / args.steps / step_ms, 4)} for k, v in kern.items()}

    extra["pcie"] = bench_pcie()
    extra["scalar_api"] = bench_scalar_api(db)
    if not args.no_bitset:
        extra.update(legs)
        extra["score_sources"] = bench_score_sources(peak)

    # ---- CPU baseline (bounded sample, single thread = the reference's only native mode) -----------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        cpu = cpu_baseline(db, qq)
        extra["cpu_reference"] = cpu_reference_extras()

    # ---- BASELINE.json's metric has two halves (queries/s and bitset-AND GB/s against the HBM roof) and two more configs
    #      on N GPUs: flat scalar keys next to the find kernel's entry, at every N (all ranks took part) ---------------------
    if legs:
        bs, c4, c5 = legs["bitset"], legs["bed_intersect"], legs["aggregate"]
        roofline.update({
            "bitset_and_gbs": bs["and_genome"]["gbs"], "bitset_and_frac": bs["and_genome"]["frac"],
            "bitset_and_ms": bs["and_genome"]["ms_per_pass"], "bitset_and_peak_gbs": peak * world,
            "bitset_and_bytes": bs["algorithmic_bytes_per_pass"], "bitset_and_traffic": tj.get("k_binop_batch<0, 0>"),
            "bitset_and_count_gbs": bs["and_count_genome"]["gbs"], "bitset_and_count_frac": bs["and_count_genome"]["frac"],
            "bitset_and_per_pair_gbs": bs["and_per_pair"]["gbs"], "bitset_and_per_pair_frac": bs["and_per_pair"]["frac"],
            "bitset_and_parity_ok": bs["parity_ok"],
            "bitset_and_workload": "a &= b over 24 BinnedBitSet(250000000) pairs, all ranks, one launch per rank",
            "c4_ms": c4["ms_per_step"], "c4_intervals_per_s": c4["intervals_per_s"], "c4_gbs": c4["gbs"], "c4_frac": c4["frac"],
            "c4_parity_ok": c4["parity_ok"], "c4_workload": c4["workload"],
            "c5_ms": c5["ms_per_step"], "c5_windows_per_s": c5["windows_per_s"], "c5_gbs": c5["gbs"], "c5_frac": c5["frac"],
            "c5_parity_ok": c5["parity_ok"], "c5_workload": c5["workload"],
        })
        cr = (extra.get("cpu_reference") or {})
        if "bitset_and" in cr:                               # the metric's "vs Cython CPU" for this half (one thread)
            roofline["bitset_and_cpu_gbs"] = cr["bitset_and"]["gbs"]
            roofline["bitset_and_cpu_cores"] = 1
        for k_cpu, k_out in (("bed_intersect_intervals_per_s", "c4_cpu_intervals_per_s"),
                             ("aggregate_windows_per_s", "c5_cpu_windows_per_s")):
            if k_cpu in cr:
                roofline[k_out] = cr[k_cpu]

    e2e = {"value": e2e_value, "unit": UNIT,
           "h2d_bytes_per_step": int(12 * nq), "d2h_bytes_per_step": int(off_bytes * (nq + 1) + 4 * hits_total),
           "steps": e2e_steps,
           "timing": "host wall clock around bxg_itree_find_host32 (pinned host arrays in, pinned host CSR out with int32 offsets; "
                     "copies overlapped with kernels), max over ranks; extra.e2e_int64_offsets is the int64-offset call",
           # what the box sustains when all ranks copy at once, and how much of it the e2e call uses (per rank)
           "copy_ceiling_gbs_per_rank": probe["bidir_gbs_min"], "copy_ceiling_gbs_aggregate": probe["bidir_gbs_sum"],
           "copy_gbs_achieved_per_rank": e2e_gbs_rank,
           # the call is device-to-host bound: the slowest rank's D2H rate while every rank copies down, and what the call
           # needs of it (its result bytes / the step time)
           "d2h_ceiling_gbs_slowest_rank": probe["d2h_gbs_min"], "d2h_ceiling_gbs_fastest_rank": probe["d2h_gbs_max"],
           "d2h_gbs_achieved_per_rank": (off_bytes * (nq + 1) + 4 * hits_total) / (nq / (e2e_value / world)) / 1e9,
           "frac_of_copy_ceiling": e2e_gbs_rank / probe["bidir_gbs_min"] if probe["bidir_gbs_min"] else None,
           "count_only_value": e2e_count, "count_only_h2d_bytes_per_step": int(12 * nq),
           "count_only_d2h_bytes_per_step": int(4 * nq)}
    if strong:
        e2e["strong_scaling_value"] = strong["value"]
        e2e["strong_scaling_e2e"] = strong["e2e"]
        e2e["strong_scaling_queries_total"] = strong["queries_total"]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic", "config": workload_config(args, world),
        "clocks": clock_summary, "gpu_launches": launches,
        "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu, "extra": extra,
    }
    print(json.dumps(line))
    comm.close()


def copy_probe(comm):
    """Every rank copies 256 MB blocks through pinned memory at the same time: H2D alone, D2H alone, both at once."""
    import ctypes as C

    from bx_python_b200 import _lib
    L = _lib.lib()
    h2d, d2h, bi = C.c_double(), C.c_double(), C.c_double()
    _lib.sync()
    comm.barrier()
    _lib.check(L.bxg_copy_probe(256 << 20, 6, C.byref(h2d), C.byref(d2h), C.byref(bi)))
    vals = np.array([h2d.value, d2h.value, bi.value])
    mx = comm.allreduce_max_f64(vals.copy())
    mn = -comm.allreduce_max_f64(-vals.copy())
    # sums through the int64 reduction (MB/s resolution is plenty)
    sm = comm.allreduce_sum_i64((vals * 1000).astype(np.int64)) / 1000.0
    return {"h2d_gbs_min": float(mn[0]), "h2d_gbs_max": float(mx[0]), "h2d_gbs_sum": float(sm[0]),
            "d2h_gbs_min": float(mn[1]), "d2h_gbs_max": float(mx[1]), "d2h_gbs_sum": float(sm[1]),
            "bidir_gbs_min": float(mn[2]), "bidir_gbs_max": float(mx[2]), "bidir_gbs_sum": float(sm[2]),
            "ranks": comm.world, "bytes_per_copy": 256 << 20,
            "note": "all ranks probe simultaneously after a barrier; bidir = H2D + D2H bytes on two streams / the later finish"}


def spot_check(forest, tid, s, e, qt, qs, qe, total):
    """Ordered hit lists of a query sample vs the oracle restatement; returns a short report."""
    from bx_python_b200 import _lib
    from bx_python_b200._lib import check, ptr
    from oracle import oracle as orc
    L = _lib.lib()
    nq = len(qs)
    off = np.empty(nq + 1, np.int64)
    hits = np.empty(total, np.int32)
    check(L.bxg_itree_fetch(forest.handle, ptr(off), ptr(hits)))
    checked = 0
    for c in (0, len(np.unique(tid)) - 1):
        items = np.nonzero(tid == c)[0]
        o = orc.OracleIntervalTree(s[items], e[items])
        qsel = np.nonzero(qt == c)[0][:20000]
        ooff, ohits = o.find(qs[qsel], qe[qsel])
        ohits = items[ohits].astype(np.int32)
        for j, q in enumerate(qsel):
            if not np.array_equal(hits[off[q]:off[q + 1]], ohits[ooff[j]:ooff[j + 1]]):
                raise AssertionError(f"parity failure: tree {c} query {q}")
        checked += len(qsel)
    return f"{checked} queries bit-identical to oracle (ordered hit lists)"


def cpu_baseline(db, qq):
    from oracle import oracle as orc
    c = 20   # chr21: the smallest chromosome keeps the sample at a few seconds; full database density
    if orc.ref_available():
        r = reference_rate(db, qq, 1.0, 1, repeats=1, chroms=[c])
        return {"value": r["rates"][0], "unit": UNIT, "cores": 1, "kind": "reference",
                "sample": f"chr21 shard at full density: {len(db[c][0])} intervals (build {r['build_s']:.1f} s, not timed), "
                          f"{r['queries']} queries, {r['hits']} hits; unmodified bx.intervals.intersection (oracle/_ref), "
                          f"1 thread (the reference holds the GIL)"}
    t = orc.OracleIntervalTree(db[c][0], db[c][1])
    t0 = time.perf_counter()
    off, _ = t.find(qq[c][0], qq[c][1])
    dt = time.perf_counter() - t0
    return {"value": len(qq[c][0]) / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"chr21 shard: {len(db[c][0])} intervals, {len(qq[c][0])} queries; C restatement (oracle/bx_oracle.c)"}


def cpu_reference_extras():
    """The reference's own CPU path for the other halves of the metric, timed on this box's host cores (one thread: the
    reference holds the GIL) on bounded samples: bitset AND / count over one chr1-length BinnedBitSet pair (249 Mbp,
    built from the C3 law), the set_range / count_range call loop of bed_intersect, and the per-base loop of
    aggregate_scores_in_intervals.  Uses oracle/_ref (the unmodified reference compiled here); returns {} without it.
    No GPU work; never raises (a failure is reported as a string)."""
    try:
        from oracle import oracle as orc
        if not orc.ref_available():
            return {}
        bs, _ = orc.ref_modules()
        size, nranges = C3_SIZE, C3_RANGES                                       # one pair of configs[2]
        (sa, ca), (sb, cb), (ps, pc) = synth.c3_case(size, nranges, 0, nq=20000)
        a, b = bs.BinnedBitSet(size), bs.BinnedBitSet(size)
        t0 = time.perf_counter()
        for s_, c_ in zip(sa.tolist(), ca.tolist()):
            a.set_range(s_, c_)
        t_set = time.perf_counter() - t0
        for s_, c_ in zip(sb.tolist(), cb.tolist()):
            b.set_range(s_, c_)
        t0 = time.perf_counter()
        a.iand(b)
        t_and = time.perf_counter() - t0
        t0 = time.perf_counter()
        total = a.count_range(0, size)
        t_cnt = time.perf_counter() - t0
        t0 = time.perf_counter()
        for s_, c_ in zip(ps.tolist(), pc.tolist()):
            a.count_range(s_, c_)
        t_q = time.perf_counter() - t0
        nbytes = 3 * ((size + 63) // 64) * 8
        out = {"kind": "reference", "cores": 1,
               "sample": f"one BinnedBitSet({size}) pair of configs[2] ({nranges} set_range calls each), unmodified bx.bitset",
               "bitset_and": {"ms": t_and * 1e3, "gbs": nbytes / t_and / 1e9, "algorithmic_bytes": nbytes},
               "count_all": {"ms": t_cnt * 1e3, "gbs": nbytes / 3 / t_cnt / 1e9, "bits_set": int(total)},
               "set_range": {"calls_per_s": len(sa) / t_set}, "count_range": {"calls_per_s": len(ps) / t_q},
               # configs[3] is one set_range per file-2 line plus one count_range per file-1 line
               "bed_intersect_intervals_per_s": 2.0 / (t_set / len(sa) + t_q / len(ps))}
        # aggregate: the script's per-base loop (scripts/aggregate_scores_in_intervals.py:107-124) over a dense float32 list
        rng = np.random.default_rng(5003)
        v = synth.aggregate_scores(rng, 200_000)
        ws = rng.integers(0, len(v) - 50, 20_000)
        we = ws + rng.integers(1, 41, 20_000)
        t0 = time.perf_counter()
        acc = 0
        for a_, b_ in zip(ws.tolist(), we.tolist()):
            tot, cnt = 0, 0
            for i in range(a_, b_):
                sc = v[i]
                if sc and sc == sc:
                    tot += sc
                    cnt += 1
            acc += cnt
        dt = time.perf_counter() - t0
        out["aggregate_windows_per_s"] = len(ws) / dt
        out["aggregate"] = {"windows_per_s": len(ws) / dt, "bases_per_s": float((we - ws).sum()) / dt,
                            "sample": "20000 windows of 1..40 bases over 200000 float32 scores; the script's per-base Python loop without "
                                      "mask, reading a plain numpy array -- a LOWER bound of the reference's cost (its loop goes "
                                      "through BinnedArray.__getitem__ twice per base; that pure-Python class does not travel to "
                                      "the GPU box)"}
        return out
    except Exception as e:                             # noqa: BLE001 -- reported, never fatal for the bench line
        return {"error": f"{type(e).__name__}: {e}"}


# ---------------------------------------------------------------------------------------------------------------------
# configs[2..4] on every rank: chromosomes sharded over the ranks, per-chromosome counters reduced with NCCL inside the
# device-timed step.  Every leg is called by ALL ranks and returns its report on rank 0.
# ---------------------------------------------------------------------------------------------------------------------
C3_SIZE = 250_000_000              # configs[2]: 24 x BinnedBitSet(250 Mbp)
C3_RANGES = 400_000
N_CHROM = 24


def load_full_golden():
    """tests/golden/full_size.json: digests of the compiled reference's answers at full size (make_golden_full.py)."""
    try:
        return json.load(open(os.path.join(ROOT, "tests", "golden", "full_size.json")))
    except (OSError, ValueError):
        return {}


def sha_arrays(*arrays):
    import hashlib
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


class DevTimer:
    """barrier -> CUDA events on the library stream around `reps` calls -> max over ranks, in ms per call."""

    def __init__(self, comm):
        from bx_python_b200 import _lib
        self.comm, self.t, self._lib = comm, _lib.Timer(), _lib

    def __call__(self, fn, reps, warm=1):
        for _ in range(warm):
            fn()
        self._lib.sync()
        self.comm.barrier()
        self.t.start()
        for _ in range(reps):
            fn()
        self.t.stop()
        ms = self.t.elapsed_ms() / reps
        return float(self.comm.allreduce_max_f64(np.array([ms]))[0])


def all_ok(comm, ok):
    """True when every rank's check passed."""
    return int(comm.allreduce_sum_i64(np.array([0 if ok else 1]))[0]) == 0


def leg_bitset(comm, peak, args):
    """configs[2]: a &= b (and the fused a &= b; count) over 24 BinnedBitSet(250 000 000) pairs, 400 k set_range calls per
    operand; the pairs are dealt round-robin to the ranks.  GB/s = 3 x 8 bytes per word of every pair on every
    rank / the slowest rank's device time."""
    import ctypes as C

    from bx_python_b200 import _lib
    from bx_python_b200._lib import check, ptr
    from bx_python_b200.bitset import BinnedBitSet
    L = _lib.lib()
    rank, world = comm.rank, comm.world
    mine = lpt_assign([1.0] * N_CHROM, world)[rank]
    A, B = [], []
    for c in mine:
        (sa, ca), (sb, cb), _ = synth.c3_case(C3_SIZE, C3_RANGES, c, nq=1)
        a, b = BinnedBitSet(C3_SIZE), BinnedBitSet(C3_SIZE)
        a.set_ranges(sa, ca)
        b.set_ranges(sb, cb)
        A.append(a)
        B.append(b)
    words = (C3_SIZE + 63) // 64
    n = len(mine)
    ha = (C.c_void_p * max(n, 1))(*[a._h for a in A])
    hb = (C.c_void_p * max(n, 1))(*[b._h for b in B])
    counts = np.zeros(max(n, 1), np.int64)
    timed = DevTimer(comm)

    def batch(count):
        if n:
            check(L.bxg_bits_binop_batch(0, ha, hb, n, ptr(counts) if count else None))

    def per_pair(count):
        for a, b in zip(A, B):
            check(L.bxg_bits_and_count(a._h, b._h, None) if count else L.bxg_bits_and(a._h, b._h))
    total_bytes = 3 * words * 8 * N_CHROM
    res = {"workload": f"{N_CHROM} x BinnedBitSet({C3_SIZE}) pairs, {C3_RANGES} set_range calls per operand, "
                       f"{N_CHROM // world if N_CHROM % world == 0 else f'{N_CHROM}/{world}'} pairs per rank",
           "algorithmic_bytes_per_pass": total_bytes, "unit": "GB/s", "bound": "hbm",
           "l2_policy": f"each rank streams {3 * words * 8 * max(1, N_CHROM // world) / 1e6:.0f} MB per pass (> 126 MB L2)"}
    for name, fn, count, reps in (("and_genome", batch, False, 10), ("and_count_genome", batch, True, 10),
                                  ("and_per_pair", per_pair, False, 5), ("and_count_per_pair", per_pair, True, 5)):
        ms = timed(lambda: fn(count), reps, warm=2)
        gbs = total_bytes / (ms * 1e-3) / 1e9
        res[name] = {"ms_per_pass": ms, "gbs": gbs, "frac": gbs / (peak * world),
                     "launches_per_pass_per_rank": 1 if fn is batch else n}
    # parity: the fused popcounts of the AND against the compiled reference's answer at full size (golden), all ranks
    gold = load_full_golden().get("c3")
    batch(True)
    ok, how = True, "no golden file: unchecked"
    if gold:
        ok = all(int(counts[k]) == gold[c]["dense"]["count_and"] for k, c in enumerate(mine))
        how = "popcount(a & b) of every pair == compiled reference (tests/golden/full_size.json c3.dense.count_and)"
    res["parity_ok"] = all_ok(comm, ok)
    res["parity"] = how
    if rank != 0:
        return None
    # kernel-only launch durations (CUDA events around each launch) on rank 0
    _lib.profile_enable(True)
    for _ in range(5):
        batch(False)
        batch(True)
        per_pair(False)
    prof = _lib.profile_report()
    _lib.profile_enable(False)
    for kname, (nl, tot) in prof.items():
        per_launch = 3 * words * 8 * (n if "batch" in kname else 1)
        avg_ms = tot / nl
        res.setdefault("kernels_rank0", {})[kname.strip("()")] = {
            "launches": nl, "avg_ms": avg_ms, "gbs": per_launch / (avg_ms * 1e-3) / 1e9,
            "frac": per_launch / (avg_ms * 1e-3) / 1e9 / peak}
    # single-bitmap operations on one 250 Mbp pair (31 MB per bitmap: these FIT the 126 MB L2, so they are L2 figures)
    t = _lib.Timer()

    def timed1(fn, reps=10):
        fn()
        _lib.sync()
        t.start()
        for _ in range(reps):
            fn()
        t.stop()
        return t.elapsed_ms() / reps
    n64, nr = C.c_int64(), C.c_int64()
    ms = timed1(lambda: check(L.bxg_bits_not(A[0]._h)))
    res["invert_one"] = {"ms": ms, "gbs": 2 * words * 8 / (ms * 1e-3) / 1e9, "resident": "L2 (31 MB bitmap)"}
    ms = timed1(lambda: check(L.bxg_bits_count_all(A[0]._h, C.byref(n64))))
    res["count_all_one"] = {"ms": ms, "gbs": words * 8 / (ms * 1e-3) / 1e9, "resident": "L2 (31 MB bitmap)",
                            "note": "includes the 8-byte D2H + stream sync of the scalar result"}

    def runs_pass():
        check(L.bxg_bits_not(B[0]._h))              # invalidates the cached runs so each repetition extracts again
        check(L.bxg_bits_runs_count(B[0]._h, C.byref(nr)))
    ms_pair = timed1(runs_pass, reps=6)
    ms_not = timed1(lambda: check(L.bxg_bits_not(B[0]._h)), reps=6)
    res["runs_one"] = {"ms": ms_pair - ms_not, "runs": nr.value, "resident": "L2 (31 MB bitmap)",
                       "gbs": (2 * words * 8 + 8 * nr.value) / max(1e-6, (ms_pair - ms_not) * 1e-3) / 1e9}
    return res


def c4_shards(f1, f2, world):
    """bed_intersect shards: chromosome weight = bitmap words + intervals of both files (LPT)."""
    w = [float((int(L) + 63) // 64) / 8 + len(f1[c][0]) + len(f2[c][0]) for c, L in enumerate(synth.HG38_LENS)]
    return lpt_assign(w, world)


def flat_file(per_chrom, chroms, seed):
    """(chrom id, start, count) of the selected chromosomes in shuffled (file) order, chromosome ids GLOBAL."""
    which = np.concatenate([np.full(len(per_chrom[c][0]), c, np.int32) for c in chroms] or [np.zeros(0, np.int32)])
    s = np.concatenate([per_chrom[c][0] for c in chroms] or [np.zeros(0, np.int32)])
    cnt = np.concatenate([(per_chrom[c][1] - per_chrom[c][0]).astype(np.int32) for c in chroms] or [np.zeros(0, np.int32)])
    perm = np.random.default_rng(seed).permutation(len(which))
    return which[perm], s[perm], cnt[perm], perm


def leg_bed_intersect(comm, peak, args):
    """configs[3] (scripts/bed_intersect.py:42-53 over lib/bx/bitset_builders.py:17-54): file 2 (50 M intervals) -> one
    BinnedBitSet per chromosome (set_range per line); file 1 (50 M intervals) -> count_range(start, end - start) >= 1 per
    line.  Chromosomes are sharded over the ranks (no exchange); the step ends with ONE NCCL all-reduce of the
    per-chromosome counters [overlapping lines, overlapping bases summed over lines, covered bases] and their D2H copy.
    A step = clear bitmaps, set ranges, build rank tables, count, reduce -- everything on the library stream."""
    import ctypes as C

    from bx_python_b200 import _lib
    from bx_python_b200._lib import check, ptr
    from bx_python_b200.bitset import BinnedBitSet
    from oracle import oracle as orc
    L = _lib.lib()
    rank, world = comm.rank, comm.world
    n = args.c4_n
    f2 = synth.genome_intervals(n, 4002)
    f1 = synth.genome_intervals(n, 4001)
    mine = c4_shards(f1, f2, world)[rank]
    bits = {c: BinnedBitSet(int(synth.HG38_LENS[c])) for c in mine}
    sets = (C.c_void_p * N_CHROM)(*[(bits[c]._h if c in bits else None) for c in range(N_CHROM)])
    w2, s2, c2, _ = flat_file(f2, mine, 40 + rank)
    w1, s1, c1, perm1 = flat_file(f1, mine, 41 + rank)
    n2, n1 = len(w2), len(w1)
    d2 = [_lib.DeviceBuffer(a) for a in (w2, s2, c2)]
    d1 = [_lib.DeviceBuffer(a) for a in (w1, s1, c1)]
    d_out = _lib.DeviceBuffer(np.zeros(max(n1, 1), np.int32))
    d_stats = _lib.DeviceBuffer(np.zeros(3 * N_CHROM, np.int64))
    h_stats = _lib.PinnedArray(3 * N_CHROM, np.int64)
    p_stats = d_stats.ptr.value

    def step():
        check(L.bxg_dev_memset(d_stats.ptr, 0, 8 * 3 * N_CHROM))
        for b in bits.values():
            check(L.bxg_bits_clear(b._h))
        check(L.bxg_bits_set_ranges_multi(sets, N_CHROM, d2[0].ptr, d2[1].ptr, d2[2].ptr, n2, _lib.DEVICE))
        check(L.bxg_bits_count_ranges_multi(sets, N_CHROM, d1[0].ptr, d1[1].ptr, d1[2].ptr, n1, d_out.ptr, 1, _lib.DEVICE))
        check(L.bxg_group_stats_i32(d1[0].ptr, d_out.ptr, n1, N_CHROM, 1, d_stats.ptr, _lib.DEVICE))
        check(L.bxg_bits_count_all_multi(sets, N_CHROM, C.c_void_p(p_stats + 16 * N_CHROM), 1, _lib.DEVICE))
        check(L.bxg_comm_allreduce_i64_dev(d_stats.ptr, 3 * N_CHROM))
        check(L.bxg_memcpy_d2h(ptr(h_stats.array), d_stats.ptr, 8 * 3 * N_CHROM))
        _lib.sync()
    timed = DevTimer(comm)
    check(L.bxg_launch_count_reset())
    step()
    launches = _lib.launch_count()
    ms = timed(step, args.c4_steps, warm=1)
    stats = h_stats.array.copy()
    per = stats[:2 * N_CHROM].reshape(N_CHROM, 2)
    covered = stats[2 * N_CHROM:]
    # ---- parity: (i) this rank's chromosomes, every line's count vs the oracle restatement; (ii) digests + reduced
    #      per-chromosome counters vs the compiled reference at full size (golden file)
    got = np.empty(n1, np.int32)
    if n1:
        check(L.bxg_memcpy_d2h(ptr(got), d_out.ptr, got.nbytes))
        _lib.sync()
    file_order = np.empty(n1, np.int32)
    file_order[perm1] = got                                # back to generated (per-chromosome) order
    gold = load_full_golden().get("c4") if n == 50_000_000 else None
    ok, pos, checked = True, 0, 0
    for c in mine:
        m = len(f1[c][0])
        mine_counts = file_order[pos:pos + m]
        pos += m
        if gold:
            ok &= sha_arrays(mine_counts) == gold[c]["sha256"]
        if (not gold) or c == mine[-1]:                    # the oracle restatement on one chromosome (all, without goldens)
            ob = orc.OracleBinnedBitSet(int(synth.HG38_LENS[c]))
            ob.set_ranges(f2[c][0], f2[c][1] - f2[c][0])
            ok &= bool(np.array_equal(mine_counts, ob.count_ranges(f1[c][0], f1[c][1] - f1[c][0])))
        checked += m
    if gold:                                               # the all-reduced counters, on every rank
        ok &= all(int(per[c, 0]) == gold[c]["overlapping"] and int(per[c, 1]) == gold[c]["sum_counts"] and
                  int(covered[c]) == gold[c]["covered"] for c in range(N_CHROM))
    ok = all_ok(comm, ok)
    # ---- per-kernel device times on rank 0 (separate pass; every rank runs it -- the step contains a collective)
    if rank == 0:
        _lib.profile_enable(True)
    step()
    prof = _lib.profile_report() if rank == 0 else None
    if rank == 0:
        _lib.profile_enable(False)
    words_written = int(comm.allreduce_sum_i64(np.array(
        [sum(int(np.sum((f2[c][1] - 1) // 64 - f2[c][0] // 64 + 1)) for c in mine)]))[0])
    if rank != 0:
        return None
    W = int(sum((int(x) + 63) // 64 for x in synth.HG38_LENS))
    # clear W*8 + set (12 n2 in + 8 per word written) + rank build (W*8 read + W words/4 x 4 B table) +
    # count (12 n1 in + 4 n1 out) + per-chromosome counters (8 n1)
    alg = 8 * W + 12 * n + 8 * words_written + 8 * W + W + 16 * n + 8 * n
    kern = {k.strip("()"): {"launches": v[0], "ms": round(v[1], 4)} for k, v in prof.items()}
    return {"workload": f"bed_intersect {n} x {n} hg38-shaped BED intervals, {N_CHROM} chromosomes sharded over {world} rank(s)",
            "ms_per_step": ms, "intervals_per_s": 2 * n / (ms * 1e-3), "lines_per_s": n / (ms * 1e-3),
            "algorithmic_bytes": alg, "gbs": alg / (ms * 1e-3) / 1e9, "frac": alg / (ms * 1e-3) / 1e9 / (peak * world),
            "launches_per_step_rank0": launches, "steps": args.c4_steps,
            "overlapping_lines": int(per[:, 0].sum()), "overlapping_bases": int(per[:, 1].sum()),
            "covered_bases": int(covered.sum()), "kernels_rank0": kern, "parity_ok": ok,
            "parity": (f"every line's count on every rank: SHA-256 per chromosome == compiled reference (full_size.json c4), "
                       f"all-reduced [overlapping, sum, covered] x {N_CHROM} == reference, + oracle restatement on one chromosome per rank"
                       if gold else "every line's count on every rank == oracle restatement (no full-size golden for this size)"),
            "collective": "one ncclAllReduce(int64 x 72) inside the timed step"}


def leg_aggregate(comm, peak, args):
    """configs[4] (scripts/aggregate_scores_in_intervals.py:107-134): 100 M float32 scores as dense per-chromosome tracks,
    5 M BED windows of 1..40 bases.  Chromosomes (tracks + their windows) are sharded over the ranks; a step = one
    genome-wide aggregate launch over the rank's windows (BED order), the per-chromosome counters
    [windows with a value, counted bases], ONE NCCL all-reduce of them and their D2H copy."""
    import ctypes as C

    from bx_python_b200 import _lib
    from bx_python_b200._lib import check, ptr
    from oracle import oracle as orc
    L = _lib.lib()
    rank, world = comm.rank, comm.world
    tracks = synth.genome_scores(args.c5_scores, args.c5_windows, 5001)
    mine = lpt_assign([len(t[1]) + 20.0 * len(t[2]) for t in tracks], world)[rank]
    handles = {}
    for c in mine:
        origin, v, _, _ = tracks[c]
        h = C.c_void_p()
        check(L.bxg_scores_create(ptr(v), len(v), origin, _lib.HOST, C.byref(h)))
        handles[c] = h
    ht = (C.c_void_p * N_CHROM)(*[(handles[c] if c in handles else None) for c in range(N_CHROM)])
    wt = np.concatenate([np.full(len(tracks[c][2]), c, np.int32) for c in mine] or [np.zeros(0, np.int32)])
    ws = np.concatenate([tracks[c][2] for c in mine] or [np.zeros(0, np.int32)])
    we = np.concatenate([tracks[c][3] for c in mine] or [np.zeros(0, np.int32)])
    nw = len(wt)
    perm = np.random.default_rng(50 + rank).permutation(nw)          # BED order, not grouped by chromosome
    d_w = [_lib.DeviceBuffer(a[perm]) for a in (wt, ws, we)]
    outs = [_lib.DeviceBuffer(np.zeros(max(nw, 1), dt)) for dt in (np.float32, np.float32, np.int32, np.float32, np.float32)]
    d_stats = _lib.DeviceBuffer(np.zeros(2 * N_CHROM, np.int64))
    h_stats = _lib.PinnedArray(2 * N_CHROM, np.int64)

    def step():
        check(L.bxg_dev_memset(d_stats.ptr, 0, 16 * N_CHROM))
        check(L.bxg_aggregate_multi(ht, None, N_CHROM, d_w[0].ptr, d_w[1].ptr, d_w[2].ptr, nw, _lib.DEVICE,
                                    outs[0].ptr, outs[1].ptr, outs[2].ptr, outs[3].ptr, outs[4].ptr))
        check(L.bxg_group_stats_i32(d_w[0].ptr, outs[2].ptr, nw, N_CHROM, 1, d_stats.ptr, _lib.DEVICE))
        check(L.bxg_comm_allreduce_i64_dev(d_stats.ptr, 2 * N_CHROM))
        check(L.bxg_memcpy_d2h(ptr(h_stats.array), d_stats.ptr, 16 * N_CHROM))
        _lib.sync()
    timed = DevTimer(comm)
    check(L.bxg_launch_count_reset())
    step()
    launches = _lib.launch_count()
    ms = timed(step, args.c5_steps, warm=1)
    stats = h_stats.array.copy().reshape(N_CHROM, 2)
    # ---- parity: every window of this rank's chromosomes, bit for bit, vs the oracle restatement; SHA-256 of the
    #      avg/min/max columns per chromosome vs what the reference script itself printed at full size (golden)
    host = [np.empty(max(nw, 1), dt) for dt in (np.float32, np.float32, np.int32, np.float32, np.float32)]
    for h, dbuf in zip(host, outs):
        if nw:
            check(L.bxg_memcpy_d2h(ptr(h), dbuf.ptr, h.nbytes))
    _lib.sync()
    inv = np.empty(nw, np.int64)
    inv[perm] = np.arange(nw)
    gold = load_full_golden().get("c5") if (args.c5_scores, args.c5_windows) == (100_000_000, 5_000_000) else None
    ok, pos = True, 0
    for c in mine:
        origin, v, cws, cwe = tracks[c]
        m = len(cws)
        sel = inv[pos:pos + m]
        pos += m
        dense = np.full(origin + len(v), np.nan, np.float32)
        dense[origin:] = v
        ref = orc.aggregate(dense, cws, cwe)
        for k, name in enumerate(("sum", "avg", "count", "min", "max")):
            ok &= bool(np.array_equal(host[k][sel].view(np.uint32), ref[name].view(np.uint32)))
        if gold:
            ok &= sha_arrays(host[1][sel], host[3][sel], host[4][sel]) == gold[c]["sha256"]
            ok &= int(m - stats[c, 0]) == gold[c]["nan_lines"]
    ok = all_ok(comm, ok)
    bases = int(comm.allreduce_sum_i64(np.array([int(np.sum(we.astype(np.int64) - ws))]))[0])
    nw_all = int(comm.allreduce_sum_i64(np.array([nw]))[0])
    if rank == 0:                                   # (every rank runs the profiled step: it contains a collective)
        _lib.profile_enable(True)
    step()
    prof = _lib.profile_report() if rank == 0 else None
    if rank == 0:
        _lib.profile_enable(False)
    for h in handles.values():
        L.bxg_scores_free(h)
    if rank != 0:
        return None
    alg = 4 * bases + 12 * nw_all + 20 * nw_all + 8 * nw_all      # scores read, windows in, 5 columns out, counters pass
    kern = {k.strip("()"): {"launches": v[0], "ms": round(v[1], 4)} for k, v in prof.items()}
    return {"workload": f"aggregate_scores_in_intervals: {args.c5_scores} float32 scores, {nw_all} windows of 1..40 bases, "
                        f"{N_CHROM} chromosomes sharded over {world} rank(s)",
            "ms_per_step": ms, "windows_per_s": nw_all / (ms * 1e-3), "bases_per_s": bases / (ms * 1e-3),
            "algorithmic_bytes": alg, "gbs": alg / (ms * 1e-3) / 1e9, "frac": alg / (ms * 1e-3) / 1e9 / (peak * world),
            "launches_per_step_rank0": launches, "steps": args.c5_steps,
            "windows_with_value": int(stats[:, 0].sum()), "counted_bases": int(stats[:, 1].sum()),
            "kernels_rank0": kern, "parity_ok": ok,
            "parity": ("every window on every rank bit-identical to the oracle restatement (sum, avg, count, min, max)" +
                       ("; SHA-256 of the avg/min/max columns per chromosome == the reference script's own output (full_size.json c5)"
                        if gold else "")),
            "collective": "one ncclAllReduce(int64 x 48) inside the timed step"}


def bench_scalar_api(db):
    """Latency of the reference-shaped scalar calls (one kernel round trip each): the price of NOT batching."""
    from bx_python_b200.bitset import BinnedBitSet
    from bx_python_b200.intervals import IntervalTree
    s, e = db[20]
    t = IntervalTree()
    t.insert_many(s, e)
    t.find(1000, 2000)
    n = 200
    t0 = time.perf_counter()
    for k in range(n):
        t.find(100000 * k, 100000 * k + 1500)
    find_us = (time.perf_counter() - t0) / n * 1e6
    # the same calls without the Python class around them: ctypes -> bxg_itree_find1 (launch + completion word) only
    import ctypes as C

    from bx_python_b200 import _lib
    fn, h, p = _lib.lib().bxg_itree_find1, t._index._h, C.c_void_p()
    ref = C.byref(p)
    t0 = time.perf_counter()
    for k in range(n):
        fn(h, 0, 100000 * k, 100000 * k + 1500, ref)
    find_raw_us = (time.perf_counter() - t0) / n * 1e6
    b = BinnedBitSet(int(synth.HG38_LENS[20]))
    b.set_ranges(np.arange(0, 40_000_000, 5000), np.full(8000, 700))       # a 700-base run every 5000 bases
    b.count_range(0, 5000)
    t0 = time.perf_counter()
    for k in range(n):
        b.count_range(1000 * k, 1500)
    count_us = (time.perf_counter() - t0) / n * 1e6
    t0 = time.perf_counter()
    for k in range(n):
        b.next_set(1000 * k + 900)                  # the run-extraction idiom: the next run is a few thousand bits away
    next_us = (time.perf_counter() - t0) / n * 1e6
    return {"find_us_per_call": find_us, "find1_ctypes_us_per_call": find_raw_us, "count_range_us_per_call": count_us,
            "next_set_us_per_call": next_us,
            "note": "scalar IntervalTree.find goes to the lingering find server (a resident one-warp kernel polling a request "
                    "line in mapped host memory: no launch per call); BinnedBitSet.count_range / next_set = one launch each, "
                    "completion word in mapped memory; results are written by the kernels into mapped pinned memory (no "
                    "copies); the reference's Cython calls take ~1-4 us -- bulk callers should still use the batched methods"}


def bench_pcie():
    """Pinned-memory copy rates of this box (the ceiling of the e2e figure)."""
    from bx_python_b200 import _lib
    from bx_python_b200._lib import check, ptr
    L = _lib.lib()
    n = 64 << 20
    h = _lib.PinnedArray(n, np.int32)
    h.array[:] = 1
    d = _lib.DeviceBuffer(np.zeros(n, np.int32))
    t = _lib.Timer()
    out = {}
    for name, fn in (("h2d", lambda: check(L.bxg_memcpy_h2d(d.ptr, ptr(h.array), n * 4))),
                     ("d2h", lambda: check(L.bxg_memcpy_d2h(ptr(h.array), d.ptr, n * 4)))):
        fn()
        _lib.sync()
        t.start()
        for _ in range(4):
            fn()
        t.stop()
        out[name + "_gbs"] = 4 * n * 4 / (t.elapsed_ms() * 1e-3) / 1e9
    return out


def bench_score_sources(peak):
    """SURVEY 8f-4 rows, device-timed: the wiggle-load span write (per-GPU share of C5: 12.5 M scored bases), the bigWig
    summary kernel and the overlap join; each with a bit-exact spot check against the oracle."""
    import ctypes as C

    from bx_python_b200 import _lib
    from bx_python_b200._lib import check
    from bx_python_b200.intervals import IntervalForest
    from oracle import oracle as orc
    L = _lib.lib()
    timer = _lib.Timer()
    rng = np.random.default_rng(6001)
    out = {}

    def timed(fn, reps=5):
        fn()
        _lib.sync()
        timer.start()
        for _ in range(reps):
            fn()
        timer.stop()
        return timer.elapsed_ms() / reps

    # ---- span write: 12.5 M bases as (a) single-base records (fixedStep span=1) and (b) 625 k records of 20 bases
    n = 12_500_000
    for tag, ln in (("span1", 1), ("span20", 20)):
        k = n // ln
        st = (np.arange(k, dtype=np.int64) * ln).astype(np.int32)
        en = (st + ln).astype(np.int32)
        v = rng.normal(size=k).astype(np.float32)
        h = C.c_void_p()
        check(L.bxg_scores_alloc(n, 0, float("nan"), C.byref(h)))
        d_s, d_e, d_v = _lib.DeviceBuffer(st), _lib.DeviceBuffer(en), _lib.DeviceBuffer(v)
        ms = timed(lambda: check(L.bxg_scores_set_spans(h, d_s.ptr, d_e.ptr, d_v.ptr, k, _lib.DEVICE)))
        got = np.empty(4096, np.float32)
        check(L.bxg_scores_get_range(h, n - 4096, n, got.ctypes.data_as(C.c_void_p)))
        assert np.array_equal(got.view(np.uint32), np.repeat(v, ln)[-4096:].view(np.uint32)), "span write parity"
        alg = 12 * k + 4 * n
        out[f"set_spans_{tag}"] = {"ms": ms, "records": k, "bases_per_s": n / (ms * 1e-3), "algorithmic_bytes": alg,
                                   "gbs": alg / (ms * 1e-3) / 1e9, "frac": alg / (ms * 1e-3) / 1e9 / peak,
                                   "note": "one pass (write + sortedness check) and the 24-byte D2H round trip of its flags"}
        L.bxg_scores_free(h)
    # overlapping batch -> owner/apply path
    k = 1_000_000
    st = rng.integers(0, n - 64, k).astype(np.int32)
    en = (st + rng.integers(1, 40, k)).astype(np.int32)
    v = rng.normal(size=k).astype(np.float32)
    h = C.c_void_p()
    check(L.bxg_scores_alloc(n, 0, float("nan"), C.byref(h)))
    d_s, d_e, d_v = _lib.DeviceBuffer(st), _lib.DeviceBuffer(en), _lib.DeviceBuffer(v)
    ms = timed(lambda: check(L.bxg_scores_set_spans(h, d_s.ptr, d_e.ptr, d_v.ptr, k, _lib.DEVICE)))
    ref = np.full(n, np.nan, np.float32)
    orc.scores_set_spans(ref, 0, st, en, v)
    got = np.empty(n, np.float32)
    check(L.bxg_scores_get_range(h, 0, n, got.ctypes.data_as(C.c_void_p)))
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), "ordered span write parity"
    out["set_spans_overlapping"] = {"ms": ms, "records": k, "records_per_s": k / (ms * 1e-3),
                                    "parity": "12.5 M cells bit-identical to the sequential loop (oracle)"}
    L.bxg_scores_free(h)

    # ---- bigWig summary: 10 M sorted disjoint intervals into 1000 bins
    k = 10_000_000
    ln = rng.integers(1, 30, k)
    st64 = np.cumsum(ln + rng.integers(0, 5, k)) - ln
    st, en = st64.astype(np.int32), (st64 + ln).astype(np.int32)
    v = rng.normal(size=k).astype(np.float32)
    size = 1000
    d_s, d_e, d_v = _lib.DeviceBuffer(st), _lib.DeviceBuffer(en), _lib.DeviceBuffer(v)
    zeros = np.zeros(size)
    d_out = [_lib.DeviceBuffer(zeros) for _ in range(5)]

    def summ():
        for b in d_out:
            check(L.bxg_memcpy_h2d(b.ptr, zeros.ctypes.data_as(C.c_void_p), zeros.nbytes))
        check(L.bxg_summarize(d_s.ptr, d_e.ptr, d_v.ptr, k, _lib.DEVICE, 0, int(en[-1]), size, *[b.ptr for b in d_out]))
    ms = timed(summ)
    got = np.empty(size)
    check(L.bxg_memcpy_d2h(got.ctypes.data_as(C.c_void_p), d_out[3].ptr, got.nbytes))
    _lib.sync()
    m = 200_000                                       # the first 200 k intervals fill the first bins completely
    nb = int(en[m - 1]) // (int(en[-1]) // size) - 1
    o = orc.summarize(st[:m], en[:m], v[:m], 0, int(en[-1]), size, 0.0, 0.0)
    assert nb > 2 and np.array_equal(got[:nb].view(np.uint64), o["sum_data"][:nb].view(np.uint64)), "summary parity"
    alg = 12 * k + 40 * size
    out["summarize"] = {"ms": ms, "intervals": k, "bins": size, "intervals_per_s": k / (ms * 1e-3),
                        "algorithmic_bytes": alg, "gbs": alg / (ms * 1e-3) / 1e9, "frac": alg / (ms * 1e-3) / 1e9 / peak,
                        "note": "one warp per bin: 32 intervals weighed in parallel, folded in file order by one lane (float64 sums are order-dependent)",
                        "parity": f"first {nb} bins bit-identical to oracle"}

    # ---- join: 1 M left x 1 M right hg38-shaped intervals, mincols 1 and 500
    chroms = list(range(len(synth.HG38_LENS)))
    (tid, s, e), (qt, qs, qe) = (flatten(synth.genome_intervals(1_000_000, 6002), chroms),
                                flatten(synth.genome_intervals(1_000_000, 6003), chroms))
    forest = IntervalForest(len(chroms)).build(tid, s, e)
    d = [_lib.DeviceBuffer(a) for a in (qt, qs, qe, s, e)]
    total = C.c_int64()
    for mincols in (1, 500):
        ms = timed(lambda: check(L.bxg_itree_join(forest.handle, d[0].ptr, d[1].ptr, d[2].ptr, len(qs), d[3].ptr, d[4].ptr,
                                                  mincols, _lib.DEVICE, C.byref(total))), reps=3)
        out[f"join_mincols{mincols}"] = {"ms": ms, "left": len(qs), "right": len(s), "pairs": total.value,
                                         "left_per_s": len(qs) / (ms * 1e-3)}
    poff = np.empty(len(qs) + 1, np.int64)
    items = np.empty(total.value, np.int32)
    vis = np.empty(len(s), np.uint8)
    check(L.bxg_itree_join_fetch(poff.ctypes.data_as(C.c_void_p), items.ctypes.data_as(C.c_void_p), vis.ctypes.data_as(C.c_void_p)))
    c = len(chroms) - 4                                                   # one small chromosome (chr21) against the oracle
    si, qi = np.nonzero(tid == c)[0], np.nonzero(qt == c)[0][:2000]
    ooff, oitems, _ = orc.join(tid[si], s[si], e[si], qt[qi], qs[qi], qe[qi], 500)
    for j, q in enumerate(qi.tolist()):
        assert sorted(items[poff[q]:poff[q + 1]].tolist()) == si[oitems[ooff[j]:ooff[j + 1]]].tolist(), "join parity"
    out["join_mincols500"]["parity"] = f"{len(qi)} left intervals of chromosome #{c}: kept sets identical to oracle"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-db", dest="n_db", type=int, default=N_DB)
    ap.add_argument("--nq", type=int, default=NQ_PER_GPU, help="queries per GPU")
    ap.add_argument("--ref-sample", type=float, default=0.1, help="fraction of each chromosome's queries per reference step")
    ap.add_argument("--ref-workers", type=int, default=0, help="processes for --impl reference (0 = usable cores - 1)")
    ap.add_argument("--no-bitset", action="store_true", help="skip the configs[2..4] legs and the score-source extras")
    ap.add_argument("--c4-n", dest="c4_n", type=int, default=50_000_000, help="intervals per BED file of configs[3]")
    ap.add_argument("--c4-steps", dest="c4_steps", type=int, default=3)
    ap.add_argument("--c5-scores", dest="c5_scores", type=int, default=100_000_000)
    ap.add_argument("--c5-windows", dest="c5_windows", type=int, default=5_000_000)
    ap.add_argument("--c5-steps", dest="c5_steps", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
