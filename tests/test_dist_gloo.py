"""
CPU coverage of the N>1 path (world size 2, gloo): chromosome sharding (LPT) and the final reduction of the
per-chromosome counters.  The per-shard compute is done by the oracle here (there is no GPU in this container);
what is tested is the host logic bench.py uses on the GPU box: every chromosome is owned by exactly one rank, the
routed queries cover the workload exactly once, and all-reduce(sum) of the per-rank counters equals the unsharded
answer.  The GPU box runs the same code with backend 'nccl' (bxg_comm_*).
"""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from bx_python_b200 import synth
from bx_python_b200.dist import lpt_assign

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lpt_assign_properties():
    w = synth.HG38_LENS.astype(float)
    for n in (1, 2, 3, 4, 8):
        shards = lpt_assign(w, n)
        flat = sorted(c for s in shards for c in s)
        assert flat == list(range(24)), "every chromosome exactly once"
        loads = [w[s].sum() for s in shards]
        assert max(loads) <= w.sum() / n + w.max(), "LPT bound"
        assert max(loads) / (w.sum() / n) < 1.08
    assert lpt_assign([5, 1, 1], 5)[0] == [0]


WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np
    sys.path.insert(0, {root!r})
    from bx_python_b200 import synth
    from bx_python_b200.dist import Comm, lpt_assign
    from oracle import oracle as orc
    comm = Comm("gloo")
    db = synth.genome_intervals(60000, 2001)
    qq = synth.genome_intervals(40000 * comm.world, 2002)
    shards = lpt_assign([len(q[0]) + 0.25 * len(d[0]) for q, d in zip(qq, db)], comm.world)
    per_chrom = np.zeros(24, np.int64)
    nq = 0
    for c in shards[comm.rank]:
        off, _ = orc.OracleIntervalTree(db[c][0], db[c][1]).find(qq[c][0], qq[c][1])
        per_chrom[c] = off[-1]
        nq += len(qq[c][0])
    comm.barrier()
    total = comm.allreduce_sum_i64(per_chrom.copy())
    q_all = comm.allreduce_sum_i64(np.array([nq]))
    tmax = comm.allreduce_max_f64(np.array([float(comm.rank + 1)]))
    if comm.rank == 0:
        print(json.dumps({{"per_chrom": total.tolist(), "q_all": int(q_all[0]), "tmax": float(tmax[0]),
                          "mine": per_chrom.tolist(), "shards": shards}}))
    comm.close()
""")


def test_two_rank_shard_and_reduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
         "127.0.0.1", "--master-port", str(port), str(script)],
        capture_output=True, text=True, timeout=600, env={**os.environ, "OMP_NUM_THREADS": "1"})
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    res = json.loads(line)
    # unsharded answer
    from oracle import oracle as orc
    db = synth.genome_intervals(60000, 2001)
    qq = synth.genome_intervals(80000, 2002)
    expect = [int(orc.OracleIntervalTree(d[0], d[1]).find(q[0], q[1])[0][-1]) for d, q in zip(db, qq)]
    assert res["per_chrom"] == expect
    assert res["q_all"] == 80000 and res["tmax"] == 2.0
    assert sorted(c for s in res["shards"] for c in s) == list(range(24))
    assert any(v == 0 for v in res["mine"]) and sum(res["mine"]) < sum(expect), "rank 0 holds only its shard"


# ---- configs[3] / [4] on two ranks: the sharding, routing (global chromosome ids, shuffled file order) and the layout of
# ---- the all-reduced counter vectors that bench.py's leg_bed_intersect / leg_aggregate use, with the oracle as compute
WORKER_C45 = textwrap.dedent("""
    import os, sys, json
    import numpy as np
    sys.path.insert(0, {root!r})
    import bench
    from bx_python_b200 import synth
    from bx_python_b200.dist import Comm, lpt_assign
    from oracle import oracle as orc
    comm = Comm("gloo")
    N = 24
    n = 30000
    f2 = synth.genome_intervals(n, 4002, max_len=300)
    f1 = synth.genome_intervals(n, 4001, max_len=300)
    mine = bench.c4_shards(f1, f2, comm.world)[comm.rank]
    w1, s1, c1, perm1 = bench.flat_file(f1, mine, 41 + comm.rank)
    w2, s2, c2, _ = bench.flat_file(f2, mine, 40 + comm.rank)
    stats = np.zeros(3 * N, np.int64)
    lines = 0
    for c in mine:
        ob = orc.OracleBinnedBitSet(int(synth.HG38_LENS[c]))
        sel2 = w2 == c
        ob.set_ranges(s2[sel2], c2[sel2])
        sel1 = w1 == c
        counts = ob.count_ranges(s1[sel1], c1[sel1])
        stats[2 * c] = int((counts >= 1).sum())
        stats[2 * c + 1] = int(counts.sum())
        stats[2 * N + c] = ob.count_range(0, ob.size)
        lines += int(sel1.sum())
    assert lines == len(w1) and set(np.unique(w1)) <= set(mine)
    stats = comm.allreduce_sum_i64(stats)
    ok = bench.all_ok(comm, comm.rank == 0)          # rank 1 reports a failure: every rank must see it
    tracks = synth.genome_scores(200000, 10000, 5001)
    tm = lpt_assign([len(t[1]) + 20.0 * len(t[2]) for t in tracks], comm.world)[comm.rank]
    agg = np.zeros(2 * N, np.int64)
    for c in tm:
        origin, v, ws, we = tracks[c]
        dense = np.full(origin + len(v), np.nan, np.float32)
        dense[origin:] = v
        r = orc.aggregate(dense, ws, we)
        agg[2 * c] = int((r["count"] >= 1).sum())
        agg[2 * c + 1] = int(r["count"].sum())
    agg = comm.allreduce_sum_i64(agg)
    if comm.rank == 0:
        print(json.dumps({{"stats": stats.tolist(), "agg": agg.tolist(), "all_ok": bool(ok), "mine": mine}}))
    comm.close()
""")


def test_two_rank_bed_intersect_and_aggregate_counters(tmp_path):
    import json
    script = tmp_path / "worker_c45.py"
    script.write_text(WORKER_C45.format(root=ROOT))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
         "127.0.0.1", "--master-port", str(port), str(script)],
        capture_output=True, text=True, timeout=600, env={**os.environ, "OMP_NUM_THREADS": "1"})
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    from oracle import oracle as orc
    n, N = 30000, 24
    f2 = synth.genome_intervals(n, 4002, max_len=300)
    f1 = synth.genome_intervals(n, 4001, max_len=300)
    expect = np.zeros(3 * N, np.int64)
    for c in range(N):
        ob = orc.OracleBinnedBitSet(int(synth.HG38_LENS[c]))
        ob.set_ranges(f2[c][0], f2[c][1] - f2[c][0])
        counts = ob.count_ranges(f1[c][0], f1[c][1] - f1[c][0])
        expect[2 * c], expect[2 * c + 1], expect[2 * N + c] = (counts >= 1).sum(), counts.sum(), ob.count_range(0, ob.size)
    assert res["stats"] == expect.tolist()
    assert 0 < len(res["mine"]) < N and res["all_ok"] is False
    tracks = synth.genome_scores(200000, 10000, 5001)
    agg = np.zeros(2 * N, np.int64)
    for c, (origin, v, ws, we) in enumerate(tracks):
        dense = np.full(origin + len(v), np.nan, np.float32)
        dense[origin:] = v
        r = orc.aggregate(dense, ws, we)
        agg[2 * c], agg[2 * c + 1] = (r["count"] >= 1).sum(), r["count"].sum()
    assert res["agg"] == agg.tolist()


def test_unique_id_exchange_without_torch():
    """The NCCL backend's rendezvous: 128 bytes from rank 0 to the other local ranks over an abstract unix socket."""
    import multiprocessing as mp

    from bx_python_b200.dist import exchange_id
    name = "bxb200-test-%d" % os.getpid()
    payload = bytes(range(128))
    ctx = mp.get_context("fork")
    q = ctx.Queue()

    def worker(r):
        q.put((r, exchange_id(payload if r == 0 else b"", r, 4, name=name, timeout=30)))
    procs = [ctx.Process(target=worker, args=(r,)) for r in (3, 1, 2, 0)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=60) for _ in procs)
    for p in procs:
        p.join()
    assert got == {r: payload for r in range(4)}
    src = open(os.path.join(ROOT, "bx_python_b200", "dist.py")).read()
    assert "import torch" not in src.split("class Comm")[0]       # torch only inside the gloo (CPU test) backend of Comm
    nccl_branch = src.split('if self.backend == "nccl":')[1].split("else:")[0]
    assert "torch" not in nccl_branch
