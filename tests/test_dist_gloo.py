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
