#!/usr/bin/env python
"""One rank of a multi-GPU drop-in sweep (SURVEY.md 8e beyond sigma): launched by torchrun, one process per GPU,

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/run_dropin_multigpu.py <case> [--out DIR]

every rank runs the SAME reference sweep (oracle/_ref/block_gpu) in its own scratch directory; the hooks partition the operator terms of
multiplyH / diagonalH and the noise operators over the ranks (b2d_plan(rank, nranks), cost-weighted ownership) and the library all-reduces
the partial results over NCCL.  Rank 0 prints one JSON line: wall time, per-sweep energies of every rank (they must be identical) and
the difference to the golden sweeps of the unmodified reference."""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_dropin as T  # noqa: E402


def main():
    args = sys.argv[1:]
    out_dir = None
    if "--out" in args:
        i = args.index("--out"); out_dir = args[i + 1]; del args[i:i + 2]
    name = args[0]
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    port = os.environ.get("MASTER_PORT", "0")
    share = os.path.join(tempfile.gettempdir(), "b2d_multigpu_%s_%s" % (name, port))
    os.makedirs(share, exist_ok=True)
    idfile = os.path.join(share, "nccl_id")
    if rank == 0 and os.path.exists(idfile):
        os.unlink(idfile)
    env = {"B2D_NCCL_ID_FILE": idfile}
    t0 = time.time()
    out, golden, stats = T.run_case(name, env, timeout=3000)
    dt = time.time() - t0
    got = T.parse_sweeps(out.stdout)
    res = {"rank": rank, "rc": out.returncode, "wall_s": dt, "energies": [e for _, _, _, e in got]}
    json.dump(res, open(os.path.join(share, "rank%d.json" % rank), "w"))
    if out.returncode != 0:
        sys.stderr.write("rank %d: exit %d\n%s\n" % (rank, out.returncode, out.stderr[-2000:]))
    if rank == 0:
        deadline = time.time() + 600
        ranks = []
        for r in range(world):
            p = os.path.join(share, "rank%d.json" % r)
            while not os.path.exists(p) and time.time() < deadline:
                time.sleep(0.2)
            time.sleep(0.2)
            ranks.append(json.load(open(p)))
        ref = [e for _, _, _, e in golden]
        line = {"case": name, "n_gpus": world, "wall_s": max(r["wall_s"] for r in ranks), "rc": [r["rc"] for r in ranks],
                "sweeps": len(ranks[0]["energies"]),
                "ranks_identical": all(r["energies"] == ranks[0]["energies"] for r in ranks),
                "max_abs_dE_vs_reference": max([abs(a - b) for a, b in zip(ranks[0]["energies"], ref)] + [0.0]) if len(ranks[0]["energies"]) == len(ref) else None,
                "final_energy": ranks[0]["energies"][-1] if ranks[0]["energies"] else None}
        print(json.dumps(line), flush=True)
        if out_dir:
            os.makedirs(out_dir, exist_ok=True)
            open(os.path.join(out_dir, "%s_n%d.json" % (name, world)), "w").write(json.dumps(line) + "\n")
            open(os.path.join(out_dir, "%s_n%d.stats.txt" % (name, world)), "w").write(stats)


if __name__ == "__main__":
    main()
