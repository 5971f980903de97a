"""Driver for oracle/_ref/ref_bench: the UNMODIFIED reference's operatorfunctions::TensorMultiply timed (and, with the
shared counter-based operator fill, evaluated) on a synthetic big block.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/dmrg_oracle.py header): imported by bench.py's CPU leg and by tests/."""
from __future__ import annotations

import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BENCH = os.path.join(HERE, "_ref", "ref_bench")


def available() -> bool:
    return os.path.exists(REF_BENCH)


def host_cores() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def run(big, terms, indices, cores=None, fills=None, psi=None, reps=1, timeout=1800):
    """Run ref_bench on terms[i] for i in indices (terms = dmrg_oracle.h_terms(big)).
    fills: optional {i: (lseed, lamp, rseed, ramp)} selecting the CUDA library's counter-based operator values;
    psi: optional flat wavefunction (FlattenInto order) -> the summed sigma contribution is returned.
    Returns (seconds, flops, sigma or None)."""
    cores = cores or host_cores()
    tmp = tempfile.mkdtemp(prefix="refbench")
    spec = os.path.join(tmp, "spec.txt")
    with open(spec, "w") as f:
        for blk in (big.left, big.right):
            f.write("%d\n" % len(blk.dims))
            for q, d in zip(blk.q, blk.dims):
                f.write("%d %d %d\n" % (q[0], q[1], d))
        f.write("%d %d\n%d\n" % (big.psi_dq[0], big.psi_dq[1], len(indices)))
        for i in indices:
            lop, rop, scale = terms[i]
            ls, la, rs, ra = fills[i] if fills else (0, 0.0, 0, 0.0)
            f.write("%d %d %d %d  %d %d %d %d  %.17g  %d %.17g %d %.17g\n" % (
                lop.op.dq[0], lop.op.dq[1], int(lop.op.fermion), int(lop.t), rop.op.dq[0], rop.op.dq[1], int(rop.op.fermion), int(rop.t), scale,
                ls, la, rs, ra))
    cmd = [REF_BENCH, spec, str(reps), str(cores)]
    out_path = None
    if psi is not None:
        psi_path, out_path = os.path.join(tmp, "psi.bin"), os.path.join(tmp, "sigma.bin")
        np.ascontiguousarray(psi, dtype=np.float64).tofile(psi_path)
        cmd += [psi_path, out_path]
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS=str(cores))
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout)
    sigma = np.fromfile(out_path) if out_path and os.path.exists(out_path) else None
    for name in os.listdir(tmp):
        os.unlink(os.path.join(tmp, name))
    os.rmdir(tmp)
    for ln in res.stdout.splitlines():
        if ln.startswith("REFBENCH"):
            kv = dict(x.split("=") for x in ln.split()[1:])
            return float(kv["seconds"]), float(kv["flops"]), sigma
    raise RuntimeError("ref_bench failed (rc %d): %s" % (res.returncode, res.stderr[-400:]))
