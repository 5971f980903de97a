#!/usr/bin/env python
"""Benchmark of the DMRG sweep hot path on B200: sigma = H.psi (SpinBlock::multiplyH, spinblock.C:722-789) on the
big block of one two-dot block iteration of BASELINE.json's headline configuration
("synthetic random-integral FCIDUMP, 40 orbitals / 40 electrons, M=4000").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

  * one step = one multiplyH over every operator term of the block iteration (ALL ranks' terms together);
  * value   = algorithmic sigma FP64 GFLOP/s = the dgemm flops the reference issues for that multiplyH
              (operatorfunctions.C:515,530) / device time, psi and operators resident in HBM;
  * e2e     = the same through the reference-facing call with HOST buffers (b2d_multiplyH_host: H2D of psi from pinned
              memory, sigma, D2H of the result inside the timed region);
  * N > 1   = one process per GPU (torchrun); operator terms are partitioned by the reference's ownership rule
              (para_array.h:33-42,360-383), psi is replicated, partial sigmas are summed by NCCL all-reduce inside
              the step: the total work is fixed, so "scaling" is "strong";
  * --impl reference = the CPU implementation of the same TensorMultiply calls on the box's host cores on a bounded
              sample of the terms (see cpu_sigma_sample).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


METRIC = "sigma_fp64_gflops"
UNIT = "GFLOP/s"


def workload_name(a):
    return "synthetic 40o/40e-shaped big block: norbs=%d nelec=%d M=%d, block iteration %d|%d sites, two-dot, C1, spin-adapted" % (
        a.norbs, a.nelec, a.M, a.left_sites, a.norbs - a.left_sites)


def factorised_self_check(sb, psi, nterms=4):
    """Full-size parity of the FACTORISED operators against their own dense form: `nterms` operator-pair terms of multiplyH, evenly
    spaced over the term list, through b2d_tensor_multiply - first with both operators as factor lists, then after b2d_materialise_op
    wrote the same two operators out as dense sector blocks (the form whose parity against the REAL reference's TensorMultiply
    tests/test_synthetic.py and round 1's bench pinned at 1.8e-15).  Returns (worst relative difference, #terms)."""
    lops, rops, flags, scales = sb.terms(all_ranks=False)[:4]
    pick = [int(i) for i in np.unique(np.linspace(2, len(lops) - 1, nterms).astype(int))]
    sb.reserve(4)
    sb.upload(0, psi)
    worst = 0.0
    for i in pick:
        lo, ro, fl, sc = int(lops[i]), int(rops[i]), int(flags[i]), float(scales[i])
        sb.clear(2)
        sb.tensor_multiply_slots(lo, ro, bool(fl & 1), bool(fl & 2), 0, sc, 0, 2)
        fac = sb.download(2)
        sb._ck(sb.lib.b2d_materialise_op(sb._ctx, 0, lo))
        sb._ck(sb.lib.b2d_materialise_op(sb._ctx, 1, ro))
        sb.clear(2)
        sb.tensor_multiply_slots(lo, ro, bool(fl & 1), bool(fl & 2), 0, sc, 0, 2)
        den = sb.download(2)
        worst = max(worst, float(np.linalg.norm(fac - den) / max(np.linalg.norm(den), 1e-300)))
    return worst, len(pick)


# ---------------------------------------------------------------------------------------------------------------------
# CPU leg: the reference algorithm (oracle restatement, numpy/OpenBLAS dgemm) on a bounded sample of the terms
# ---------------------------------------------------------------------------------------------------------------------
class _LazyBlocks(dict):
    """Operator sector blocks generated on demand (values do not matter for timing; shapes and sparsity do)."""

    def __init__(self, dims, seed):
        super().__init__()
        self.dims, self.rng = dims, np.random.default_rng(seed)

    def __missing__(self, key):
        a = self.rng.standard_normal((int(self.dims[key[0]]), int(self.dims[key[1]])))
        self[key] = a
        return a


def cpu_sigma_sample(a, budget_s=8.0, max_terms=64):
    """Time oracle.dmrg_oracle.tensor_multiply (the restatement of operatorfunctions.C:485-537 whose inner products are
    numpy/OpenBLAS dgemm calls on all host cores) on an evenly spaced sample of the multiplyH term list of the SAME big
    block.  Returns (GFLOP/s, cores, description, seconds, flops)."""
    from block_b200 import synthetic as S
    from oracle import dmrg_oracle as O

    nl = a.left_sites
    filling = a.nelec / a.norbs
    L = S.add_dot(S.renormalised_sectors(nl - 1, filling * (nl - 1), a.M))
    R = S.add_dot(S.renormalised_sectors(a.norbs - nl - 1, filling * (a.norbs - nl - 1), a.M))
    L = {k: d for k, d in L.items() if (a.nelec - k[0], k[1]) in R}
    R = {k: d for k, d in R.items() if (a.nelec - k[0], k[1]) in L}
    lsites, rsites = list(range(nl)), list(range(nl, a.norbs))
    blocks = []
    for spec in (S.make_block(L, lsites, rsites, True), S.make_block(R, rsites, lsites, False)):
        blk = O.Block(q=spec.q.astype(np.int64), dims=spec.dims.astype(np.int64), sites=spec.sites, loop=spec.loop)
        for k, op in enumerate(spec.ops):
            blk.ops.append(O.Op(optype=op.optype, orbs=op.orbs, comp=op.comp, dq=op.dq, fermion=op.fermion, allowed=op.allowed.astype(bool),
                                blocks=_LazyBlocks(spec.dims, 1000 * len(blocks) + k)))
        blocks.append(blk)
    big = O.Big(left=blocks[0], right=blocks[1], psi_dq=(a.nelec, 0, 0))
    terms = O.h_terms(big)
    rng = np.random.default_rng(1)
    c = big.unflatten(rng.standard_normal(big.size))
    order = [int(i) for i in np.unique(np.linspace(0, len(terms) - 1, max_terms).astype(int))]
    # interleave so that a truncated sample still covers every term family
    order = order[::4] + order[1::4] + order[2::4] + order[3::4]
    flops, secs, used = [0.0], 0.0, 0
    for i in order:
        lop, rop, scale = terms[i]
        for view in (lop, rop):           # materialise the operator blocks outside the timed region
            for (p, q) in zip(*np.nonzero(view.op.allowed)):
                view.op.blocks[(int(p), int(q))]
        v = big.zeros()
        t0 = time.perf_counter()
        O.tensor_multiply(big, lop, rop, c, v, 0, scale, flops)
        secs += time.perf_counter() - t0
        used += 1
        for view in (lop, rop):
            view.op.blocks.clear()
        if secs > budget_s:
            break
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    desc = "%d of %d TensorMultiply terms of the same multiplyH (evenly spaced over the term list), numpy/OpenBLAS dgemm, %.1f s" % (used, len(terms), secs)
    return flops[0] / secs / 1e9, cores, desc, secs, flops[0]


def _oracle_big(a):
    """The synthetic big block as oracle objects (operator blocks lazily random) and its multiplyH term list."""
    from block_b200 import synthetic as S
    from oracle import dmrg_oracle as O

    nl = a.left_sites
    filling = a.nelec / a.norbs
    L = S.add_dot(S.renormalised_sectors(nl - 1, filling * (nl - 1), a.M))
    R = S.add_dot(S.renormalised_sectors(a.norbs - nl - 1, filling * (a.norbs - nl - 1), a.M))
    L = {k: d for k, d in L.items() if (a.nelec - k[0], k[1]) in R}
    R = {k: d for k, d in R.items() if (a.nelec - k[0], k[1]) in L}
    lsites, rsites = list(range(nl)), list(range(nl, a.norbs))
    blocks = []
    for spec in (S.make_block(L, lsites, rsites, True), S.make_block(R, rsites, lsites, False)):
        blk = O.Block(q=spec.q.astype(np.int64), dims=spec.dims.astype(np.int64), sites=spec.sites, loop=spec.loop)
        for k, op in enumerate(spec.ops):
            blk.ops.append(O.Op(optype=op.optype, orbs=op.orbs, comp=op.comp, dq=op.dq, fermion=op.fermion, allowed=op.allowed.astype(bool),
                                blocks=_LazyBlocks(spec.dims, 1000 * len(blocks) + k)))
        blocks.append(blk)
    big = O.Big(left=blocks[0], right=blocks[1], psi_dq=(a.nelec, 0, 0))
    return big, O.h_terms(big)


def _sample_order(nterms, max_terms):
    order = [int(i) for i in np.unique(np.linspace(0, nterms - 1, max_terms).astype(int))]
    return order[::4] + order[1::4] + order[2::4] + order[3::4]   # a truncated sample still covers every term family


def cpu_sigma_reference(a, budget_s=10.0, max_terms=192, state={}):
    """Time the REAL reference's operatorfunctions::TensorMultiply (oracle/_ref/ref_bench: the unmodified reference
    objects compiled by oracle/Makefile, OpenMP over operator terms with single-threaded OpenBLAS dgemm inside, exactly how
    multiplyH parallelises: operatorloops.h:87-97) on an evenly spaced sample of the SAME multiplyH's term list.
    Returns (GFLOP/s, cores, description, seconds, flops)."""
    from oracle import refbench
    cores = refbench.host_cores()
    if "terms" not in state:
        state["big"], state["terms"] = _oracle_big(a)
    big, terms = state["big"], state["terms"]
    order = _sample_order(len(terms), max_terms)
    if "rate" not in state:                       # calibrate on one term per core
        s0, f0, _ = refbench.run(big, terms, order[:max(cores, 2)], cores)
        state["rate"], state["per_term"] = f0 / s0, f0 / max(cores, 2)
    n = int(min(len(order), max(cores, budget_s * state["rate"] / state["per_term"])))
    n = max(cores, n - n % cores) if n >= cores else n
    secs, flops, _ = refbench.run(big, terms, order[:n], cores)
    desc = ("%d of %d TensorMultiply terms of the same multiplyH (evenly spaced over the term list) through the unmodified reference's "
            "operatorfunctions::TensorMultiply, OpenMP over terms x %d threads, OpenBLAS 0.3.15 dgemm single-threaded inside, %.1f s" % (n, len(terms), cores, secs))
    return flops / secs / 1e9, cores, desc, secs, flops


def reference_parity(sb, a, psi, nterms=6):
    """Full-size parity against the REAL reference: the sum of `nterms` sampled operator-pair terms of multiplyH computed
    by the GPU (b2d_tensor_multiply through the C ABI) and by the unmodified reference's TensorMultiply (ref_bench) on
    identical operator values (shared counter-based stream) and the same psi.  Returns (relative error, #terms)."""
    from block_b200 import synthetic as S
    from oracle import refbench
    big, terms = _oracle_big(a)
    pair_terms = [i for i in range(len(terms)) if terms[i][0].op.optype not in (S.HAM, S.OVERLAP) and terms[i][1].op.optype not in (S.HAM, S.OVERLAP)]
    pick = [pair_terms[int(j)] for j in np.unique(np.linspace(0, len(pair_terms) - 1, nterms).astype(int))]
    fills, calls = {}, []
    for i in pick:
        lop, rop, scale = terms[i]
        kl = next(k for k, o in enumerate(big.left.ops) if o is lop.op)
        kr = next(k for k, o in enumerate(big.right.ops) if o is rop.op)
        ls, la = S.fill_params(big.left.dims, 0, kl, lop.op.optype, sb.fill_seed)
        rs, ra = S.fill_params(big.right.dims, 1, kr, rop.op.optype, sb.fill_seed)
        fills[i] = (ls, la, rs, ra)
        calls.append((sb.op_ids[0][kl], sb.op_ids[1][kr], lop.t, rop.t, scale))
    _, _, ref = refbench.run(big, terms, pick, fills=fills, psi=psi)
    sb.reserve(3)
    sb.upload(0, psi)
    sb.clear(1)
    for (lo, ro, lt, rt, scale) in calls:
        sb.tensor_multiply_slots(lo, ro, lt, rt, 0, scale, 0, 1)
    got = sb.download(1)
    return float(np.linalg.norm(got - ref) / np.linalg.norm(ref)), len(pick)


def cpu_leg(a, budget_s):
    from oracle import refbench
    if refbench.available():
        return ("reference",) + cpu_sigma_reference(a, budget_s)
    return ("port",) + cpu_sigma_sample(a, budget_s)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # a step = a bounded sample of the same multiplyH: --ref-step-s seconds of the box's host cores, shortened so that the whole
    # --steps K --warmup W run stays within --ref-total-s (default 360 s: 16 s per step, >= 128 terms on 16 cores, at the driver's K = 20, W = 5)
    step_s = min(a.ref_step_s, a.ref_total_s / max(a.steps + 0.5 * a.warmup, 1.0))
    for _ in range(a.warmup):
        cpu_leg(a, step_s / 2)
    t, fl, last = 0.0, 0.0, None
    for _ in range(a.steps):
        last = cpu_leg(a, step_s)
        t += last[4]
        fl += last[5]
    value = fl / t / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * t / max(a.steps, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": workload_name(a), "sample": last[3]},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": last[2], "kind": last[0], "sample": last[3]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ---------------------------------------------------------------------------------------------------------------------
# GPU leg
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device, self.proc, self.path = device, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for ln in open(self.path):
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
                except ValueError:
                    continue
                for nm, val in zip(names, f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), power_w_max=float(max(power)), reasons=sorted(reasons), samples=len(sm))
        return out


def block_iteration_leg(sb, a, psi, sigma_ms):
    """The rest of one RenormaliseFrom + transform_operators at the same size (renormalise.C:39-133, save_load_block.C:267):
    diag(H), a Davidson solve capped at a.davidson_iters iterations (random operators do not converge like a molecule's),
    density matrix, per-sector eigen-decomposition, state selection, operator rotation.  Times are wall-clock around the
    synchronous C calls.  Not part of `value`: it shows that nothing beside sigma dominates a block iteration."""
    import ctypes as C
    from block_b200 import _lib
    out = {}
    sb.set_option("max_davidson_iter", a.davidson_iters)
    sb.reserve(3)
    sb.upload(0, psi)

    def timed(fn):
        sb.sync()
        t0 = time.perf_counter()
        r = fn()
        sb.sync()
        return r, (time.perf_counter() - t0) * 1e3

    _, out["diag_ms"] = timed(lambda: sb.lib.b2d_diagonal(sb._ctx, 1))
    ev, nm, res = np.zeros(1), C.c_int(0), C.c_double(0.0)
    rc, out["davidson_ms"] = timed(lambda: sb.lib.b2d_davidson(sb._ctx, 1, 0, 1, 1e-14, 2, 20, ev.ctypes.data_as(_lib.c_f64p), C.byref(nm), C.byref(res)))
    out["davidson_h_applications"] = nm.value
    out["davidson_level1_ms_per_iteration"] = (out["davidson_ms"] - nm.value * sigma_ms) / max(nm.value, 1)
    out["davidson_residual_norm2"] = res.value
    w = np.ones(1)
    rc, out["density_ms"] = timed(lambda: sb.lib.b2d_make_density(sb._ctx, 1, 0, w.ctypes.data_as(_lib.c_f64p)))
    if rc:
        out["error"] = sb.lib.b2d_last_error(sb._ctx).decode()
        return out
    rc, out["eigen_ms"] = timed(lambda: sb.lib.b2d_diagonalise_dm(sb._ctx, None))
    if rc:
        out["error"] = sb.lib.b2d_last_error(sb._ctx).decode()
        return out
    kept = np.zeros(len(sb.left.dims), np.int32)
    err = C.c_double(0.0)
    rc, out["select_ms"] = timed(lambda: sb.lib.b2d_select_states(sb._ctx, a.M, kept.ctypes.data_as(_lib.c_i32p), C.byref(err)))
    out["kept_states"] = int(kept.sum())
    out["discarded_weight"] = err.value
    rc, out["rotate_ms"] = timed(lambda: sb.lib.b2d_transform_operators(sb._ctx))
    if rc:
        out["error"] = sb.lib.b2d_last_error(sb._ctx).decode()
    out["non_sigma_ms"] = out["diag_ms"] + out["davidson_level1_ms_per_iteration"] * nm.value + out["density_ms"] + out["eigen_ms"] + out["select_ms"] + out.get("rotate_ms", 0.0)
    out["sigma_share_of_block_iteration"] = nm.value * sigma_ms / (nm.value * sigma_ms + out["non_sigma_ms"])
    return out


def sweep_leg(a):
    """BASELINE.json's other headline figure, "two-site sweep wall time ... energy delta vs ref": the UNMODIFIED reference sweep
    (oracle/_ref/block.spin_adapted, all host threads) next to the same reference sweep with its hot path re-routed to this library
    (oracle/_ref/block_gpu, tests/dropin/block_gpu_hooks.cpp) on one real FCIDUMP / dmrg.conf case of tests/golden/dropin_cases.npz
    (default synthetic_16o_M300: random-integral FCIDUMP, 16 orbitals, M = 150 -> 300, Davidson tolerance 1e-12, four sweeps, golden
    sweeps committed).  Both run here, back to back, on this box; energies are compared sweep by sweep, with each other and with the
    golden sweeps.  Block construction is inside both times.  "gpu_dropin" is the default drop-in (enlarged-block operators
    materialised on the device for small blocks and FACTORISED for large ones - the binding decides per block iteration -, guess
    wavefunctions transformed on the device, renormalised blocks cached on the device); "gpu_dropin_factorised" forces the factorised
    form for every block iteration."""
    import re
    gpu_bin = os.path.join(ROOT, "oracle", "_ref", "block_gpu")
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "block.spin_adapted")
    cases = os.path.join(ROOT, "tests", "golden", "dropin_cases.npz")
    if not (os.path.exists(gpu_bin) and os.path.exists(ref_bin) and os.path.exists(cases)):
        return {"unavailable": "oracle/_ref/block_gpu or block.spin_adapted not built"}
    z = np.load(cases)
    name = a.sweep_case
    threads = os.cpu_count() or 1
    pat = re.compile(r"M = (\d+)\s+state = (\d+)\s+Largest Discarded Weight = (\S+)\s+Sweep Energy = (\S+)")
    out = {"case": name, "host_threads": threads}
    energies = {}
    # untimed warm-up of the drop-in binary on the reference's smallest case (loads the library, creates a CUDA context once: the first
    # CUDA process on a fresh box pays seconds that belong to neither arm)
    try:
        warm = "c2_d2h_M50"
        if warm + "/conf" in z.files:
            wdir = tempfile.mkdtemp(prefix="sweep_warm_")
            for f in z[warm + "/files"]:
                open(os.path.join(wdir, str(f)), "wb").write(z["%s/file/%s" % (warm, f)].tobytes())
            open(os.path.join(wdir, "dmrg.conf"), "wb").write(z[warm + "/conf"].tobytes())
            subprocess.run([gpu_bin, "dmrg.conf"], cwd=wdir, env=dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1"), capture_output=True, timeout=300)
    except Exception:   # noqa: BLE001
        pass
    for tag, exe in (("reference_cpu", ref_bin), ("gpu_dropin", gpu_bin), ("gpu_dropin_factorised", gpu_bin)):
        work = tempfile.mkdtemp(prefix="sweep_%s_" % tag)
        for f in z[name + "/files"]:
            open(os.path.join(work, str(f)), "wb").write(z["%s/file/%s" % (name, f)].tobytes())
        conf = z[name + "/conf"].tobytes().decode() + "threads_per_node %d\n" % threads
        open(os.path.join(work, "dmrg.conf"), "w").write(conf)
        env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS=str(threads), B2D_DROPIN_STATS=os.path.join(work, "stats.txt"))
        if tag == "gpu_dropin_factorised":
            env.update(B2D_DROPIN_OPTIONS="factorised=1")
        # line-buffered stdout (stdbuf), every "Sweep Energy" line stamped as it arrives: wall time PER SWEEP - the first one is the
        # warm-up sweep (CSF-built guess environments, SURVEY N4: reference host code on both arms), the others are regular sweeps
        t0 = time.perf_counter()
        stamps, lines_out = [], []
        try:
            p = subprocess.Popen(["stdbuf", "-oL", exe, "dmrg.conf"], cwd=work, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            for ln in p.stdout:
                lines_out.append(ln)
                if "Sweep Energy" in ln:
                    stamps.append(time.perf_counter() - t0)
                if time.perf_counter() - t0 > 900:
                    p.kill()
                    break
            err_text = p.stderr.read()
            p.wait(timeout=60)
        except Exception as ex:   # noqa: BLE001
            out[tag] = {"failed": repr(ex)[:300]}
            continue
        dt = time.perf_counter() - t0
        stdout_text = "".join(lines_out)
        if p.returncode != 0:
            out[tag] = {"failed": (err_text or stdout_text)[-300:]}
            continue
        e = [float(m.group(4)) for m in pat.finditer(stdout_text)]
        energies[tag] = e
        nroots = max(1, len({m.group(2) for m in pat.finditer(stdout_text)}))
        ends = stamps[nroots - 1::nroots]                                   # one stamp per sweep (the last root's line)
        per_sweep = [ends[0]] + [b - a for a, b in zip(ends[:-1], ends[1:])] if ends else []
        out[tag] = {"wall_s": dt, "sweep_lines": len(e), "final_energy": e[-1] if e else None, "wall_s_per_sweep": per_sweep,
                    "warmup_sweep_s": per_sweep[0] if per_sweep else None, "regular_sweeps_s": sum(per_sweep[1:]) if per_sweep else None}
        if tag.startswith("gpu_dropin") and os.path.exists(os.path.join(work, "stats.txt")):
            tot = {}
            for l in open(os.path.join(work, "stats.txt")):
                for k, v in re.findall(r"(\w+)=([-\d.e+]+)", l):
                    tot[k] = tot.get(k, 0.0) + float(v)
            out[tag]["hot_path_s"] = {k: tot.get(k, 0.0) for k in ("host_op_build_s", "upload_s", "guess_s", "diag_s", "davidson_s", "density_s", "eig_s", "rotate_s")}
            out[tag]["n_multiply"] = int(tot.get("n_multiply", 0))
            out[tag]["kernel_launches"] = int(tot.get("launches", 0))
            out[tag]["blocks_taken_from_device_cache"] = int(tot.get("cache_uses", 0))
    if "reference_cpu" in energies and "gpu_dropin_factorised" in energies and len(energies["reference_cpu"]) == len(energies["gpu_dropin_factorised"]):
        out["gpu_dropin_factorised"]["max_abs_dE_per_sweep"] = max(abs(x - y) for x, y in zip(energies["reference_cpu"], energies["gpu_dropin_factorised"]))
    try:
        out["speedup_whole_run"] = out["reference_cpu"]["wall_s"] / out["gpu_dropin"]["wall_s"]
        out["speedup_regular_sweeps"] = out["reference_cpu"]["regular_sweeps_s"] / out["gpu_dropin"]["regular_sweeps_s"]
        out["speedup_warmup_sweep"] = out["reference_cpu"]["warmup_sweep_s"] / out["gpu_dropin"]["warmup_sweep_s"]
    except Exception:   # noqa: BLE001
        pass
    if "reference_cpu" in energies and "gpu_dropin" in energies and len(energies["reference_cpu"]) == len(energies["gpu_dropin"]):
        out["max_abs_dE_per_sweep"] = max(abs(x - y) for x, y in zip(energies["reference_cpu"], energies["gpu_dropin"]))
        golden = [float(m.group(4)) for m in pat.finditer(z[name + "/sweeps"].tobytes().decode())] if name + "/sweeps" in z.files else []
        if golden and len(golden) == len(energies["gpu_dropin"]):
            out["max_abs_dE_vs_golden"] = max(abs(x - y) for x, y in zip(golden, energies["gpu_dropin"]))
    return out


def next_rows_leg(a):
    """Rows of SURVEY.md 8f behind the C ABI at the benchmark's size - the guess-wavefunction transform (N1) and, under
    "operator_construction", the construction of the enlarged block's operators (N2: kron_scatter_kernel against the HBM roofline) - in a
    process of its own: a failure there is reported under "error" and never touches the measurements above."""
    cmd = [sys.executable, os.path.join(ROOT, "scripts", "bench_next_rows.py"), "--norbs", str(a.norbs), "--nelec", str(a.nelec), "--M", str(a.M),
           "--left-sites", str(a.left_sites)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"error": (r.stderr or r.stdout)[-400:]}
        return json.loads(lines[-1])
    except Exception as e:   # noqa: BLE001
        return {"error": repr(e)[:400]}


def run_ours(a):
    import torch
    import torch.distributed as dist

    from block_b200 import synthetic as S
    from block_b200.hotpath import SpinBlock

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU leg)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t_setup = time.time()
    options = dict({"workspace_mb": a.workspace_mb, "slice_iters": a.slice_iters}, **{kv.split("=")[0]: float(kv.split("=")[1]) for kv in a.opt})
    if world > 1 and "balance_terms" not in options:
        options["balance_terms"] = 1.0   # cost-weighted ownership of the operator terms (SURVEY 8e allows it: the sum over ranks is the same multiplyH)
    if a.mode == "factorised":
        # the blocks a sweep holds before a block iteration: renormalised M-state blocks + one-site dots; every operator of the two
        # enlarged blocks is built on the device as a list of scaled sub-blocks of the renormalised operators (never materialised)
        case = S.make_product_case(norbs=a.norbs, nelec=a.nelec, M=a.M, left_sites=a.left_sites)
        sb = S.make_big_block_from_products(case, device=local, options=options, factorised=True, rank=rank, nranks=world)
        sb.fill_seed = case["seed"]
    else:
        sb = S.make_big_block(norbs=a.norbs, nelec=a.nelec, M=a.M, left_sites=a.left_sites, device=local, rank=rank, nranks=world, options=options)
    if world > 1:   # the partial sigmas are summed by the library's own NCCL communicator (dlopen'ed libnccl of the torch wheel)
        ident = [SpinBlock.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        sb.attach_communicator(ident[0], rank, world)
    stats = sb.plan_stats()
    flops_alg = sb.sigma_flops(all_ranks=True)          # the reference's dgemm flops for the whole multiplyH
    flops_mine = sb.sigma_flops(all_ranks=False)
    W = sb.size
    rng = np.random.default_rng(5)
    psi_host = torch.from_numpy(rng.standard_normal(W)).pin_memory()
    sig_host = torch.empty(W, dtype=torch.float64).pin_memory()
    psi = psi_host.numpy()
    psi /= np.linalg.norm(psi)
    sb.reserve(3)
    sb.upload(0, psi)
    t_setup = time.time() - t_setup
    stream = torch.cuda.ExternalStream(sb.lib.b2d_stream(sb._ctx), device=torch.device("cuda", local))

    def barrier():
        sb.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        e1.synchronize()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    if a.profile_mode:
        sb.sigma(0, 1)
        sb.sync()
        for _ in range(a.steps):
            sb.sigma(0, 1)
        sb.sync()
        sb.measure_level1(reps=2)     # the HBM-bound Davidson kernels, for the same ncu launch list
        sb.sync()
        print("profile-mode: %d launches per sigma" % int(stats["launches_per_sigma"]), flush=True)
        sb.close()
        return

    # parity at full size (the numpy oracle cannot run this size): (1) linearity H(2x - 3y) = 2Hx - 3Hy,
    # (2) a sample of operator-pair terms against the REAL reference's TensorMultiply on identical inputs
    y = rng.standard_normal(W); y /= np.linalg.norm(y)
    sb.upload(2, y)
    sb.sigma(0, 1)
    hx = sb.download(1)
    sb.sigma(2, 1)
    hy = sb.download(1)
    sb.upload(2, 2.0 * psi - 3.0 * y)
    sb.sigma(2, 1)
    lin_err = float(np.linalg.norm(sb.download(1) - (2.0 * hx - 3.0 * hy)) / np.linalg.norm(2.0 * hx - 3.0 * hy))
    # sigma_norm / sigma_probe are pure functions of the (seeded) workload: they must agree across --gpus 1/2/4/8 runs
    parity = {"linearity_rel": lin_err, "sigma_norm": float(np.linalg.norm(hx)), "sigma_probe": float(np.dot(hx, np.cos(np.arange(W))))}
    sb.sigma(0, 1)
    parity["bit_reproducible"] = bool(np.array_equal(sb.download(1), hx))
    if world == 1 and not a.no_cpu and a.mode == "materialised":
        from oracle import refbench
        if refbench.available():
            err, nt = reference_parity(sb, a, psi)
            parity["vs_reference_TensorMultiply_rel"] = err
            parity["vs_reference_terms"] = nt
        sb.upload(0, psi)

    for _ in range(a.warmup):
        sb.sigma(0, 1)
    clocks = ClockSampler(local)
    clocks.start()
    l0 = sb.kernel_launches()
    ms = timed(lambda: sb.sigma(0, 1), a.steps)
    launches = sb.kernel_launches() - l0
    clk = clocks.stop()
    ms_step = ms / a.steps
    value = flops_alg / (ms_step * 1e-3) / 1e9

    # end to end through the reference-facing call with host buffers (pinned), copies inside the timed region
    sigp = sig_host.numpy()
    for _ in range(min(a.warmup, 2)):
        sb.multiplyH_into(psi, sigp)
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        sb.multiplyH_into(psi, sigp)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = flops_alg / (float(e2e_s.item()) / a.steps) / 1e9
    parity["e2e_vs_resident_rel"] = float(np.linalg.norm(sigp - hx) / np.linalg.norm(hx))

    line = None
    if rank == 0 or world == 1:
        # roofline of the dominant kernel (128x128 DMMA tile class of the grouped contraction), timed live with events
        prof = sb.sigma_profile(0, 1)
        dmma_live, dfma = sb.measure_fp64_peak()
        # the roofline denominator is PINNED: profiles/r02_fp64_yardstick.json (tracked; the same register-loop yardstick measured on a B200 of this
        # pool with its clock record); the live measurement of this run is reported next to it
        dmma, pinned = dmma_live, False
        try:
            dmma = float(json.load(open(os.path.join(ROOT, "profiles", "r02_fp64_yardstick.json")))["fp64_dmma_tflops"])
            pinned = True
        except Exception:   # noqa: BLE001
            pass
        # the dominant kernel = the tile class of the grouped contraction with the largest share of the step's time
        cls_ms = {c: sum(prof[(st, c)][0] for st in range(2)) for c in range(10)}
        kc = max(cls_ms, key=cls_ms.get)
        k_ms = sum(prof[(st, kc)][0] for st in range(2))
        k_fl = sum(prof[(st, kc)][1] for st in range(2))
        k_pad = sum(prof[(st, kc)][2] for st in range(2))
        k_n = sum(prof[(st, kc)][3] for st in range(2))
        tot_ms = sum(v[0] for v in prof.values())
        tot_fl = sum(v[1] for v in prof.values())
        kname = "grouped_gemm_kernel<%d,%d,*> (FP64 DMMA m16n8k8, %d warps x 32x32)" % (128 >> (kc // 3), 128 >> (kc % 3), (128 >> (kc // 3)) * (128 >> (kc % 3)) // 1024) if kc < 9 else "tiny_gemm_kernel"
        achieved = k_fl / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
        # DRAM bytes per launch of that kernel from the committed ncu capture of this same workload (profiles/README.md); other
        # workloads have no capture
        default_workload = (a.norbs, a.nelec, a.M, a.left_sites, world, a.mode) == (40, 40, 4000, 18, 1, "materialised")
        default_factorised = (a.norbs, a.nelec, a.M, a.left_sites, world, a.mode) == (40, 40, 4000, 20, 1, "factorised")
        roof = {"bound": "tensor", "achieved": achieved, "peak": dmma, "unit": "TFLOP/s", "frac": achieved / dmma if dmma else None,
                "traffic": 16.12e9 if default_workload else (14.27e9 if default_factorised and kc == 0 else None),
                "traffic_source": ("profiles/r01_ncu_launches_sigma_fullsize.csv: dram__bytes_read.sum + dram__bytes_write.sum of the 36 launches / 36 (bytes per launch)" if default_workload else
                                   ("profiles/r02_ncu_launches_sigma_factorised.csv: dram__bytes_read.sum + dram__bytes_write.sum of the 88 launches of grouped_gemm_kernel<128,128,1> / 88 "
                                    "(bytes per launch); whole sigma 1.09 TB = 0.86 TB/s: FP64-pipe bound" if default_factorised and kc == 0 else None)),
                "kernel": kname, "launches": int(k_n), "avg_launch_ms": k_ms / max(k_n, 1),
                "flops_basis": "EXECUTED flops of that kernel (useful 2mnk inside its tiles; structural zeros the factorised form skips are not credited)",
                "whole_sigma_executed_tflops": tot_fl / (ms_step * 1e-3) / 1e12, "whole_sigma_frac_of_peak": tot_fl / (ms_step * 1e-3) / 1e12 / dmma if dmma else None,
                "share_of_sigma": k_ms / tot_ms if tot_ms else None, "tile_fill": k_fl / k_pad if k_pad else None,
                "peak_live": dmma_live,
                "peak_source": ("profiles/r02_fp64_yardstick.json (pinned FP64 DMMA register-loop yardstick with its clock record; MEASURED_PEAKS.json has no FP64 entry); "
                                if pinned else "live FP64 DMMA register-loop yardstick of this library on this GPU (MEASURED_PEAKS.json has no FP64 entry); ") + "live this run: DMMA %.2f, DFMA %.1f TFLOP/s" % (dmma_live, dfma),
                "per_class": {("step%d_%dx%d" % (st + 1, 128 >> (c // 3), 128 >> (c % 3)) if c < 9 else "step%d_tiny8x8_warp" % (st + 1)): {"ms": v[0], "tflops": (v[1] / (v[0] * 1e-3) / 1e12 if v[0] > 0 else 0.0), "tile_fill": (v[1] / v[2] if v[2] else None),
                                                               "launches": int(v[3])} for (st, c), v in prof.items() if v[3] > 0}}
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "executed": {"flops_per_step": stats["flops_executed"], "gflops": stats["flops_executed"] / (ms_step * 1e-3) / 1e9,
                             "ratio_to_algorithmic": stats["flops_executed"] / flops_mine if flops_mine else None,
                             "note": "value = ALGORITHMIC flops (the dgemm flops the reference issues for this multiplyH, SURVEY 8d) / time; executed = what the kernels "
                                     "compute: the factorised operators skip the structural zeros of the Kronecker blocks, T blocks nobody consumes are skipped in both modes"},
                "config": {"workload": workload_name(a), "operators": a.mode, "psi_doubles": int(W), "terms": int(len(sb.terms(all_ranks=True)[0])),
                           "left_sectors": int(len(sb.left.dims)), "right_sectors": int(len(sb.right.dims)), "left_states": int(sb.left.dims.sum()),
                           "right_states": int(sb.right.dims.sum()), "sigma_flops": flops_alg, "rank0_flops": flops_mine,
                           "operator_arena_gb_rank0": stats["arena_doubles"] * 8 / 1e9, "presummed_factor_blocks_gb": stats["combo_doubles"] * 8 / 1e9,
                           "factors": int(stats["factors_direct"] + stats["factors_combo"]), "chunks": int(stats["chunks"]),
                           "parallelism": ("operator-term partition x%d (%s) + NCCL all-reduce of partial sigma" % (world, "cost-weighted ownership" if options.get("balance_terms") else "reference's ownership rule")) if world > 1 else "single GPU",
                           "l2": "inputs larger than L2 (operator arena %.0f GB + %.1f GB of T workspace per step)" % (stats["arena_doubles"] * 8 / 1e9, stats["workspace_doubles"] * 8 / 1e9),
                           "setup_s": t_setup},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(W * 8), "d2h_bytes_per_step": int(W * 8)},
                "gpu_launches": int(launches), "clocks": clk, "roofline": roof,
                "parity": parity,
                "hbm_peak_gbs": peaks.get("hbm_gbs")}
    if line is not None and world == 1 and a.mode == "factorised":
        err, nt = factorised_self_check(sb, psi)
        line["parity"]["factorised_vs_dense_TensorMultiply_rel"] = err
        line["parity"]["factorised_vs_dense_terms"] = nt
        sb.upload(0, psi)
    if line is not None and world == 1 and not a.no_block_iteration:
        # a first pass with ONE Davidson iteration loads every kernel of the leg (CUDA loads modules lazily: the first launch of
        # a kernel costs tens of milliseconds) and sizes the scratch buffers; the second pass is the one reported
        iters = a.davidson_iters
        a.davidson_iters = 1
        block_iteration_leg(sb, a, psi, ms_step)
        a.davidson_iters = iters
        line["block_iteration"] = block_iteration_leg(sb, a, psi, ms_step)
        # HBM-bound phases (north_star: achieved HBM GB/s for the bandwidth-bound phases): the Davidson level-1 kernels alone
        hbm = peaks.get("hbm_gbs") or 6534.5
        lv = sb.measure_level1(reps=10)
        line["level1_hbm"] = {"peak_gbs": hbm, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks.get("hbm_gbs") else "fallback 6534.5 GB/s",
                              "vector_mb": W * 8 / 1e6,
                              "kernels": {k: {"ms": ms_, "algorithmic_gb": by / 1e9, "gbs": by / (ms_ * 1e-3) / 1e9 if ms_ > 0 else None,
                                              "frac": by / (ms_ * 1e-3) / 1e9 / hbm if ms_ > 0 else None} for k, (ms_, by) in lv.items()}}
    if line is not None and not a.no_cpu and world == 1:
        kind, v, cores, desc, _, _ = cpu_leg(a, a.cpu_budget_s)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc}
    sb.close()
    if line is not None and world == 1 and not a.no_sweep:
        line["sweep"] = sweep_leg(a)
    if line is not None and world == 1 and not a.no_block_iteration:
        line["next_rows"] = next_rows_leg(a)      # guess-wavefunction transform (N1) + "operator_construction" (N2) at this size
    if line is not None:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    # stdout carries exactly ONE JSON line: everything libraries print (NCCL banners, torchrun notices) goes to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--norbs", type=int, default=40)
    ap.add_argument("--nelec", type=int, default=40)
    ap.add_argument("--M", type=int, default=4000)
    # 18|22: the heaviest block iteration of the 40-orbital M=4000 sweep whose materialised operator arenas (152 GB) fit
    # ONE 180 GB B200; the mid-chain 20|20 iteration (185 GB) needs the term partition over >= 2 GPUs
    ap.add_argument("--left-sites", type=int, default=None)
    # factorised (default): the operators of the two enlarged blocks are lists of scaled sub-blocks of the renormalised operators
    # (SURVEY.md 7 "hard parts") - the mid-chain 20|20 block iteration, the heaviest of the sweep, fits one GPU; materialised: round 1's
    # dense enlarged-block operators (18|22 is the heaviest that fits one 180 GB GPU)
    ap.add_argument("--mode", default="factorised", choices=["factorised", "materialised"])
    ap.add_argument("--workspace-mb", type=float, default=8192.0)
    ap.add_argument("--slice-iters", type=int, default=256)
    ap.add_argument("--opt", action="append", default=[], help="extra library option key=value (b2d_set_option)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-block-iteration", action="store_true", help="skip the Davidson / density / eigen / rotation leg")
    ap.add_argument("--davidson-iters", type=int, default=6)
    ap.add_argument("--no-sweep", action="store_true", help="skip the whole-sweep leg (reference sweep vs the same sweep with the GPU hot path)")
    ap.add_argument("--sweep-case", default="synthetic_16o_M300", help="case of tests/golden/dropin_cases.npz for the sweep leg")
    ap.add_argument("--profile-mode", action="store_true", help="for ncu: 1 warm-up sigma + --steps sigmas, nothing else, no JSON line")
    ap.add_argument("--cpu-budget-s", type=float, default=20.0)
    ap.add_argument("--ref-total-s", type=float, default=360.0, help="--impl reference: CPU time budget of the whole run (steps are shortened to fit)")
    ap.add_argument("--ref-step-s", type=float, default=20.0)   # >= 128 terms of the multiplyH per reference step on a 16-core host
    a = ap.parse_args()
    if a.left_sites is None:
        a.left_sites = a.norbs // 2 if a.mode == "factorised" else 18
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
