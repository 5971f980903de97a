#!/usr/bin/env python
"""bench.py's leg for the rows of SURVEY.md 8f that are behind the C ABI but were finished after the hot path: the guess-wavefunction
transform (N1) at the benchmark's size.  Runs in its own process (bench.py starts it after the sigma context is closed) and prints
ONE JSON object; a failure here never touches the headline measurement.  usage: bench_next_rows.py --norbs 40 --nelec 40 --M 4000 --left-sites 18"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from block_b200 import _lib, hotpath, synthetic  # noqa: E402


def opbuild_leg(a):
    """Construction of the enlarged left block's operators on the device at the benchmark's sector sizes: host planning +
    kron_scatter_kernel, one launch per product (default) and batched (option opbuild_batch); algorithmic bytes from the library
    (b2d_product_stats), device time of the batched flush from CUDA events (b2d_last_timing)."""
    lib = _lib.load()

    def run(case, batch):
        left, dot, ops, pt, h1, h2 = case
        o2n = [list(pt["old_to_new"][pt["old_to_new_begin"][i]:pt["old_to_new_begin"][i + 1]]) for i in range(len(pt["dims"]))]
        t0 = time.perf_counter()
        pb = hotpath.ProductBlock(left, dot, pt["q"], pt["dims"], pt["unc.lmap"], pt["unc.rmap"], pt["unc.dims"], o2n, device=a.device)
        try:
            assert lib.b2d_set_option(pb._ctx, b"opbuild_batch", float(batch)) == 0
            pb.set_integrals(h1, h2, np.zeros(h1.shape[0], np.int32))
            lib.b2d_sync(pb._ctx)
            upload_s = time.perf_counter() - t0
            l0 = pb.kernel_launches()
            t0 = time.perf_counter()
            for op in ops:
                pb.build(op.optype, op.orbs, op.dq, op.fermion, False, comp=op.comp)
            sites = np.asarray(list(left.sites) + list(dot.sites), dtype=np.int32)
            pb._ck(lib.b2d_stash_product(pb._ctx, 0, 1, len(sites), sites.ctypes.data_as(_lib.c_i32p)))     # runs whatever was deferred
            lib.b2d_sync(pb._ctx)
            wall = time.perf_counter() - t0
            st, tm = np.zeros(4), np.zeros(4)
            lib.b2d_product_stats(pb._ctx, st.ctypes.data_as(_lib.c_f64p), 4)
            if batch:
                lib.b2d_last_timing(pb._ctx, tm.ctypes.data_as(_lib.c_f64p), 4)
            return {"wall_ms": wall * 1e3, "children_upload_s": upload_s, "launches": pb.kernel_launches() - l0, "products": int(st[0]), "scatter_tasks": int(st[1]),
                    "algorithmic_gb": st[2] / 1e9, "rounds": int(st[3]), "device_ms": float(tm[0]) if batch else None}
        finally:
            pb.close()

    run(synthetic.make_opbuild_case(12, 12, 60, 6, op_sites=3, other_sites=3), 1)          # loads the kernels
    case = synthetic.make_opbuild_case(a.norbs, a.nelec, a.M, a.left_sites)
    dims = case[3]["dims"].astype(np.int64)
    out = {"workload": "operators of the enlarged left block (%d states in %d sectors) from an M = %d block x one site: %d operators, %.1f GB" %
                       (int(dims.sum()), len(dims), a.M, len(case[2]), sum(int((op.allowed.astype(np.int64) * np.outer(dims, dims)).sum()) for op in case[2]) * 8 / 1e9),
           "one_launch_per_product": run(case, 0), "batched": run(case, 1)}
    b = out["batched"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:   # noqa: BLE001
        pass
    hbm = peaks.get("hbm_gbs") or 6534.5
    if b["device_ms"]:
        out["roofline"] = {"bound": "hbm", "kernel": "kron_scatter_kernel", "achieved": b["algorithmic_gb"] / (b["device_ms"] * 1e-3), "peak": hbm, "unit": "GB/s",
                           "frac": b["algorithmic_gb"] / (b["device_ms"] * 1e-3) / hbm, "launches": b["rounds"],
                           "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks.get("hbm_gbs") else "fallback 6534.5 GB/s"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--norbs", type=int, default=40)
    ap.add_argument("--nelec", type=int, default=40)
    ap.add_argument("--M", type=int, default=4000)
    ap.add_argument("--left-sites", type=int, default=18)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args()
    dq, tables, allowed, lcols, rcols, old, lrot, rrot = synthetic.make_guess_case(a.norbs, a.nelec, a.M, a.left_sites)
    t0 = time.perf_counter()
    gt = hotpath.GuessTransform(dq, tables, allowed, lcols, rcols, device=a.device)
    plan_s = time.perf_counter() - t0
    out = {"workload": "guess wavefunction of the %d|%d block iteration, norbs=%d nelec=%d M=%d (GuessWave::transform_previous_wavefunction)" %
                       (a.left_sites, a.norbs - a.left_sites, a.norbs, a.nelec, a.M),
           "old_wave_doubles": gt.old_size, "trial_doubles": gt.trial_size, "gemm_flops": gt.flops, "shuffle_algorithmic_bytes": gt.shuffle_bytes,
           "shuffle_tasks": gt.shuffle_tasks, "shuffle_rounds": gt.shuffle_rounds, "plan_host_s": plan_s}
    try:
        guess_device_leg(a, gt, out, old, lrot, rrot)
    except Exception as e:   # noqa: BLE001
        out["error"] = repr(e)[:400]
    finally:
        gt.close()
    try:
        out["operator_construction"] = opbuild_leg(a)
    except Exception as e:   # noqa: BLE001
        out["operator_construction"] = {"error": repr(e)[:400]}
    print(json.dumps(out))


def guess_device_leg(a, gt, out, old, lrot, rrot):
    l0 = gt.kernel_launches()
    first = gt.transform(old, lrot, rrot)                      # loads the kernels, sizes the buffers
    out["gpu_launches_per_call"] = gt.kernel_launches() - l0
    dev_ms, e2e_ms = [], []
    tm = np.zeros(4)
    for _ in range(a.reps):
        t0 = time.perf_counter()
        got = gt.transform(old, lrot, rrot)
        e2e_ms.append((time.perf_counter() - t0) * 1e3)
        gt.lib.b2d_last_timing(gt._ctx, tm.ctypes.data_as(_lib.c_f64p), 4)
        dev_ms.append(float(tm[0]))
    out["device_ms"] = float(np.median(dev_ms))                # CUDA events around stage 1 + shuffle + stage 3
    out["e2e_ms"] = float(np.median(e2e_ms))                   # host arrays in, trial vector out (upload, pack, transform, unpack, download)
    out["h2d_bytes"] = int((gt.old_size + gt.lrot_size + gt.rrot_size) * 8)
    out["d2h_bytes"] = int(gt.trial_size * 8)
    out["bit_reproducible"] = bool(np.array_equal(first, got))
    twice = gt.transform(2.0 * old, lrot, rrot)
    out["linearity_rel"] = float(np.linalg.norm(twice - 2.0 * got) / max(np.linalg.norm(got), 1e-300))
    out["trial_norm"] = float(np.linalg.norm(got))


if __name__ == "__main__":
    main()
