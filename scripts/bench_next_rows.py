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
    finally:
        gt.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
