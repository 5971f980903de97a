#!/usr/bin/env python
"""Run one drop-in case (tests/golden/dropin_cases.npz) through oracle/_ref/block_gpu and print the per-sweep comparison with the
unmodified reference's golden energies.  usage: run_dropin_case.py <case> [KEY=VALUE ...] [--out DIR]"""
import os
import sys
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
    env = dict(a.split("=", 1) for a in args[1:])
    t0 = time.time()
    out, golden, stats = T.run_case(name, env, timeout=3000)
    dt = time.time() - t0
    got = T.parse_sweeps(out.stdout)
    print("%s: exit %d, %.1f s wall, %d/%d sweep lines" % (name, out.returncode, dt, len(got), len(golden)))
    for a, b in zip(got, golden):
        print("  M=%d state=%d  E_gpu=%.10f  E_ref=%.10f  dE=%+.2e  dw_gpu=%.3e dw_ref=%.3e" % (a[0], a[1], a[3], b[3], a[3] - b[3], a[2], b[2]))
    if out_dir:
        os.makedirs(out_dir, exist_ok=True)
        tag = name + ("_" + "_".join(sorted(env)) if env else "")
        open(os.path.join(out_dir, tag + ".stdout.txt"), "w").write(out.stdout)
        open(os.path.join(out_dir, tag + ".stderr.txt"), "w").write(out.stderr)
        open(os.path.join(out_dir, tag + ".stats.txt"), "w").write(stats)
    if out.returncode != 0:
        print(out.stdout[-1500:]); print(out.stderr[-3000:])


if __name__ == "__main__":
    main()
