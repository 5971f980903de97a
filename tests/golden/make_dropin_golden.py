#!/usr/bin/env python
"""Golden per-sweep energies for the drop-in test (tests/test_gpu_dropin.py): the UNMODIFIED reference
(oracle/_ref/block.spin_adapted, no hooks) is run on each case; its "Sweep Energy" lines, the inputs it read and its wall
time are stored in tests/golden/dropin_cases.npz.  The GPU test feeds the same inputs to oracle/_ref/block_gpu (same reference
objects, hot path re-routed to libblockb200.so) and compares sweep by sweep (north_star: per-sweep energies within 1e-8 Eh).

Run in the build container only (needs /root/reference and `make -C oracle ref`).  Inputs that come from the reference's
dmrg_tests/ are stored as bytes inside the .npz (fixtures, not sources); the Hubbard chain and the synthetic random-integral
FCIDUMPs are generated here (SURVEY.md 8d, P3 / P5 recipes).
"""
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("BLOCK_REFERENCE", "/root/reference")
BLOCK = os.path.join(ROOT, "oracle", "_ref", "block.spin_adapted")


def hubbard_chain_fcidump(L, U=2.0, t=1.0):
    """Open 1-D chain in the format / sign convention of dmrg_tests/hubbard/FCIDUMP (which is the periodic 8-ring)."""
    lines = [" &FCI NORB= %d,NELEC= %d,MS2= 0," % (L, L), "  ORBSYM=" + ",".join(["1"] * L) + ",", "  ISYM=1", " &END"]
    for i in range(1, L + 1):
        lines.append("%.2f\t%d\t%d\t%d\t%d" % (U, i, i, i, i))
    for i in range(1, L):
        lines.append("%.2f\t%d\t%d\t0\t0" % (t, i, i + 1))
    return "\n".join(lines) + "\n"


def synthetic_fcidump(norb, nelec, seed=20260):
    """h symmetric N(0,1)*0.5 with a diagonal ramp, (ij|kl) = sum_P L^P_ij L^P_kl (positive, 8-fold symmetric), C1."""
    rng = np.random.default_rng(seed)
    h = rng.normal(size=(norb, norb)) * 0.5
    h = 0.5 * (h + h.T) + np.diag(np.linspace(-2.0, 2.0, norb))
    Lp = rng.normal(size=(3 * norb, norb, norb)) * 0.1
    Lp = 0.5 * (Lp + Lp.transpose(0, 2, 1))
    eri = np.einsum("pij,pkl->ijkl", Lp, Lp)
    out = [" &FCI NORB= %d,NELEC=%d,MS2= 0," % (norb, nelec), "  ORBSYM=" + ",".join(["1"] * norb) + ",", "  ISYM=1", " &END"]
    for i in range(norb):
        for j in range(i + 1):
            for k in range(i + 1):
                for l in range(k + 1):
                    if i * (i + 1) // 2 + j < k * (k + 1) // 2 + l:
                        continue
                    out.append("%.15e %d %d %d %d" % (eri[i, j, k, l], i + 1, j + 1, k + 1, l + 1))
    for i in range(norb):
        for j in range(i + 1):
            out.append("%.15e %d %d 0 0" % (h[i, j], i + 1, j + 1))
    out.append("%.15e 0 0 0 0" % 0.0)
    return "\n".join(out) + "\n"


def ref_file(*parts):
    return open(os.path.join(REF, "dmrg_tests", *parts)).read()


def cases():
    c = {}
    # P1 (BASELINE configs[0]): C2 / D2h, small M two-site sweeps, two state-averaged roots
    c["c2_d2h_M50"] = dict(files={"FCIDUMP": ref_file("c2_d2h_smallM", "FCIDUMP"), "reorder.dat": ref_file("c2_d2h_smallM", "reorder.dat")}, conf="""nelec 8
spin 0
irrep 1
hf_occ integral
schedule
0 50 1.0e-16 0.0
end
maxiter 4
twodot
sweep_tol 1e-12
sym d2h
orbitals FCIDUMP
nroots 2
weights 0.5 0.5
reorder reorder.dat
outputlevel 0
""")
    # the same with perturbative noise in the first sweeps (DensityMatrix::add_onedot_noise inside the sweep)
    c["c2_d2h_M50_noise"] = dict(files=c["c2_d2h_M50"]["files"], conf="""nelec 8
spin 0
irrep 1
hf_occ integral
schedule
0 50 1.0e-16 1.0e-4
2 50 1.0e-16 0.0
end
maxiter 4
twodot
sweep_tol 1e-12
sym d2h
orbitals FCIDUMP
nroots 2
weights 0.5 0.5
reorder reorder.dat
outputlevel 0
""")
    # RANDOM noise as well (twodot_noise keyword -> DensityMatrix::add_twodot_noise, density.C:92-165): the hook draws the random
    # wavefunctions with the reference's own Randomise, so both runs consume the same glibc rand() stream
    c["c2_d2h_M50_twodotnoise"] = dict(files=c["c2_d2h_M50"]["files"], conf="""nelec 8
spin 0
irrep 1
hf_occ integral
schedule
0 50 1.0e-16 1.0e-4
2 50 1.0e-16 0.0
end
maxiter 4
twodot
twodot_noise 1.0e-4 0.3
sweep_tol 1e-12
sym d2h
orbitals FCIDUMP
nroots 2
weights 0.5 0.5
reorder reorder.dat
outputlevel 0
""")
    # the reference's default algorithm: two-dot sweeps, then one-dot sweeps (twodot_to_onedot).  In the one-dot sweeps the dot
    # alternates between the system and the environment side; on the environment side RenormaliseFrom solves on
    # system x (dot+environment) and reshuffles the solution to (system+dot) x environment before the density matrix (renormalise.C:64-79)
    c["c2_d2h_M50_onedot_tail"] = dict(files=c["c2_d2h_M50"]["files"], conf="""nelec 8
spin 0
irrep 1
hf_occ integral
schedule
0 50 1.0e-16 1.0e-4
2 50 1.0e-16 0.0
end
maxiter 6
twodot_to_onedot 3
sweep_tol 1e-12
sym d2h
orbitals FCIDUMP
nroots 2
weights 0.5 0.5
reorder reorder.dat
outputlevel 0
""")
    # P2 (configs[1]): H2O, no symmetry, M = 500
    c["h2o_nosym_M500"] = dict(files={"FCIDUMP": ref_file("h2o_nosym", "FCIDUMP")}, conf="""nelec 10
spin 0
irrep 1
hf_occ integral
schedule
0 500 1.0e-16 0.0
end
maxiter 8
twodot
sweep_tol 1e-12
orbitals FCIDUMP
noreorder
outputlevel 0
""")
    # the same molecule with M small enough that the top-M cut (not the 1e-13 weight threshold) decides the retained basis
    c["h2o_nosym_M60"] = dict(files=c["h2o_nosym_M500"]["files"], conf=c["h2o_nosym_M500"]["conf"].replace("0 500 1.0e-16 0.0", "0 60 1.0e-16 0.0").replace("maxiter 8", "maxiter 4"))
    # P3 (configs[2]): 1-D Hubbard chain, spin-adapted, M = 1000, noise in the first sweeps
    c["hubbard_L16_M1000"] = dict(files={"FCIDUMP": hubbard_chain_fcidump(16)}, conf="""nelec 16
spin 0
hf_occ integral
schedule
0 200 1.0e-16 1e-4
4 1000 1.0e-16 0.0
end
maxiter 8
twodot
sweep_tol 1e-12
orbitals FCIDUMP
noreorder
outputlevel 0
warmup local_2site
""")
    c["hubbard_L16_M80"] = dict(files=c["hubbard_L16_M1000"]["files"], conf=c["hubbard_L16_M1000"]["conf"].replace("0 200 1.0e-16 1e-4", "0 80 1.0e-16 1e-4").replace("4 1000 1.0e-16 0.0", "4 80 1.0e-16 0.0"))
    # P5 scaled down (configs[4] shape: random-integral FCIDUMP, half filling, C1) so that the CPU reference finishes in a minute
    c["synthetic_14o_M200"] = dict(files={"FCIDUMP": synthetic_fcidump(14, 14)}, conf="""nelec 14
spin 0
irrep 1
hf_occ integral
schedule
0 100 1.0e-16 0.0
2 200 1.0e-16 0.0
end
maxiter 4
twodot
sweep_tol 1e-12
orbitals FCIDUMP
noreorder
outputlevel 0
""")
    # bench.py's sweep leg only (no golden sweeps stored: the reference and the GPU drop-in are run side by side on the box):
    # large enough that sigma dominates the CPU run
    c["synthetic_18o_M500"] = dict(golden=False, files={"FCIDUMP": synthetic_fcidump(18, 18)}, conf="""nelec 18
spin 0
irrep 1
hf_occ integral
schedule
0 200 1.0e-8 0.0
1 500 1.0e-8 0.0
end
maxiter 2
twodot
sweep_tol 1e-12
orbitals FCIDUMP
noreorder
outputlevel 0
""")
    # bench.py's sweep leg WITH golden sweeps (round 2): Davidson converged to 1e-12 so that the energies are determined to well below 1e-8,
    # four sweeps, the top-M cut decides the retained basis (well conditioned); generated with 8 host threads (the thread count only changes
    # the summation order of the thread-private sigma accumulators)
    c["synthetic_16o_M300"] = dict(threads=8, files={"FCIDUMP": synthetic_fcidump(16, 16)}, conf="""nelec 16
spin 0
irrep 1
hf_occ integral
schedule
0 150 1.0e-12 0.0
2 300 1.0e-12 0.0
end
maxiter 4
twodot
sweep_tol 1e-12
orbitals FCIDUMP
noreorder
outputlevel 0
""")
    # P4 stand-in (BASELINE configs[3]; SURVEY 8d: the C2 cc-pVDZ file is not in the tree): arenes/28_28_fie, 28 orbitals / 28 electrons, C1,
    # at an M the CPU reference still finishes here (8 host threads); the M = 2000 run of the same FCIDUMP is GPU-only (scripts/gpu_p4.sh)
    c["arenes28_M400"] = dict(threads=8, files={"FCIDUMP": open(os.path.join(REF, "dmrg_tests", "dmrg_parameters", "arenes", "28_28_fie", "FCIDUMP")).read()}, conf="""nelec 28
spin 0
irrep 1
hf_occ integral
schedule
0 200 1.0e-10 1.0e-4
2 400 1.0e-10 0.0
end
maxiter 4
twodot
sweep_tol 1e-12
orbitals FCIDUMP
noreorder
outputlevel 0
warmup local_2site
""")
    # The reference's OWN known-answer test, verbatim (dmrg_tests/runtest:19-23: h2o_nosym, default schedule with noise, default orbital
    # reordering, two-dot -> one-dot, `test_energy.py 1 1.0e-6 -76.11460447`).  Not a per-sweep golden case (random noise, threshold
    # regime): its sweeps are stored as "/ref_sweeps" and the GPU test applies the reference's own acceptance criterion.
    c["h2o_nosym_runtest"] = dict(golden="runtest", runtest_energy=-76.11460447, runtest_tol=1.0e-6,
                                  files={"FCIDUMP": open(os.path.join(REF, "dmrg_tests", "h2o_nosym", "FCIDUMP")).read()},
                                  conf=open(os.path.join(REF, "dmrg_tests", "h2o_nosym", "dmrg.conf")).read())
    return c


def run_reference(name, case, threads=8):
    work = tempfile.mkdtemp(prefix="dropin_")
    for f, text in case["files"].items():
        open(os.path.join(work, f), "w").write(text)
    nthreads = int(case.get("threads", 1))
    open(os.path.join(work, "dmrg.conf"), "w").write(case["conf"] + ("threads_per_node %d\n" % nthreads if nthreads > 1 else ""))
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS=str(nthreads))
    t0 = time.time()
    out = subprocess.run([BLOCK, "dmrg.conf"], cwd=work, env=env, capture_output=True, text=True)
    dt = time.time() - t0
    shutil.rmtree(work)
    if out.returncode != 0:
        print(out.stdout[-3000:], out.stderr[-3000:])
        raise SystemExit("reference run failed for " + name)
    sweeps = [l.strip() for l in out.stdout.splitlines() if "Sweep Energy" in l]
    return sweeps, dt


def main():
    only = sys.argv[1:]
    path = os.path.join(HERE, "dropin_cases.npz")
    store = dict(np.load(path, allow_pickle=False)) if os.path.exists(path) else {}
    for name, case in cases().items():
        if only and name not in only:
            continue
        if case.get("golden", True) == "runtest":
            sweeps, dt = run_reference(name, case)
            print(name, "%.1f s" % dt, sweeps[-1])
            store[name + "/ref_sweeps"] = np.frombuffer("\n".join(sweeps).encode(), dtype=np.uint8)
            store[name + "/ref_wall_s"] = np.array([dt])
            store[name + "/runtest"] = np.array([case["runtest_energy"], case["runtest_tol"]])
        elif case.get("golden", True):
            sweeps, dt = run_reference(name, case)
            print(name, "%.1f s" % dt)
            for s in sweeps:
                print("   ", s)
            store[name + "/sweeps"] = np.frombuffer("\n".join(sweeps).encode(), dtype=np.uint8)
            store[name + "/ref_wall_s"] = np.array([dt])
        store[name + "/conf"] = np.frombuffer(case["conf"].encode(), dtype=np.uint8)
        store[name + "/files"] = np.array(sorted(case["files"]))
        for f, text in case["files"].items():
            store[name + "/file/" + f] = np.frombuffer(text.encode(), dtype=np.uint8)
    np.savez_compressed(path, **store)
    print("wrote", path, "%.1f kB" % (os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    main()
