"""P5 (BASELINE configs[4]): the synthetic 40-orbital / 40-electron C1 FCIDUMP (seeded generator, the same one the golden
synthetic cases use) and the P5 schedule  0 250 1e-5 / 2 1000 1e-6 / 4 4000 1e-7, two-dot, noise 0.   python scripts/p5_case.py DIR [maxiter]
writes DIR/FCIDUMP and DIR/dmrg.conf.  `warmup local_2site` as in the P4 run (the default warm-up builds determinant blocks whose cost
explodes with 40 orbitals); everything else is SURVEY 8d's configuration."""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def generator():
    spec = importlib.util.spec_from_file_location("make_dropin_golden", os.path.join(HERE, "..", "tests", "golden", "make_dropin_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.synthetic_fcidump


def conf(maxiter=6, threads=1, norb=40):
    return """nelec %d
spin 0
irrep 1
hf_occ integral
schedule
0 250 1.0e-5 0.0
2 1000 1.0e-6 0.0
4 4000 1.0e-7 0.0
end
maxiter %d
twodot
sweep_tol 1e-12
orbitals FCIDUMP
noreorder
outputlevel 0
warmup local_2site
%s""" % (norb, maxiter, "threads_per_node %d\n" % threads if threads > 1 else "")


if __name__ == "__main__":
    out = sys.argv[1]
    maxiter = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    threads = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    os.makedirs(out, exist_ok=True)
    open(os.path.join(out, "FCIDUMP"), "w").write(generator()(40, 40))
    open(os.path.join(out, "dmrg.conf"), "w").write(conf(maxiter, threads))
    print("wrote", out)
