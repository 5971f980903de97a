mkdir -p gpurun_out/r7
timeout 600 python -m pytest tests/test_synthetic.py tests/test_gpu_hotpath.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r7/pytest_core.txt
for w in "--opt persistent=0" "--opt persistent=1" "--opt persistent=1 --workspace-mb 12288" "--opt persistent=1 --slice-iters 512"; do timeout 600 python bench.py --no-cpu --no-block-iteration --no-sweep --steps 3 --warmup 3 $w 2> gpurun_out/r7/exp.err | python -c "
import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=l['roofline']['per_class']; print('$w', round(l['value']), round(l['ms_per_step'],1), 'frac', round(l['roofline']['frac'],4), 'chunks', l['config']['chunks'], 's1_128', round(p['step1_128x128']['tflops'],2), 's2_128', round(p['step2_128x128']['tflops'],2), 'lin', l['parity']['linearity_rel'])" | tee -a gpurun_out/r7/sigma_experiments.txt; tail -2 gpurun_out/r7/exp.err; done
timeout 900 python -m pytest tests/test_gpu_dropin.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r7/pytest_dropin.txt; cat gpurun_out/r7/pytest_dropin.txt
python - <<'PY' 2>&1 | tee gpurun_out/r7/sweep18.txt
import sys, types, json, os
sys.path.insert(0, '.')
import bench
for opts in ("", "eig_jacobi_max=256"):
    if opts: os.environ["B2D_DROPIN_OPTIONS"] = opts
    a = types.SimpleNamespace(sweep_case="synthetic_18o_M500")
    r = bench.sweep_leg(a)
    print(opts or "default", json.dumps(r["gpu_dropin"]), r.get("max_abs_dE_per_sweep"), "ref", r["reference_cpu"]["wall_s"], flush=True)
PY
