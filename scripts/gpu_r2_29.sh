# round 2, call 29: final validation - the whole GPU suite, then the default bench line
mkdir -p gpurun_out/r2_29
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/r2_29/pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/r2_29/bench_n1.json 2> gpurun_out/r2_29/bench_n1.err; tail -c 800 gpurun_out/r2_29/bench_n1.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2_29/bench_n1.json") if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"], {k: d["roofline"][k] for k in ("achieved", "peak", "frac")}, d["clocks"])
print("block_iteration", d["block_iteration"])
sw = d.get("sweep", {})
for k in ("reference_cpu", "gpu_dropin", "gpu_dropin_factorised"):
    print(k, {x: sw.get(k, {}).get(x) for x in ("wall_s", "wall_s_per_sweep", "final_energy", "hot_path_s", "failed")})
print({k: sw.get(k) for k in ("speedup_whole_run", "speedup_regular_sweeps", "max_abs_dE_per_sweep", "max_abs_dE_vs_golden")})
print("next_rows", json.dumps(d.get("next_rows"))[:600])
PY
