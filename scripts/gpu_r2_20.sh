# round 2, call 20: tiled rotation-matrix gather (parity + time of b2d_select_states at the benchmark size), per-launch timeline of one sigma
mkdir -p gpurun_out/r2_20
timeout 900 python -m pytest tests/test_gpu_hotpath.py tests/test_gpu_block_cache.py tests/test_gpu_eig.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2_20/pytest.txt
timeout 600 python -m pytest tests/test_gpu_dropin.py -m gpu -x -q -k "match_reference and (c2_d2h_M50 or hubbard_L16_M1000 or synthetic_16o)" 2>&1 | tail -3 | tee -a gpurun_out/r2_20/pytest.txt
timeout 600 python bench.py --no-cpu --no-sweep --steps 3 --warmup 3 > gpurun_out/r2_20/bench.json 2> gpurun_out/r2_20/bench.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_20/bench.json") if l.startswith("{")][-1])
    print("value", d["value"], "block_iteration", d["block_iteration"])
except Exception as e:
    print("bench failed", e, open("gpurun_out/r2_20/bench.err").read()[-1500:])
PY
B2D_TRACE=gpurun_out/r2_20/trace.csv timeout 600 python bench.py --profile-mode --steps 1 > gpurun_out/r2_20/trace.log 2>&1; tail -3 gpurun_out/r2_20/trace.log; wc -l gpurun_out/r2_20/trace.csv
