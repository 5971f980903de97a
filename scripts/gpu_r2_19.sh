# round 2, call 19: sigma kernel with the prefetched segment descriptor - parity tests of every tile class, then A/B of the split-K
# slice length of narrow sigma blocks (option slice_iters_narrow) on the benchmark workload
mkdir -p gpurun_out/r2_19
timeout 900 python -m pytest tests/test_gpu_hotpath.py tests/test_gpu_factorised.py tests/test_synthetic.py tests/test_gpu_eig.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2_19/pytest.txt
for OPT in "" "slice_iters_narrow=64" "slice_iters_narrow=24"; do
  TAG=${OPT:-base}
  timeout 600 python bench.py --no-cpu --no-block-iteration --no-sweep --steps 4 --warmup 3 ${OPT:+--opt $OPT} > gpurun_out/r2_19/bench_$TAG.json 2> gpurun_out/r2_19/bench_$TAG.err
  python - "$TAG" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r2_19/bench_%s.json" % tag) if l.startswith("{")][-1])
    pc = d["roofline"]["per_class"]
    narrow = sum(v["ms"] for k, v in pc.items() if k.startswith("step2") and not k.startswith("step2_128x") and "tiny" not in k)
    print(tag, "value %.0f GFLOP/s" % d["value"], "ms_per_step %.1f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "narrow step-2 classes %.1f ms" % narrow,
          "128x128: %.1f + %.1f ms" % (pc["step1_128x128"]["ms"], pc["step2_128x128"]["ms"]), "parity", d.get("parity"))
except Exception as e:
    print(tag, "failed", e, open("gpurun_out/r2_19/bench_%s.err" % tag).read()[-800:])
PY
done
