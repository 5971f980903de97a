# round 2, call 22: a split-K family of their own for the narrow tiles of every sigma block (option slice_iters_narrow): parity, then A/B
mkdir -p gpurun_out/r2_22
run() {
  TAG=$1; shift
  timeout 600 python bench.py --no-cpu --no-block-iteration --no-sweep --steps 4 --warmup 3 "$@" > gpurun_out/r2_22/bench_$TAG.json 2> gpurun_out/r2_22/bench_$TAG.err
  python - "$TAG" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r2_22/bench_%s.json" % tag) if l.startswith("{")][-1])
    pc = d["roofline"]["per_class"]
    narrow = sum(v["ms"] for k, v in pc.items() if k.startswith("step2") and not k.startswith("step2_128x128") and "tiny" not in k)
    print(tag, "value %.0f GFLOP/s" % d["value"], "ms_per_step %.1f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "narrow step-2 classes alone %.1f ms" % narrow,
          "bit_reproducible", d["parity"]["bit_reproducible"], "linearity %.1e" % d["parity"]["linearity_rel"], "fact-vs-dense %.1e" % d["parity"]["factorised_vs_dense_TensorMultiply_rel"])
except Exception as e:
    print(tag, "failed", e, open("gpurun_out/r2_22/bench_%s.err" % tag).read()[-600:])
PY
}
run base
run narrow64 --opt slice_iters_narrow=64
run narrow32 --opt slice_iters_narrow=32
run narrow128 --opt slice_iters_narrow=128
