# round 2, call 21: who gets the SMs - A/B of stream priorities (narrow tile classes first) and of the persistent kernel's CTA count
# (the B2D_NARROW_PRIORITY / B2D_PERSISTENT_CTAS knobs this script sets existed only for this experiment: no effect, reverted - profiles/README.md)
mkdir -p gpurun_out/r2_21
run() {
  TAG=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu --no-block-iteration --no-sweep --steps 4 --warmup 3 > gpurun_out/r2_21/bench_$TAG.json 2> gpurun_out/r2_21/bench_$TAG.err
  python - "$TAG" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/r2_21/bench_%s.json" % tag) if l.startswith("{")][-1])
    print(tag, "value %.0f GFLOP/s" % d["value"], "ms_per_step %.1f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "bit_reproducible", d["parity"]["bit_reproducible"], "linearity %.1e" % d["parity"]["linearity_rel"])
except Exception as e:
    print(tag, "failed", e, open("gpurun_out/r2_21/bench_%s.err" % tag).read()[-600:])
PY
}
run base B2D_X=0
run narrow_first B2D_NARROW_PRIORITY=1
run ctas132 B2D_PERSISTENT_CTAS=132
run ctas116 B2D_PERSISTENT_CTAS=116
run narrow_first_ctas140 B2D_NARROW_PRIORITY=1 B2D_PERSISTENT_CTAS=140
