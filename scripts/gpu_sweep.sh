cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run() {
  timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --workspace-mb 8192 "$@" > gpurun_out/sweep.json 2> gpurun_out/sweep.err || tail -2 gpurun_out/sweep.err
  python - "$*" <<'P'
import json, sys
try:
    d=json.load(open('gpurun_out/sweep.json')); r=d['roofline']
    print(sys.argv[1], "| TF %.2f"%(d['value']/1e3), "ms %.1f"%d['ms_per_step'], "class0: ach %.2f share %.3f fill %.3f"%(r['achieved'], r['share_of_sigma'], r['tile_fill'] or 0), "launches", d['gpu_launches'])
    for k,v in r['per_class'].items():
        if v['ms']>50: print("    ", k, "ms %.1f tf %.2f fill %.3f"%(v['ms'], v['tflops'], v['tile_fill']))
except Exception as e: print("failed", sys.argv[1:], e)
P
}
run
run --opt tile_class=11
run --opt tile_class=20
B2D_TRACE=gpurun_out/trace_sigma.csv timeout 600 python bench.py --profile-mode --steps 1 --workspace-mb 8192 > /dev/null 2>&1; wc -l gpurun_out/trace_sigma.csv
