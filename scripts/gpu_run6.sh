cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for M in 1000 4000; do
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu --M $M > gpurun_out/bench_bi_$M.json 2> gpurun_out/bench_bi.err; python - $M <<'P'
import json, sys
d=json.load(open('gpurun_out/bench_bi_%s.json'%sys.argv[1]))
print("M", sys.argv[1], "TF %.2f"%(d['value']/1e3), "ms %.1f"%d['ms_per_step'])
print(json.dumps(d.get('block_iteration'), indent=1))
P
tail -3 gpurun_out/bench_bi.err
done
