cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_n1_mbar.json 2> gpurun_out/bench_n1.err; python - <<'P'
import json
d=json.load(open('gpurun_out/bench_n1_mbar.json'))
r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['parity'], d['clocks'])
print(r['achieved'], r['frac'], r['tile_fill'], r['share_of_sigma'])
for k,v in r['per_class'].items(): print(k, v)
P
tail -3 gpurun_out/bench_n1.err
