cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"grouped_gemm_kernel<.int.128, .int.128" -s 4 -c 2 -o gpurun_out/prof_gemm_v3 python bench.py --profile-mode --steps 1 --M 2000 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
timeout 1200 python bench.py --steps 3 --warmup 3 --no-cpu --workspace-mb 8192 > gpurun_out/bench_prio.json 2> gpurun_out/bench_n1.err; python - <<'P'
import json
d=json.load(open('gpurun_out/bench_prio.json'))
r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['parity'])
print(r['achieved'], r['frac'], r['tile_fill'], r['share_of_sigma'])
P
tail -3 gpurun_out/bench_n1.err
