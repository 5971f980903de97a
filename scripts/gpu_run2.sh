# first full GPU pass of the round: parity tests, smoke, bench (N=1), ncu launch list + one full capture
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm,power.limit --format=csv
nproc; free -g | head -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:grouped_gemm -s 432 -c 432 --csv --log-file gpurun_out/launches_sigma.csv python bench.py --profile-mode --steps 1 > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:grouped_gemm -s 200 -c 6 -o gpurun_out/prof_gemm python bench.py --profile-mode --steps 1 --M 2000 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
