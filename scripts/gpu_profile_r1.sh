# round-1 evidence: launch list with DRAM traffic + DMMA pipe activity of one full-size sigma, then the bench lines
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --kernel-name-base demangled -k regex:grouped_gemm -s 324 -c 324 --csv --log-file gpurun_out/launches_sigma.csv python bench.py --profile-mode --steps 1 > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log; wc -l gpurun_out/launches_sigma.csv
timeout 1200 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 1500 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; head -c 300 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
