# Round 2, GPU call 9: full GPU suite, default bench (factorised 20|20 + sweep leg with per-sweep times), ncu launch list
O=gpurun_out/r2_09
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/pytest_gpu.txt
timeout 1500 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
tail -c 600 $O/bench_n1.json; tail -3 $O/bench_n1.err
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv \
  --log-file $O/ncu_launches_sigma_factorised.csv python bench.py --profile-mode --steps 1 > $O/ncu_launches.log 2>&1
tail -2 $O/ncu_launches.log; wc -l $O/ncu_launches_sigma_factorised.csv
