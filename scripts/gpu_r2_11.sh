# Round 2, GPU call 11 (2 GPUs): multi-process drop-in sweep; bench at N = 2 (factorised, cost-weighted ownership)
O=gpurun_out/r2_11
mkdir -p $O
nvidia-smi -L | tee $O/gpus.txt
timeout 900 python -m pytest tests/test_gpu_multirank_dropin.py -m gpu -x -q -s 2>&1 | tail -12 | tee $O/pytest_multirank.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/run_dropin_multigpu.py synthetic_16o_M300 --out $O 2>&1 | tail -3 | tee $O/multigpu_16o.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
tail -c 1200 $O/bench_n2.json; tail -3 $O/bench_n2.err
