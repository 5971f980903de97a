cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 2500 gpurun_out/bench_n2.json | head -c 2500; tail -5 gpurun_out/bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --left-sites 20 > gpurun_out/bench_n2_mid.json 2> gpurun_out/bench_n2_mid.err; head -c 1200 gpurun_out/bench_n2_mid.json; tail -3 gpurun_out/bench_n2_mid.err
