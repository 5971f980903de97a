# round 2, call 24 (2 GPUs): partitioned eigen-decomposition + rotation - the 2-GPU drop-in tests, then P4 on 2 GPUs
mkdir -p gpurun_out/r2_24
timeout 900 python -m pytest tests/test_gpu_multirank_dropin.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2_24/pytest.txt
P4_GPUS=2 P4_OUT=gpurun_out/r2_24/p4_n2 timeout 900 bash scripts/gpu_p4.sh 2>&1 | tail -34
