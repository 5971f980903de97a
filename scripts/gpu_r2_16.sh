# round 2, call 16: block-cache tests (spill / eviction), then P4 through the drop-in with FACTORISED enlarged-block operators on one GPU
mkdir -p gpurun_out/r2_16
timeout 600 python -m pytest tests/test_gpu_block_cache.py tests/test_gpu_factorised.py -x -q 2>&1 | tail -15 | tee gpurun_out/r2_16/pytest.txt
P4_OPTIONS="factorised=1" P4_OUT=gpurun_out/r2_16/p4_fact timeout 900 bash scripts/gpu_p4.sh 2>&1 | tail -30
