mkdir -p gpurun_out/r15
timeout 125 python -m pytest tests/test_gpu_dropin.py -m gpu -n 8 -v -p no:cacheprovider 2>&1 | tee gpurun_out/r15/pytest_dropin.txt | grep -v "^$" | tail -40
