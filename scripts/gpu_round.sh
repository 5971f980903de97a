mkdir -p gpurun_out/r10
python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r10/pytest_gpu.txt; cat gpurun_out/r10/pytest_gpu.txt
python bench.py > gpurun_out/r10/bench_n1.json 2> gpurun_out/r10/bench_n1.err; python -c "
import json; l=json.loads(open('gpurun_out/r10/bench_n1.json').read().strip().splitlines()[-1]); print(l['value'], l['e2e']['value'], l['roofline']['frac']); print(l['block_iteration']); print(l['sweep'])"; tail -2 gpurun_out/r10/bench_n1.err
