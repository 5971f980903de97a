cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
python - <<'P'
import numpy as np, glob
from block_b200 import hotpath
from oracle import dumpio
rec = dumpio.read_records(sorted(glob.glob('tests/golden/*.npz'))[0])
sb = hotpath.spinblock_from_record(rec, device=0)
print("FP64 peaks (DMMA, DFMA) TFLOP/s:", sb.measure_fp64_peak())
print("FP64 peaks again:", sb.measure_fp64_peak())
P
