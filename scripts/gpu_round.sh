# First GPU call of the next round: everything that was finished after round 1's GPU budget was spent.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/gpu_round.sh'
mkdir -p gpurun_out/r16
# 1. the never-run device tests, with their real outcome (--runxfail turns the non-strict xfail marker off)
timeout 300 python -m pytest tests/test_z_gpu_next_rows.py -m gpu --runxfail -q -s 2>&1 | tail -40 | tee gpurun_out/r16/pytest_next_rows.txt
# 2. the drop-in with the device guess transform and the batched operator construction, check mode: every hook (and every transformed
#    guess: "B2D_CHECK ... guess_transform") against the reference's own CPU function
for c in "c2_d2h_M50_noise B2D_DROPIN_CHECK=1 B2D_DROPIN_GUESS=device" "c2_d2h_M50_onedot_tail B2D_DROPIN_CHECK=1 B2D_DROPIN_GUESS=device" \
         "synthetic_14o_M200 B2D_DROPIN_GUESS=device B2D_DROPIN_OPTIONS=opbuild_batch=1" "hubbard_L16_M80 B2D_DROPIN_GUESS=device B2D_DROPIN_OPTIONS=opbuild_batch=1" \
         "synthetic_14o_M200"; do
  timeout 600 python scripts/run_dropin_case.py $c --out gpurun_out/r16/dropin 2>&1 | grep -v "^  M=\(50\|80\) " | tee -a gpurun_out/r16/dropin_summary.txt
done
grep -h "guess_transform" gpurun_out/r16/dropin/*CHECK*.stderr.txt | grep -o "max_abs_diff=[^ ]*" | cut -d= -f2 | sort -g | tail -1
for f in gpurun_out/r16/dropin/synthetic_14o_M200*.stats.txt; do python - "$f" <<'PY'
import re, sys
tot = {}
for l in open(sys.argv[1]):
    for k, v in re.findall(r"(\w+)=([-\d.e+]+)", l): tot[k] = tot.get(k, 0) + float(v)
print(sys.argv[1].split("/")[-1], {k: round(v, 2) for k, v in tot.items() if k.endswith("_s") or k == "launches"})
PY
done
# 3. the bench's next-rows leg alone
timeout 300 python scripts/bench_next_rows.py 2>&1 | tail -2 | tee gpurun_out/r16/bench_next_rows.json
# 4. ncu launch list of the next-rows leg: DRAM traffic of kron_scatter_kernel (roofline.traffic of the operator construction) and of the
#    guess transform's launches; a number printed under ncu is never a bench value
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
  --log-file gpurun_out/r16/ncu_next_rows_launches.csv python scripts/bench_next_rows.py --reps 1 > gpurun_out/r16/ncu_next_rows.log 2>&1
grep -c kron_scatter gpurun_out/r16/ncu_next_rows_launches.csv
