# Round 2, GPU call 10: P4 stand-in (arenes 28 orbitals) through the drop-in at M <= 400 against golden sweeps; sigma tuning experiments
O=gpurun_out/r2_10
mkdir -p $O
for c in "arenes28_M400" "arenes28_M400 B2D_DROPIN_OPTIONS=factorised=1"; do
  timeout 1500 python scripts/run_dropin_case.py $c --out $O/dropin 2>&1 | cut -c1-300 | tee -a $O/dropin.txt
done
python - <<'PY'
import re, glob
for f in sorted(glob.glob("gpurun_out/r2_10/dropin/*.stats.txt")):
    tot = {}
    for l in open(f):
        for k, v in re.findall(r"(\w+)=([-\d.e+]+)", l): tot[k] = tot.get(k, 0) + float(v)
    print(f.split("/")[-1], {k: round(v, 2) for k, v in tot.items() if k.endswith("_s") or k in ("launches", "cache_uses", "n_multiply")})
PY
grep -h "Elapsed Sweep Wall" $O/dropin/arenes28*.stdout.txt | head -12
for o in "" "--opt slice_iters=128" "--opt slice_iters=512" "--opt presum_identity=0" "--workspace-mb 16384"; do
  echo "== bench $o"
  timeout 600 python bench.py --no-sweep --no-block-iteration --no-cpu --steps 3 $o 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=d['roofline']
print('value %.0f ms %.1f exec_tflops %.2f frac_whole %.3f top %s %.3f arena %.1f GB launches %d' % (d['value'], d['ms_per_step'], r['whole_sigma_executed_tflops'], r['whole_sigma_frac_of_peak'], r['kernel'][:30], r['frac'], d['config']['operator_arena_gb_rank0'], d['gpu_launches']))
print(' '.join('%s:%.0fms/%.1f' % (k.replace('step','s'), v['ms'], v['tflops']) for k, v in sorted(r['per_class'].items(), key=lambda kv: -kv[1]['ms'])[:10]))
" | tee -a $O/tuning.txt
done
