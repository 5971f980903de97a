mkdir -p gpurun_out/r14
for c in "c2_d2h_M50_noise B2D_DROPIN_CHECK=1" "c2_d2h_M50_onedot_tail" "synthetic_14o_M200" "hubbard_L16_M80" "synthetic_18o_M500" "synthetic_18o_M500 B2D_DROPIN_OPBUILD=host"; do timeout 600 python scripts/run_dropin_case.py $c --out gpurun_out/r14/dropin 2>&1 | grep -v "^  M=\(50\|80\) " | tee -a gpurun_out/r14/dropin_summary.txt; done
grep -h "opbuild" gpurun_out/r14/dropin/*CHECK*.stderr.txt | awk '{print $4,$5,$6}' | sort | uniq -c | sort -k4 | tail -5
grep -h "opbuild" gpurun_out/r14/dropin/*CHECK*.stderr.txt | grep -o "max_abs_diff=[^ ]*" | cut -d= -f2 | sort -g | tail -1
for f in synthetic_18o_M500 synthetic_18o_M500_B2D_DROPIN_OPBUILD; do python - <<PY
import re
tot={}
for l in open("gpurun_out/r14/dropin/$f.stats.txt"):
    for k,v in re.findall(r"(\w+)=([-\d.e+]+)", l): tot[k]=tot.get(k,0)+float(v)
print("$f", {k:round(v,2) for k,v in tot.items() if k.endswith("_s")})
PY
done
grep -h "Sweep Energy" gpurun_out/r14/dropin/synthetic_18o_M500*.stdout.txt
