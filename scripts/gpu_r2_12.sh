# Round 2, GPU call 12: ncu --set full of the dominant kernel of the factorised sigma (128x128 DMMA class, step 2 and step 1)
O=gpurun_out/r2_12
mkdir -p $O
timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled --kernel-name regex:"grouped_gemm_kernel<.int.128, .int.128" -s 3 -c 2 \
  -o $O/ncu_full_gemm128_factorised python bench.py --profile-mode --steps 1 > $O/ncu_full.log 2>&1
tail -3 $O/ncu_full.log
ls -la $O
ncu -i $O/ncu_full_gemm128_factorised.ncu-rep --page raw --csv > $O/ncu_full_gemm128_factorised_raw.csv 2>/dev/null
wc -l $O/ncu_full_gemm128_factorised_raw.csv
