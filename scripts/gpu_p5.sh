# P5 (BASELINE configs[4]): synthetic 40 orbitals / 40 electrons, C1, schedule 0 250 / 2 1000 / 4 4000, two-dot, through the drop-in
# on ONE GPU with factorised enlarged-block operators and the device block cache.  Bounded: MAXITER sweeps, optional STOP_AFTER block iterations,
# a wall-clock cap and a host-memory watchdog (the run ends itself before the box is short of memory).
#   gpurun --timeout 2400 -- 'MAXITER=6 CAP_S=1500 bash scripts/gpu_p5.sh'
R=${GRAFT_REPO_ROOT:-/root/repo}
O=$R/gpurun_out/p5
MAXITER=${MAXITER:-6}
CAP_S=${CAP_S:-1500}
mkdir -p $O /tmp/p5
python $R/scripts/p5_case.py /tmp/p5 $MAXITER > /dev/null
cd /tmp/p5
{ nproc; free -g | head -2; df -h /tmp | tail -1; nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; } > $O/box.txt 2>&1
T0=$(date +%s.%N)
env OPENBLAS_NUM_THREADS=1 OMP_NUM_THREADS=1 B2D_DROPIN_STATS=/tmp/p5/stats.txt B2D_DROPIN_TIMING=1 B2D_DROPIN_OPTIONS="${P5_OPTIONS:-factorised=1}" \
    ${STOP_AFTER:+B2D_DROPIN_STOP_AFTER=$STOP_AFTER} stdbuf -oL -eL $R/oracle/_ref/block_gpu dmrg.conf > $O/stdout.txt 2> $O/stderr.txt &
PID=$!
# watchdog: wall-clock cap and available host memory (GB)
while kill -0 $PID 2>/dev/null; do
  sleep 5
  NOW=$(date +%s.%N)
  AVAIL=$(awk '/MemAvailable/ {print int($2/1048576)}' /proc/meminfo)
  if [ "$(python -c "print(int($NOW - $T0 > $CAP_S))")" = "1" ]; then echo "watchdog: wall-clock cap $CAP_S s reached" >> $O/stderr.txt; kill $PID; break; fi
  if [ "$AVAIL" -lt 16 ]; then echo "watchdog: host memory available $AVAIL GB < 16 GB" >> $O/stderr.txt; kill $PID; break; fi
done
wait $PID; RC=$?
T1=$(date +%s.%N)
cd $R
{ grep -E "Sweep Energy|Elapsed Sweep Wall" $O/stdout.txt; echo "exit $RC, total wall $(python -c "print('%.1f' % ($T1 - $T0))") s"; grep -E "B2D_TIMING|watchdog|STOP_AFTER" $O/stderr.txt; grep B2D_PROGRESS $O/stderr.txt | tail -3; } | tee $O/sweeps.txt
cp /tmp/p5/stats.txt $O/stats.txt 2>/dev/null
# keep the log small: block-iteration headers only
grep -E "Block Iteration|# states|Sweep Energy|Elapsed|Davidson|watchdog" $O/stdout.txt | cut -c1-200 > $O/stdout_short.txt; rm -f $O/stdout.txt
tail -c 20000 $O/stderr.txt > $O/stderr_tail.txt; rm -f $O/stderr.txt
python - <<'PY'
import re
rows = [dict((k, float(v)) for k, v in re.findall(r"(\w+)=([-\d.e+]+)", l)) for l in open("gpurun_out/p5/stats.txt")]
keys = ['n_multiply', 'upload_s', 'diag_s', 'davidson_s', 'density_s', 'eig_s', 'rotate_s', 'launches', 'guess_s', 'cache_uses']
print(len(rows), "block iterations;", {k: round(sum(r.get(k, 0) for r in rows), 2) for k in keys})
rows.sort(key=lambda r: -r.get('sigma_flops', 0))
for r in rows[:6]:
    print("lsites %d W %d sigma_flops %.3e n_multiply %d davidson_dev_ms %.1f -> %.0f GFLOP/s algorithmic; upload %.2f eig %.2f rotate %.2f s" % (
        r['lsites'], r['W'], r['sigma_flops'], r['n_multiply'], r['davidson_dev_ms'], r['sigma_flops'] * r['n_multiply'] / max(r['davidson_dev_ms'], 1e-9) / 1e6,
        r['upload_s'], r['eig_s'], r['rotate_s']))
PY
