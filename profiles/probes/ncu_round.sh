# Round evidence: launch list of the default bench command (gpu__time_duration per launch) and
# one `--set full` capture each of the three dominant kernels.  Every ncu pass runs only after the
# same command has exited 0 without ncu.  usage: ncu_round.sh <tag>
TAG="$1"
B="python bench.py --steps 2 --warmup 1 --skip-e2e --skip-cpu --parity-gaussians 1000 --parity-rows 1000"
$B > gpurun_out/plain_$TAG.json 2>gpurun_out/plain_$TAG.err || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
for K in lift_gather_kernel lift_majority_kernel kmeans_step_tc_kernel; do
  ncu --set full --clock-control none --import-source on -k "regex:$K" -s 1 -c 1 -f -o gpurun_out/prof_${K}_$TAG $B > gpurun_out/ncu_${K}_$TAG.log 2>&1
  echo "$K rc=$?"
done
ls -la gpurun_out/*_$TAG.ncu-rep gpurun_out/launches_$TAG.csv
