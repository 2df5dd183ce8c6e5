# One `ncu --set full` capture of one kernel of the bench command (source view on), after the same
# command has exited 0 without ncu.  usage: ncu_one.sh <kernel regex> <out name> [bench flags...]
K="$1"; OUT="$2"; shift 2
B="python bench.py --steps 1 --warmup 1 --skip-e2e --skip-cpu --parity-gaussians 1000 --parity-rows 1000 $*"
$B > /dev/null 2>gpurun_out/ncu_one_err.log || { echo "plain run failed"; tail -5 gpurun_out/ncu_one_err.log; exit 1; }
ncu --set full --clock-control none --import-source on -k "regex:$K" -s 1 -c 1 -f -o gpurun_out/$OUT $B > gpurun_out/ncu_one.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/$OUT.ncu-rep
