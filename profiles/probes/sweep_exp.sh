# lifting-only bench runs for kernel experiments (no CPU legs): prints ms per step and per phase
B="python bench.py --steps 5 --warmup 2 --skip-kmeans --skip-e2e --skip-cpu"
pick() { python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); k=d['kernels_ms']; print('$1', 'step', round(d['ms_per_step'],3), [round(v,3) for v in k.values() if isinstance(v,float)])"; }
$B 2>/dev/null | pick default
for c in "$@"; do env $c $B 2>/dev/null | pick "$c"; done
