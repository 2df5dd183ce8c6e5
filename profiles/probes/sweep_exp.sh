# lifting-only bench runs for kernel experiments (no CPU legs): prints ms per step and per phase
B="python bench.py --steps 5 --warmup 2 --skip-kmeans --skip-e2e --skip-cpu"
pick() { python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('$1', round(d['ms_per_step'],3), d['kernels_ms'])"; }
$B 2>/dev/null | pick default
for c in "$@"; do env $c $B 2>/dev/null | pick "$c"; done
