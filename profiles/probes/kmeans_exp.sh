# k-means-only timing for kernel experiments: iterations/s and ms per iteration, parity fields
B="python bench.py --steps 10 --warmup 3 --skip-e2e --gaussians 200000 --views 16 --parity-gaussians 1000"
pick() { python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); k=d['kmeans']; print('$1', 'ms/iter', round(k['ms_per_iter'],4), 'frac', round(k['roofline']['frac'],3), 'ordered', k.get('ordered_ms_per_iter'), {a:b for a,b in d['parity']['kmeans'].items() if a in ('labels_match','rel_err_vs_exact_mean','ordered_bit_exact')})"; }
$B 2>gpurun_out/k_err.log | pick default
for c in "$@"; do env $c $B 2>/dev/null | pick "$c"; done
