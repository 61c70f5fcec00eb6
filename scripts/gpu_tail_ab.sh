# tail tests + tail micro-bench (blobs / stress) + per-kind timing; compute-sanitizer racecheck skipped (too slow)
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tail_golden.py -m gpu -q -k "detect or tail or peak or jaccard or golden or plateau or tie" 2>&1 | tail -2
timeout 300 python scripts/time_tail.py
for i in 1; do timeout 300 python bench.py --workload tail 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); t=d.get('tail',d)
for k in ('blobs','stress'): print(k, 'ms %.3f frac %.3f detect_ms %.3f'%(t[k]['ms'],t[k]['frac_of_hbm'],t[k]['detect_ms']), t[k]['parity_spot_check'])"; done
