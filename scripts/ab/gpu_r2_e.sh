mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tail_golden.py -m gpu -q -x -k "detect or tail or golden or jaccard" 2>&1 | tail -3
for occ in 3 2; do
CROG_SCAN_OCC=$occ python bench.py --workload tail --steps 20 > gpurun_out/r2_tail_occ$occ.json 2>gpurun_out/r2_tail_occ$occ.err; python -c "import json;d=json.loads(open('gpurun_out/r2_tail_occ$occ.json').read().strip().splitlines()[-1]);print('OCC',$occ,'blobs',round(d['blobs']['ms'],3),round(d['blobs']['detect_ms'],3),d['blobs']['parity_spot_check'],'stress',round(d['stress']['ms'],3),round(d['stress']['detect_ms'],3),d['stress']['parity_spot_check'])"
done
