timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -k "dft_inv or inverse or transform_pair or full_size" 2>&1 | tail -3
timeout 120 python tools/kernel_times.py tf32 tf32x3 | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l[l.index('{'):l.rindex('}')+1]); print(l.split()[0], {k:v['us'] for k,v in d.items()})"
