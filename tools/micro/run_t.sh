timeout 300 python -m pytest tests/test_gpu_tc.py -x -q -k "dft_fwd or transform_pair" 2>&1 | tail -5
for n in 1 0; do echo "== NCAT=$n"; PNO_FWD_NCAT=$n timeout 120 python tools/kernel_times.py tf32x3 | tail -1 | python -c "import sys,json; l=sys.stdin.read(); d=json.loads(l[l.index('{'):l.rindex('}')+1]); print({k:v['us'] for k,v in d.items()})"; done
