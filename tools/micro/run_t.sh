set -x
for t in 0 2 3; do echo "== PNO_MIX_TRUNC=$t"; PNO_MIX_TRUNC=$t python tools/kernel_times.py tf32x3 | tail -1 | python -c "import sys,json,re; l=sys.stdin.read(); d=json.loads(l[l.index('{'):l.rindex('}')+1]); print({k:v['us'] for k,v in d.items()})"; done
for t in 0 2 3; do PNO_MIX_TRUNC=$t timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -k "test_model_tensor_core_engines_against_oracle" 2>&1 | grep -E "passed|failed|assert .* <" ; done
timeout 900 python -m pytest tests -m gpu -x -q -k "dw or step or layer or cfg2 or cfg3 or model" 2>&1 | tail -3
