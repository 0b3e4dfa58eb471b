#!/bin/bash
# A/B sweep of k422_warp variants (experiment builds: LUTR_WK_EXPERIMENTS=1): wk_var = 100 * BM + VAR, see csrc/k422_warp.cu
out=${1:-gpurun_out/var_sweep.jsonl}; shift
: > $out
B="python bench.py --no-cpu-baseline --no-e2e --steps 100"
for v in "$@"; do
  for ct in natural noise; do
    if [ "$v" = r1 ]; then LUTR_LIB=$PWD/lut-renderer_b200/liblutr_r1.so $B --content $ct | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'var':'r1','content':'$ct','fps':d['value'],'parity':d['config']['parity_spot_check']}))" >> $out
    elif [ "$v" = default ]; then $B --content $ct | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'var':'default','content':'$ct','fps':d['value'],'parity':d['config']['parity_spot_check']}))" >> $out
    else $B --content $ct --opt wk_var=$v | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'var':$v,'content':'$ct','fps':d['value'],'parity':d['config']['parity_spot_check']}))" >> $out; fi
  done
done
cat $out
