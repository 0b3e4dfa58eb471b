B="python bench.py --no-cpu-baseline --no-e2e --steps 100"
run(){ name=$1; shift; for ct in natural noise; do "$@" --content $ct | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'var':'$name','content':'$ct','fps':d['value'],'kernel':d['roofline']['kernel'],'parity':d['config']['parity_spot_check']}))"; done; }
LUTR_LIB=$PWD/lut-renderer_b200/liblutr_r1.so run r1 $B
run warp $B --opt wk_tma=0
run tma $B
run tma_copy $B --opt wk_var=1000
