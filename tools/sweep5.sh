B="python bench.py --no-cpu-baseline --no-e2e --steps 100 --content natural"
run(){ name=$1; shift; "$@" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'var':'$name','fps':d['value'],'kernel':d['roofline']['kernel'],'parity':d['config']['parity_spot_check']}))"; }
for r in 24 32 48 64; do for g in 1776 4440 8880 17760 1000000; do run warp_r${r}_g$g $B --opt wk_tma=0 --rows-per-cta $r --opt grid_ctas=$g; done; done
for r in 32 64; do for g in 4440 8880 1000000; do run tma_r${r}_g$g $B --rows-per-cta $r --opt grid_ctas=$g; done; done
