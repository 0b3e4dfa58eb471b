B="python bench.py --no-cpu-baseline --no-e2e --steps 100 --content natural"
run(){ name=$1; shift; "$@" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'var':'$name','fps':d['value'],'kernel':d['roofline']['kernel'],'parity':d['config']['parity_spot_check']}))"; }
for g in 444 888 1776 4440 8880 34816; do run warp_grid$g $B --opt wk_tma=0 --opt grid_ctas=$g; done
for g in 444 1776 4440 34816; do run tma_grid$g $B --opt grid_ctas=$g; done
for r in 8 16 32; do run warp_rows$r $B --opt wk_tma=0 --rows-per-cta $r --opt grid_ctas=8880; done
