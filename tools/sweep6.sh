B="python bench.py --no-cpu-baseline --no-e2e --steps 100 --content natural"
run(){ name=$1; shift; "$@" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'var':'$name','fps':d['value'],'kernel':d['roofline']['kernel'],'parity':d['config']['parity_spot_check']}))"; }
for r in 64 96 128 192; do for g in 8880 17760 1000000; do run warp_r${r}_g$g $B --opt wk_tma=0 --rows-per-cta $r --opt grid_ctas=$g; done; done
for r in 64 128; do for g in 444 8880 17760; do run tma_r${r}_g$g $B --rows-per-cta $r --opt grid_ctas=$g; done; done
run tma_copy_r64_g17760 $B --rows-per-cta 64 --opt grid_ctas=17760 --opt wk_var=1000
