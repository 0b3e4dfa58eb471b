B="python bench.py --no-cpu-baseline --no-e2e --steps 100"
run(){ name=$1; shift; "$@" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'var':'$name','fps':d['value'],'kernel':d['roofline']['kernel'],'parity':d['config']['parity_spot_check']}))"; }
for ct in natural noise; do
run warp_$ct $B --content $ct --opt wk_tma=0 --rows-per-cta 128 --opt grid_ctas=8880
run bulk_$ct $B --content $ct --rows-per-cta 128 --opt grid_ctas=8880 --opt wk_var=1002
run cpasync_$ct $B --content $ct --rows-per-cta 128 --opt grid_ctas=8880 --opt wk_var=1003
done
run copy_bulk $B --rows-per-cta 128 --opt grid_ctas=8880 --opt wk_var=1000
run copy_cpasync $B --rows-per-cta 128 --opt grid_ctas=8880 --opt wk_var=1001
run copy_bulk_g444 $B --opt wk_var=1000
run copy_cpasync_g444 $B --opt wk_var=1001
