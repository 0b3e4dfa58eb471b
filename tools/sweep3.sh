B="python bench.py --no-cpu-baseline --no-e2e --steps 100 --content natural"
run(){ name=$1; shift; "$@" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'var':'$name','fps':d['value'],'kernel':d['roofline']['kernel'],'parity':d['config']['parity_spot_check']}))"; }
for c in 0 15 29 44 58; do run warp_carve$c $B --opt wk_tma=0 --opt wk_carveout=$c; done
for c in 0 35 44 58; do run tma_carve$c $B --opt wk_carveout=$c; done
