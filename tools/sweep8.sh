B="python bench.py --no-cpu-baseline --no-e2e --steps 50"
run(){ name=$1; shift; "$@" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'var':'$name','fps':d['value'],'kernel':d['roofline']['kernel'],'parity':d['config']['parity_spot_check']}))"; }
for cfg in cfg1 cfg2 master8 master10 hevc10 deliver8; do
  run ${cfg}_default $B --config $cfg
  for g in 888 1776 3552 7104; do run ${cfg}_g$g $B --config $cfg --opt grid_ctas=$g; done
done
