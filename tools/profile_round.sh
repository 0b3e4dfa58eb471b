#!/bin/bash
# ncu evidence of one round (run under gpurun, one GPU): per-launch times of the default bench command, and one `--set full`
# capture of the dominant kernel of each named config (the launch after the warm-up ones).  Outputs land in gpurun_out/.
#   bash tools/profile_round.sh r01v7 cfg3 deliver8 hevc10
tag=$1; shift
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_launches.log 2>&1
for cfg in "$@"; do
  extra=""
  case $cfg in any444) cfgname=cfg3; extra="--in-fmt yuv444p10le --out-fmt yuv444p10le --frames-per-step 16";; *) cfgname=$cfg;; esac
  ncu --set full --clock-control none --import-source on --launch-skip 4 -c 1 -k regex:'k422_warp|k420_warp|k10_warp|k_any' \
      -o gpurun_out/${tag}_${cfg} -f python bench.py --config $cfgname $extra --steps 2 --warmup 3 --no-cpu-baseline --no-e2e \
      > gpurun_out/${tag}_${cfg}_ncu.log 2>&1
  # summarise on the box (gpurun brings back at most 64 MiB): keep the report of the first config only
  python tools/ncu_summary.py gpurun_out/${tag}_${cfg}.ncu-rep gpurun_out/${tag}_${cfg}_ncu_full.json $([ $cfgname = cfg1 ] && echo 64 || ([ "$extra" ] && echo 16 || ([ $cfgname = cfg4 ] && echo 16 || echo 64))) > /dev/null
  [ "$cfg" = "$1" ] || rm -f gpurun_out/${tag}_${cfg}.ncu-rep
done
