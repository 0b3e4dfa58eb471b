#!/bin/bash
# ncu evidence of one round (run under gpurun, one GPU): per-launch times of the default bench command, and one `--set full`
# capture of the dominant kernel of each named config (a launch after the warm-up ones).  Outputs land in gpurun_out/; the
# condensed summaries and profiles/traffic.json (DRAM bytes per frame, read by bench.py for roofline.traffic) are refreshed from them.
#   bash tools/profile_round.sh r02v5 cfg3 cfg1 cfg2 master10 hevc10 deliver8
tag=$1; shift
B="--no-cpu-baseline --no-e2e --no-classes --steps 2 --warmup 3 --min-seconds 0.01"
python bench.py $B > gpurun_out/${tag}_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py $B > gpurun_out/${tag}_ncu_launches.log 2>&1
for cfg in "$@"; do
  extra=""
  case $cfg in any444) cfgname=cfg3; extra="--in-fmt yuv444p10le --out-fmt yuv444p10le --frames-per-step 16";; any422) cfgname=cfg3; extra="--opt any_kernel=1 --frames-per-step 16";; *) cfgname=$cfg;; esac
  frames=$([ "$extra" ] && echo 16 || ([ $cfgname = cfg4 ] && echo 16 || ([ $cfgname = cfg1 -o $cfgname = cfg2 ] && echo 256 || echo 64)))
  python bench.py --config $cfgname $extra $B > gpurun_out/${tag}_${cfg}_plain.log 2>&1 || continue
  ncu --set full --clock-control none --import-source on --launch-skip 5 -c 1 -k regex:'k422|k420_warp|k10_warp|k_any|k_convert' \
      -o gpurun_out/${tag}_${cfg} -f python bench.py --config $cfgname $extra $B > gpurun_out/${tag}_${cfg}_ncu.log 2>&1
  python tools/ncu_summary.py gpurun_out/${tag}_${cfg}.ncu-rep gpurun_out/${tag}_${cfg}_ncu_full.json $frames > /dev/null
  case $cfgname in cfg1|cfg2) px=$((frames * 1920 * 1080));; cfg4) px=$((frames * 7680 * 4320));; *) px=$((frames * 3840 * 2160));; esac
  python tools/opcode_mix.py gpurun_out/${tag}_${cfg}.ncu-rep $px > gpurun_out/${tag}_${cfg}_opcode_mix.txt 2>/dev/null
  # gpurun brings back at most 64 MiB: keep the report of the first config only
  [ "$cfg" = "$1" ] || rm -f gpurun_out/${tag}_${cfg}.ncu-rep
done
python - "$tag" "$@" <<'PY'
import json, sys
tag, cfgs = sys.argv[1], sys.argv[2:]
out = {}
for c in cfgs:
    try:
        d = json.load(open(f"gpurun_out/{tag}_{c}_ncu_full.json"))[0]
        out[c] = {"dram_bytes_per_frame": d["dram_bytes_per_frame"], "kernel": d["kernel"],
                  "source": f"profiles/{tag}_{c}_ncu_full.json (ncu --set full of `bench.py --config {c}`, one launch; tools/profile_round.sh)"}
    except Exception as e:
        print("no summary for", c, e)
json.dump(out, open(f"gpurun_out/{tag}_traffic.json", "w"), indent=1)
PY
