#!/bin/bash
# generic-kernel throughput on a few cells of the format matrix (4K geometry of cfg3, 16 frames per step)
out=gpurun_out/bench_matrix.jsonl; : > $out
run() { python bench.py --config cfg3 --steps 5 --warmup 3 --frames-per-step 16 --no-cpu-baseline --no-e2e --no-classes --min-seconds 0.3 "$@" >> $out 2>> gpurun_out/bench_matrix.err; }
run --in-fmt yuv444p10le --out-fmt yuv444p10le
run --in-fmt yuv422p10le --out-fmt yuv422p
run --in-fmt yuv444p --out-fmt yuv444p
run --in-fmt yuv420p --out-fmt yuv420p10le
run --in-fmt yuv420p --out-fmt yuv444p
run --in-fmt yuv422p --out-fmt yuv422p
run --in-fmt yuv420p10le --out-fmt yuv420p --opt any_kernel=1
run --in-fmt yuv422p10le --out-fmt yuv422p10le --opt any_kernel=1
run --in-fmt yuv422p10le --out-fmt yuv422p10le --interp prism
run --in-fmt yuv420p --out-fmt yuv420p --opt any_kernel=1
python - <<'PY'
import json
for l in open("gpurun_out/bench_matrix.jsonl"):
    d = json.loads(l); print(d["metric"], d["value"], d["gpix_per_s"], d["roofline"]["kernel"], d["parity"]["spot_check"], d["config"].get("options"))
PY
tail -3 gpurun_out/bench_matrix.err
