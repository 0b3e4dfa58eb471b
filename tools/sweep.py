#!/usr/bin/env python
"""A/B runs of bench.py on one GPU box: every case is one bench.py process (device-resident measurement only), one compact JSON
line per case -- name, frames/s, kernel, parity spot check.

    python tools/sweep.py gpurun_out/out.jsonl "name|bench.py arguments" ["name|LUTR_LIB=other.so bench.py arguments" ...]
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    out, cases = sys.argv[1], sys.argv[2:]
    with open(out, "w") as f:
        for case in cases:
            name, _, rest = case.partition("|")
            env = dict(os.environ)
            words = rest.split()
            while words and "=" in words[0] and not words[0].startswith("-"):
                k, _, v = words.pop(0).partition("=")
                env[k] = v
            cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline", "--no-e2e", "--no-classes", "--steps", "20", "--min-seconds", "0.5"] + words
            res = subprocess.run(cmd, capture_output=True, text=True, env=env)
            try:
                d = json.loads(res.stdout.strip().splitlines()[-1])
                parity = d.get("parity", {}).get("spot_check") if "parity" in d else d["config"].get("parity_spot_check")
                line = {"case": name, "fps": d["value"], "gpix_per_s": d["gpix_per_s"], "kernel": d["roofline"]["kernel"], "frac": d["roofline"]["frac"], "parity": parity}
            except Exception:
                line = {"case": name, "error": (res.stderr or res.stdout)[-300:]}
            print(json.dumps(line), flush=True)
            f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
