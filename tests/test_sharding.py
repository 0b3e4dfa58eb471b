"""Frame-range / clip sharding (no collective on the data path) -- host logic, plus a world_size-2 gloo run."""
import os
import socket
import subprocess
import sys

from lut_renderer_b200.sharding import Clip, assign_clips, frame_range, gop_aligned_ranges, luts_per_rank

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frame_ranges_tile_the_clip():
    for n in (0, 1, 7, 64, 1001):
        for world in (1, 2, 3, 4, 8):
            rs = [frame_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1


def test_gop_aligned():
    rs = gop_aligned_ranges(1000, 8, 24)
    assert rs[0][0] == 0 and rs[-1][1] == 1000 and all(a % 24 == 0 for a, _ in rs)
    assert all(x[1] == y[0] for x, y in zip(rs, rs[1:]))


def test_clip_assignment_cfg5():
    clips = [Clip(f"c{i}", 3840 if i % 2 else 1920, 2160 if i % 2 else 1080, "yuv422p10le" if i % 2 else "yuv420p", 120, f"look{i % 5}.cube")
             for i in range(64)]
    plan = assign_clips(clips, 8)
    assert sorted(i for r in plan for i in r) == list(range(64))
    loads = [sum(clips[i].cost for i in r) for r in plan]
    assert max(loads) / min(loads) < 1.05
    assert plan == assign_clips(clips, 8)                      # deterministic
    per = luts_per_rank(clips, plan)
    assert all(sum(d.values()) == len(r) for d, r in zip(per, plan))


WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import torch, torch.distributed as dist
from lut_renderer_b200.sharding import frame_range
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
total = 101
a, b = frame_range(total, rank, world)
# every rank "processes" its own frames (independent work), then the harness reduces counts and the max time
done = torch.tensor([b - a], dtype=torch.int64)
dist.all_reduce(done)
t = torch.tensor([0.01 * (rank + 1)], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
owners = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
dist.all_gather(owners, torch.tensor([a, b], dtype=torch.int64))
if rank == 0:
    assert int(done) == total, done
    assert abs(float(t) - 0.01 * world) < 1e-9
    spans = [tuple(int(x) for x in o) for o in owners]
    assert spans[0][0] == 0 and spans[-1][1] == total and all(x[1] == y[0] for x, y in zip(spans, spans[1:]))
    print("SHARDING_OK", spans)
dist.destroy_process_group()
'''


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "SHARDING_OK" in res.stdout


def test_bench_reference_arm_runs_only_on_rank0():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=120)
    assert res.returncode == 0 and res.stdout.strip() == ""
