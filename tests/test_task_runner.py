"""TaskRunner keeps the reference's status / progress / cancel / clean-up behaviour (task_manager.py:46-132); the ffmpeg
child processes are faked here (no ffmpeg binary exists in the image) and stages without a LUT never touch the GPU."""
import io
from pathlib import Path

from lut_renderer_b200 import ProcessingParams, Task, VideoInfo
from lut_renderer_b200.task_manager import TaskCallbacks, TaskRunner


class Rec(TaskCallbacks):
    def __init__(self):
        self.events = []

    def progress(self, tid, pct):
        self.events.append(("progress", pct))

    def status(self, tid, text):
        self.events.append(("status", text))

    def finished(self, tid, st):
        self.events.append(("finished", st))

    def log(self, tid, msg):
        self.events.append(("log", msg))


class FakeProc:
    def __init__(self, lines, rc):
        self.stdout = io.StringIO("".join(lines))
        self._rc = rc
        self.returncode = None

    def poll(self):
        return self.returncode

    def wait(self):
        self.returncode = self._rc
        return self._rc

    def terminate(self):
        self.returncode = -15

    def kill(self):
        self.returncode = -9


def _task(tmp_path, **kw):
    return Task("t1", tmp_path / "in.mov", tmp_path / "out.mp4", None, None, ProcessingParams(**kw), source_info=VideoInfo(fps=25.0))


def test_plain_stage_progress_and_completion(tmp_path):
    rec = Rec()
    lines = ["Duration: 00:00:10.00, start\n", "frame=1 time=00:00:02.50 x\n", "frame=2 time=00:00:05.00 x\n", "time=00:00:10.00\n"]
    cmds = []
    r = TaskRunner(_task(tmp_path), callbacks=rec, popen=lambda cmd, **k: (cmds.append(cmd), FakeProc(lines, 0))[1])
    r.run()
    assert [e for e in rec.events if e[0] == "progress"] == [("progress", 25), ("progress", 50), ("progress", 100), ("progress", 100)]
    assert ("status", "running") in rec.events and rec.events[-1] == ("finished", "completed")
    assert cmds[0][0] == "ffmpeg" and "-vf" not in cmds[0]


def test_failure_reports_exit_code(tmp_path):
    rec = Rec()
    TaskRunner(_task(tmp_path), callbacks=rec, popen=lambda cmd, **k: FakeProc(["boom\n"], 3)).run()
    assert ("status", "failed: exit 3") in rec.events and rec.events[-1] == ("finished", "failed")


def test_exception_becomes_failed_status(tmp_path):
    rec = Rec()
    t = _task(tmp_path, processing_mode="pro")      # pro mode without intermediate_path raises inside build_pipeline
    TaskRunner(t, callbacks=rec).run()
    assert any(e[0] == "status" and e[1].startswith("failed: ") for e in rec.events) and rec.events[-1] == ("finished", "failed")


def test_cancel_before_start(tmp_path):
    rec = Rec()
    r = TaskRunner(_task(tmp_path), callbacks=rec, popen=lambda cmd, **k: FakeProc(["x\n"], 0))
    r.cancel()
    r.run()
    assert rec.events[-1] == ("finished", "completed") or rec.events[-1] == ("finished", "canceled")


# ---- the pipelined GPU pump (reader thread | stage | writer thread) with the stage itself replaced by a byte-flipping stand-in: the
# ---- threading, pipe, cancel, failure and sharding logic is host code and must hold without a GPU ---------------------------------
class _FakeCtx:
    def lut_from_file(self, path):
        return object()


class _FakeStage:
    """Stand-in for ColourStage on a box without a GPU: same `pump` contract (read whole frames from fd_in until EOF, write the
    processed frames to fd_out, poll the cancel flag, raise on a broken pipe), the 'colour stage' being a byte flip."""

    def __init__(self, plan):
        self.ctx = _FakeCtx()
        self.in_frame_bytes = self.out_frame_bytes = plan.width * plan.height * 3 // 2

    @classmethod
    def from_plan(cls, plan, device=0, float_mode="c"):
        return cls(plan)

    def process(self, data):
        import numpy as np
        return (255 - np.frombuffer(data, np.uint8)).tobytes()

    def pump(self, lut, fd_in, fd_out, frames_per_batch=0, cancel=None, on_frames=None):
        import os
        import select
        from lut_renderer_b200._native import ERR_CANCELLED, ERR_IO, LutrError
        fb, nb, total, pending = self.in_frame_bytes, max(1, frames_per_batch), 0, b""
        while True:
            if cancel is not None and cancel.value:
                raise LutrError(ERR_CANCELLED, "cancelled")
            if not select.select([fd_in], [], [], 0.1)[0]:
                continue
            chunk = os.read(fd_in, nb * fb - len(pending))
            pending += chunk
            if chunk and len(pending) < nb * fb:
                continue
            n = len(pending) // fb
            if n:
                out = memoryview(self.process(pending[:n * fb]))
                while out:
                    if cancel is not None and cancel.value:
                        raise LutrError(ERR_CANCELLED, "cancelled")
                    if not select.select([], [fd_out], [], 0.1)[1]:
                        continue
                    try:
                        out = out[os.write(fd_out, out[:1 << 16]):]
                    except OSError as exc:
                        raise LutrError(ERR_IO, f"write to the encoder pipe failed: {exc}")
                total += n
                pending = pending[n * fb:]
                if on_frames:
                    on_frames(total)
            if not chunk:
                return {"frames_in": total, "frames_out": total}

    def close(self):
        pass


def _patched_runner(monkeypatch, tmp_path, popen, duration=4.0, devices=None, batch=4):
    import lut_renderer_b200.stage as stage_mod
    monkeypatch.setattr(stage_mod, "ColourStage", _FakeStage)
    info = VideoInfo(width=64, height=32, pix_fmt="yuv420p", bit_depth=8, colorspace="bt709", color_range="tv", fps=10.0, duration=duration)
    task = Task("t", tmp_path / "in.mov", tmp_path / "out.mp4", tmp_path / "l.cube", None, ProcessingParams(), source_info=info)
    rec = Rec()
    return TaskRunner(task, callbacks=rec, frames_per_batch=batch, popen=popen, devices=devices), rec


def test_pump_streams_frames_in_order_and_shards(monkeypatch, tmp_path):
    import math
    import subprocess
    import sys
    import numpy as np
    fb, n, fps = 64 * 32 * 3 // 2, 43, 10.0
    data = np.random.default_rng(0).integers(0, 256, n * fb, dtype=np.uint8)
    (tmp_path / "in.raw").write_bytes(data.tobytes())

    def popen(cmd, **kw):
        if "pipe:1" in cmd:
            # accurate seek: the first frame whose pts is not before the target (the runner seeks half a frame early)
            start = math.ceil(float(cmd[cmd.index("-ss") + 1]) * fps - 1e-6) if "-ss" in cmd else 0
            count = int(cmd[cmd.index("-frames:v") + 1]) if "-frames:v" in cmd else n - start
            code = f"import sys; f=open({str(tmp_path / 'in.raw')!r},'rb'); f.seek({start * fb}); sys.stdout.buffer.write(f.read({count * fb}))"
        elif "pipe:0" in cmd:
            code = f"import sys, shutil; shutil.copyfileobj(sys.stdin.buffer, open({cmd[-1]!r}, 'wb'))"
        else:
            lst = cmd[cmd.index("-i") + 1]
            code = ("parts = [l.strip()[6:-1] for l in open(%r) if l.strip()]\n"
                    "open(%r, 'wb').write(b''.join(open(p, 'rb').read() for p in parts))") % (lst, cmd[-1])
        return subprocess.Popen([sys.executable, "-c", code], **kw)

    for devices in (None, [0, 1, 2]):
        r, rec = _patched_runner(monkeypatch, tmp_path, popen, duration=n / fps, devices=devices)
        r.run()
        assert rec.events[-1] == ("finished", "completed"), rec.events[-4:]
        out = np.frombuffer((tmp_path / "out.mp4").read_bytes(), np.uint8)
        assert out.size == n * fb and np.array_equal(out, 255 - data), devices
        assert ("progress", 100) in rec.events
        (tmp_path / "out.mp4").unlink()


def test_pump_cancel_and_encoder_failure_do_not_hang(monkeypatch, tmp_path):
    import subprocess
    import sys
    import threading
    import time
    fb = 64 * 32 * 3 // 2
    endless = f"import sys\nb = bytes({fb})\nwhile True: sys.stdout.buffer.write(b)"
    sink = "import sys\nwhile sys.stdin.buffer.read(1 << 20): pass"
    dies = "import sys\nsys.stdin.buffer.read(1 << 16)\nsys.exit(3)"

    def mk(encoder):
        return lambda cmd, **kw: subprocess.Popen([sys.executable, "-c", endless if "pipe:1" in cmd else encoder], **kw)

    r, rec = _patched_runner(monkeypatch, tmp_path, mk(sink), duration=1e6)
    threading.Timer(1.0, r.cancel).start()
    t0 = time.perf_counter()
    r.run()
    assert rec.events[-1] == ("finished", "canceled") and time.perf_counter() - t0 < 30, rec.events[-3:]
    r, rec = _patched_runner(monkeypatch, tmp_path, mk(dies), duration=1e6)
    t0 = time.perf_counter()
    r.run()
    assert rec.events[-1] == ("finished", "failed") and time.perf_counter() - t0 < 30, rec.events[-3:]
    assert any(e[0] == "status" and e[1].startswith("failed") for e in rec.events)


def test_short_shard_is_refused_instead_of_stitched(monkeypatch, tmp_path):
    """A part that delivers fewer frames than its planned range (a seek that dropped a frame, a truncated source) would lose frames
    at the cut: the runner fails the task instead of concatenating (frame counts are checked before the concat)."""
    import math
    import subprocess
    import sys
    fb, n, fps = 64 * 32 * 3 // 2, 40, 10.0
    (tmp_path / "in.raw").write_bytes(bytes(n * fb))

    def popen(cmd, **kw):
        if "pipe:1" in cmd:
            start = math.ceil(float(cmd[cmd.index("-ss") + 1]) * fps - 1e-6) if "-ss" in cmd else 0
            count = int(cmd[cmd.index("-frames:v") + 1]) - 1 if "-frames:v" in cmd else n - start      # every closed range is one short
            code = f"import sys; f=open({str(tmp_path / 'in.raw')!r},'rb'); f.seek({start * fb}); sys.stdout.buffer.write(f.read({count * fb}))"
        else:
            code = "import sys\nwhile sys.stdin.buffer.read(1 << 20): pass"
        return subprocess.Popen([sys.executable, "-c", code], **kw)

    r, rec = _patched_runner(monkeypatch, tmp_path, popen, duration=n / fps, devices=[0, 1])
    r.run()
    assert rec.events[-1] == ("finished", "failed"), rec.events[-4:]
    assert any(e[0] == "log" and "帧数不符" in e[1] for e in rec.events)


def test_stage_failure_kills_the_children_and_fails_the_task(monkeypatch, tmp_path):
    """run_host raising (a CUDA error, an unsupported stride) must not leave the decoder blocked on a full pipe and the task
    'running' forever: the children are killed and the task reports failed: <text>."""
    import subprocess
    import sys
    import time
    import lut_renderer_b200.stage as stage_mod
    fb = 64 * 32 * 3 // 2
    endless = f"import sys\nb = bytes({fb})\nwhile True: sys.stdout.buffer.write(b)"
    sink = "import sys\nwhile sys.stdin.buffer.read(1 << 20): pass"

    class Boom(_FakeStage):
        calls = 0

        def process(self, data):
            Boom.calls += 1
            if Boom.calls == 2:
                raise RuntimeError("CUDA error: injected")
            return super().process(data)

    procs = []

    def popen(cmd, **kw):
        p = subprocess.Popen([sys.executable, "-c", endless if "pipe:1" in cmd else sink], **kw)
        procs.append(p)
        return p

    r, rec = _patched_runner(monkeypatch, tmp_path, popen, duration=1e6)
    monkeypatch.setattr(stage_mod, "ColourStage", Boom)
    t0 = time.perf_counter()
    r.run()
    assert time.perf_counter() - t0 < 30
    assert rec.events[-1] == ("finished", "failed") and any(e[0] == "status" and "injected" in e[1] for e in rec.events), rec.events[-3:]
    assert all(p.poll() is not None for p in procs)


def test_vfr_passthrough_source_keeps_the_reference_command(monkeypatch, tmp_path):
    """rawvideo has no timestamps: a VFR source in passthrough mode cannot cross the pipe, the stage runs as the reference's own
    single ffmpeg command (with the lut3d filter) and the log says so."""
    cmds = []
    info = VideoInfo(width=64, height=32, pix_fmt="yuv420p", bit_depth=8, colorspace="bt709", color_range="tv", fps=25.0, duration=1.0, is_vfr=True)
    task = Task("t", tmp_path / "in.mov", tmp_path / "out.mp4", tmp_path / "l.cube", None, ProcessingParams(force_cfr=False), source_info=info)
    rec = Rec()
    TaskRunner(task, callbacks=rec, popen=lambda cmd, **k: (cmds.append(cmd), FakeProc(["time=00:00:01.00\n"], 0))[1]).run()
    assert rec.events[-1] == ("finished", "completed")
    assert len(cmds) == 1 and any("lut3d" in tok for tok in cmds[0]) and "pipe:1" not in cmds[0]
    assert any(e[0] == "log" and "GPU 色彩阶段不适用" in e[1] for e in rec.events)
