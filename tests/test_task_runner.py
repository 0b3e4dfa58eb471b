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
