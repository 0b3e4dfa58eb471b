"""Operator surface -> pixels: ProcessingParams + VideoInfo exactly as the reference's GUI / TaskRunner would hold them
(reference models.py:19-56, media_info.py:12-36) -> plan_colour_stage -> ColourStage.from_plan -> GPU, checked against the
oracle driven by the same plan.  One case per decision branch of build_command's colour block (ffmpeg.py:196-247, 287-310,
348-383)."""
from pathlib import Path

import numpy as np
import pytest

from lut_renderer_b200 import ProcessingParams, VideoInfo, plan_colour_stage, synth

from helpers import assert_bit_exact, oracle_lut, oracle_stage

pytestmark = pytest.mark.gpu

W, H = 264, 40


def _info(pix_fmt, depth, colorspace="bt709", color_range="tv"):
    return VideoInfo(width=W, height=H, pix_fmt=pix_fmt, bit_depth=depth, colorspace=colorspace, color_range=color_range, fps=25.0)


CASES = {
    # fast delivery, defaults: 8-bit source, libx264, nothing forced -> the encoder negotiates yuv444p behind lut3d's RGB
    # (SURVEY App. A cfg 1); an encoder outside the negotiation table gets the working layout back
    "default_8bit": (_info("yuv420p", 8), ProcessingParams(), ("yuv420p", "yuv444p", "tv", "tv", "tv", "tetrahedral")),
    "default_8bit_hw": (_info("yuv420p", 8), ProcessingParams(video_codec="h264_videotoolbox"),
                        ("yuv420p", "yuv420p", "tv", "tv", "tv", "tetrahedral")),
    # BASELINE cfg 2: full-range source, forced bt709, trilinear, force_8bit
    "cfg2_pc_to_tv": (_info("yuvj420p", 8, "bt470bg", "pc"),
                      ProcessingParams(bit_depth_policy="force_8bit", lut_interp="trilinear", lut_input_matrix="bt709"),
                      ("yuv420p", "yuv420p", "pc", "tv", "tv", "trilinear")),
    # 4:2:2 full-range source: the range pass lands on yuv422p (ffmpeg.py:137-143); user asks for yuv444p out
    "pc_422_to_444": (_info("yuvj422p", 8, "bt709", "pc"), ProcessingParams(pix_fmt="yuv444p"),
                      ("yuv422p", "yuv444p", "pc", "tv", "tv", "tetrahedral")),
    # tags=inherit keeps a pc source pc through the stage and out of it
    "inherit_pc": (_info("yuvj420p", 8, "bt709", "pc"), ProcessingParams(lut_output_tags="inherit", pix_fmt="yuv420p"),
                   ("yuv420p", "yuv420p", "pc", "pc", "pc", "tetrahedral")),
    # 10-bit source through the default libx264: no 10-bit support -> yuv420p with swscale's ordered dither (ffmpeg.py:299-302)
    "hevc10_to_x264": (_info("yuv420p10le", 10), ProcessingParams(), ("yuv420p10le", "yuv420p", "tv", "tv", "tv", "tetrahedral")),
    # 10-bit source, libx265 preserve -> yuv420p10le (ffmpeg.py:297-298)
    "hevc10_to_x265": (_info("yuv420p10le", 10), ProcessingParams(video_codec="libx265"),
                       ("yuv420p10le", "yuv420p10le", "tv", "tv", "tv", "tetrahedral")),
    "c422_10_to_x265": (_info("yuv422p10le", 10), ProcessingParams(video_codec="libx265"),
                        ("yuv422p10le", "yuv420p10le", "tv", "tv", "tv", "tetrahedral")),
    # pro-mode master parameters (ffmpeg.py:417-433) on 10-bit 4:2:2 and 4:4:4 sources
    "master_422": (_info("yuv422p10le", 10), ProcessingParams(video_codec="prores_ks"),
                   ("yuv422p10le", "yuv422p10le", "tv", "tv", "tv", "tetrahedral")),
    "master_444": (_info("yuv444p10le", 10, "bt2020nc"), ProcessingParams(video_codec="prores_ks"),
                   ("yuv444p10le", "yuv422p10le", "tv", "tv", "tv", "tetrahedral")),
    # interp outside the GUI's two choices but inside the reference's whitelist (ffmpeg.py:243)
    "prism_8bit": (_info("yuv420p", 8), ProcessingParams(lut_interp="prism", pix_fmt="yuv422p"),
                   ("yuv420p", "yuv422p", "tv", "tv", "tv", "prism")),
    # untagged output: swscale's default (BT.601) matrix behind the LUT
    "tags_none": (_info("yuv420p", 8, "smpte170m"), ProcessingParams(lut_output_tags="none", pix_fmt="yuv420p"),
                  ("yuv420p", "yuv420p", "tv", "tv", "tv", "tetrahedral")),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_reference_options_to_pixels(name, tmp_path):
    from lut_renderer_b200.stage import ColourStage, DeviceContext
    import torch
    info, params, expect = CASES[name]
    table = synth.look_lut(17).astype(np.float32)
    cube = tmp_path / "look.cube"
    synth.write_cube(cube, table)
    plan = plan_colour_stage(params, Path(cube), info)
    assert (plan.in_pix_fmt, plan.out_pix_fmt, plan.in_range, plan.norm_range, plan.out_range, plan.interp) == expect
    st = ColourStage.from_plan(plan)
    lut = DeviceContext.get(0).lut_from_cube(plan.lut_path)
    frames = [synth.natural_frame(plan.in_pix_fmt, W, H, t=1, full_range=plan.in_range == "pc"), synth.noise_frame(plan.in_pix_fmt, W, H, 6)]
    src = torch.from_numpy(np.concatenate(frames)).cuda()
    dst = torch.zeros(len(frames) * st.out_frame_bytes, dtype=torch.uint8, device="cuda")
    st.run(lut, src, dst, len(frames), stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    out = dst.cpu().numpy()
    from oracle import oracle as O
    L = O.OracleLut.from_path(cube)
    for i, f in enumerate(frames):
        exp = oracle_stage(plan.in_pix_fmt, W, H, f, L, interp=plan.interp, in_matrix=plan.in_matrix, out_matrix=plan.out_matrix,
                           in_range=plan.in_range, norm_range=plan.norm_range, out_pix_fmt=plan.out_pix_fmt, out_range=plan.out_range)
        assert_bit_exact(plan.out_pix_fmt, W, H, out[i * st.out_frame_bytes:(i + 1) * st.out_frame_bytes], exp, f"{name} frame {i}")
    st.close()


def test_task_runner_over_real_pipes(tmp_path):
    """TaskRunner end to end over real OS pipes: a stand-in decoder process streams raw frames to stdout, the runner pushes them
    through the GPU stage in pipelined batches (reader thread | GPU | writer thread) and a stand-in encoder process collects its
    stdin.  (The image has no ffmpeg: the two child processes are python one-liners, everything between them is the product.)
    The bytes that reach the 'encoder' must be the oracle's, frame for frame, and the reference's status / progress protocol holds."""
    import subprocess
    import sys
    from lut_renderer_b200 import Task
    from lut_renderer_b200.task_manager import TaskCallbacks, TaskRunner
    from oracle import oracle as O
    w, h, nframes = 328, 38, 21                       # 21 frames in batches of 4: a ragged last batch
    info = VideoInfo(width=w, height=h, pix_fmt="yuv420p10le", bit_depth=10, colorspace="bt709", color_range="tv", fps=25.0,
                     duration=nframes / 25.0)
    cube = tmp_path / "look.cube"
    synth.write_cube(cube, synth.look_lut(17).astype(np.float32))
    frames = [synth.natural_frame("yuv420p10le", w, h, t=t) for t in range(nframes)]
    raw_in, raw_out = tmp_path / "in.raw", tmp_path / "out.raw"
    raw_in.write_bytes(np.concatenate(frames).tobytes())
    task = Task("t1", tmp_path / "in.mov", tmp_path / "out.mp4", cube, None, ProcessingParams(), source_info=info)   # libx264 -> yuv420p
    cmds = []

    def popen(cmd, **kw):
        cmds.append(list(cmd))
        if "pipe:1" in cmd:      # decoder
            fake = [sys.executable, "-c", f"import sys, shutil; shutil.copyfileobj(open({str(raw_in)!r}, 'rb'), sys.stdout.buffer)"]
        else:                    # encoder
            fake = [sys.executable, "-c", f"import sys, shutil; shutil.copyfileobj(sys.stdin.buffer, open({str(raw_out)!r}, 'wb'))"]
        return subprocess.Popen(fake, **kw)

    class Rec(TaskCallbacks):
        events = []

        def progress(self, tid, pct):
            self.events.append(("progress", pct))

        def status(self, tid, text):
            self.events.append(("status", text))

        def finished(self, tid, st):
            self.events.append(("finished", st))

    rec = Rec()
    TaskRunner(task, callbacks=rec, frames_per_batch=4, popen=popen).run()
    assert rec.events[-1] == ("finished", "completed"), rec.events[-5:]
    assert ("progress", 100) in rec.events
    dec_cmd = next(c for c in cmds if "pipe:1" in c)
    enc_cmd = next(c for c in cmds if "pipe:0" in c)
    assert dec_cmd[dec_cmd.index("-pix_fmt") + 1] == "yuv420p10le" and enc_cmd[enc_cmd.index("-pix_fmt") + 1] == "yuv420p"
    out = np.frombuffer(raw_out.read_bytes(), np.uint8)
    fb = w * h * 3 // 2
    assert out.size == nframes * fb
    L = O.OracleLut.from_path(cube)
    for i in (0, 3, 4, 19, 20):
        exp = oracle_stage("yuv420p10le", w, h, frames[i], L, out_pix_fmt="yuv420p")
        assert_bit_exact("yuv420p", w, h, out[i * fb:(i + 1) * fb], exp, f"piped frame {i}")


def test_task_runner_shards_one_clip_over_devices(tmp_path):
    """One clip cut into GOP-aligned frame ranges, one decoder -> GPU -> encoder chain per range running concurrently (here both
    ranges on cuda:0: the test box has one GPU), parts stitched afterwards.  Stand-in processes honour the same argv a real ffmpeg
    would get: the decoder starts at -ss and stops after -frames:v, the encoder writes its part file, the concat step joins the
    files of the list.  The stitched output must be the oracle's, frame for frame, in order."""
    import subprocess
    import math
    import sys
    from lut_renderer_b200 import Task
    from lut_renderer_b200.task_manager import TaskCallbacks, TaskRunner
    from oracle import oracle as O
    w, h, fps, nframes = 264, 40, 10.0, 47               # GOP = round(fps) = 10 -> ranges [0, 30) and [30, 47)
    info = VideoInfo(width=w, height=h, pix_fmt="yuv422p10le", bit_depth=10, colorspace="bt709", color_range="tv", fps=fps,
                     duration=nframes / fps)
    cube = tmp_path / "look.cube"
    synth.write_cube(cube, synth.look_lut(17).astype(np.float32))
    frames = [synth.natural_frame("yuv422p10le", w, h, t=t) for t in range(nframes)]
    fb = frames[0].nbytes
    raw_in = tmp_path / "in.raw"
    raw_in.write_bytes(np.concatenate(frames).tobytes())
    out_path = tmp_path / "out.mov"
    task = Task("t1", tmp_path / "in.mov", out_path, cube, None, ProcessingParams(video_codec="prores_ks"), source_info=info)
    cmds = []

    def popen(cmd, **kw):
        cmds.append(list(cmd))
        if "pipe:1" in cmd:      # decoder: frames [ss * fps, + frames:v)
            # accurate seek: the first frame whose pts is not before the target (the runner seeks half a frame early)
            start = math.ceil(float(cmd[cmd.index("-ss") + 1]) * fps - 1e-6) if "-ss" in cmd else 0
            count = int(cmd[cmd.index("-frames:v") + 1]) if "-frames:v" in cmd else nframes - start
            code = (f"import sys; f=open({str(raw_in)!r},'rb'); f.seek({start * fb}); sys.stdout.buffer.write(f.read({count * fb}))")
        elif "pipe:0" in cmd:    # encoder: its stdin into the part file named last
            code = f"import sys, shutil; shutil.copyfileobj(sys.stdin.buffer, open({cmd[-1]!r}, 'wb'))"
        else:                    # concat: join the files of the list into the output
            lst = cmd[cmd.index("-i") + 1]
            code = ("parts = [l.strip()[6:-1] for l in open(%r) if l.strip()]\n"
                    "open(%r, 'wb').write(b''.join(open(p, 'rb').read() for p in parts))") % (lst, cmd[-1])
        return subprocess.Popen([sys.executable, "-c", code], **kw)

    class Rec(TaskCallbacks):
        events = []

        def progress(self, tid, pct):
            self.events.append(("progress", pct))

        def finished(self, tid, st):
            self.events.append(("finished", st))

    rec = Rec()
    TaskRunner(task, callbacks=rec, frames_per_batch=4, popen=popen, devices=[0, 0]).run()
    assert rec.events[-1] == ("finished", "completed"), rec.events[-5:]
    decs = sorted((c for c in cmds if "pipe:1" in c), key=lambda c: "-ss" in c)
    assert "-ss" not in decs[0] and decs[0][decs[0].index("-frames:v") + 1] == "30"
    assert decs[1][decs[1].index("-ss") + 1] == "2.950000" and "-frames:v" not in decs[1]      # the tail runs to the end of the stream
    assert sum(1 for c in cmds if "concat" in c) == 1
    assert not list(tmp_path.glob("out.part*")) and not (tmp_path / "out.parts.txt").exists()       # parts cleaned up
    out = np.frombuffer(out_path.read_bytes(), np.uint8)
    assert out.size == nframes * fb
    L = O.OracleLut.from_path(cube)
    for i in (0, 1, 29, 30, 31, 46):
        exp = oracle_stage("yuv422p10le", w, h, frames[i], L)
        assert_bit_exact("yuv422p10le", w, h, out[i * fb:(i + 1) * fb], exp, f"sharded frame {i}")


@pytest.mark.timeout(120)
def test_task_runner_cancel_and_encoder_failure(tmp_path):
    """The pipelined pump must not hang when things go wrong: cancel() while an endless decoder is streaming ends the task as
    'canceled' (reference task_manager.py:38-44, 180-183); an encoder that dies mid-stream ends it as 'failed: ...'."""
    import subprocess
    import sys
    import threading
    import time
    from lut_renderer_b200 import Task
    from lut_renderer_b200.task_manager import TaskCallbacks, TaskRunner
    w, h = 264, 40
    info = VideoInfo(width=w, height=h, pix_fmt="yuv420p", bit_depth=8, colorspace="bt709", color_range="tv", fps=25.0, duration=1000.0)
    cube = tmp_path / "look.cube"
    synth.write_cube(cube, synth.look_lut(17).astype(np.float32))
    fb = w * h * 3 // 2
    endless = f"import sys\nb = bytes({fb})\nwhile True: sys.stdout.buffer.write(b)"
    sink = "import sys\nwhile sys.stdin.buffer.read(1 << 20): pass"
    dies = "import sys\nsys.stdin.buffer.read(1 << 20)\nsys.exit(3)"

    class Rec(TaskCallbacks):
        def __init__(self):
            self.events = []

        def status(self, tid, text):
            self.events.append(("status", text))

        def finished(self, tid, st):
            self.events.append(("finished", st))

    def run(encoder_code, cancel_after=None):
        rec = Rec()
        task = Task("t", tmp_path / "in.mov", tmp_path / "out.mp4", cube, None, ProcessingParams(), source_info=info)
        popen = lambda cmd, **kw: subprocess.Popen([sys.executable, "-c", endless if "pipe:1" in cmd else encoder_code], **kw)  # noqa: E731
        r = TaskRunner(task, callbacks=rec, frames_per_batch=4, popen=popen)
        if cancel_after is not None:
            threading.Timer(cancel_after, r.cancel).start()
        t0 = time.perf_counter()
        r.run()
        return rec.events, time.perf_counter() - t0

    events, dt = run(sink, cancel_after=1.5)
    assert events[-1] == ("finished", "canceled") and dt < 60, (events[-3:], dt)
    events, dt = run(dies)
    assert events[-1] == ("finished", "failed") and any(e[0] == "status" and e[1].startswith("failed") for e in events) and dt < 60, events[-3:]


def test_pro_pipeline_with_gpu_stage2(tmp_path):
    """Pro mode end to end through the runner with gpu_stage2: stage 1 = LUT + yuv422p10le master (reference ffmpeg.py:417-433,
    436-458), stage 2 = the master converted to the delivery pix_fmt WITHOUT a LUT (ffmpeg.py:459-471) -- both through the GPU
    stage between stand-in decoder / encoder processes.  The delivered bytes are the oracle's for the two steps chained, the master
    is deleted afterwards (cleanup_on_success), progress is split 50 / 50."""
    import subprocess
    import sys
    from lut_renderer_b200 import Task
    from lut_renderer_b200.task_manager import TaskCallbacks, TaskRunner
    from oracle import oracle as O
    w, h, nframes = 264, 40, 9
    info = VideoInfo(width=w, height=h, pix_fmt="yuv422p10le", bit_depth=10, colorspace="bt709", color_range="tv", fps=25.0, duration=nframes / 25.0)
    cube = tmp_path / "look.cube"
    synth.write_cube(cube, synth.look_lut(17).astype(np.float32))
    frames = [synth.natural_frame("yuv422p10le", w, h, t=t) for t in range(nframes)]
    src, master, out = tmp_path / "in.mov", tmp_path / "master.mov", tmp_path / "out.mp4"
    src.write_bytes(np.concatenate(frames).tobytes())          # the stand-in "container" is the raw frames themselves
    params = ProcessingParams(processing_mode="pro", video_codec="libx264")        # 10-bit master, libx264 delivery: yuv420p
    task = Task("t1", src, out, cube, None, params, source_info=info, intermediate_path=master)
    cmds = []

    def popen(cmd, **kw):
        cmds.append(list(cmd))
        if "pipe:1" in cmd:      # decoder of whichever stage: stream the file named after -i
            path = cmd[cmd.index("-i") + 1]
            fake = [sys.executable, "-c", f"import sys, shutil; shutil.copyfileobj(open({path!r}, 'rb'), sys.stdout.buffer)"]
        else:                    # encoder: stdin into the file named last
            fake = [sys.executable, "-c", f"import sys, shutil; shutil.copyfileobj(sys.stdin.buffer, open({cmd[-1]!r}, 'wb'))"]
        return subprocess.Popen(fake, **kw)

    class Rec(TaskCallbacks):
        events = []

        def progress(self, tid, pct):
            self.events.append(("progress", pct))

        def finished(self, tid, st):
            self.events.append(("finished", st))

        def log(self, tid, msg):
            self.events.append(("log", msg))

    rec = Rec()
    probe = lambda path: VideoInfo(width=w, height=h, pix_fmt="yuv422p10le", bit_depth=10, colorspace="bt709", color_range="tv", fps=25.0,  # noqa: E731
                                   duration=nframes / 25.0)
    TaskRunner(task, callbacks=rec, frames_per_batch=2, popen=popen, probe=probe, gpu_stage2=True).run()
    assert rec.events[-1] == ("finished", "completed"), [e for e in rec.events if e[0] != "progress"][-6:]
    assert len([c for c in cmds if "pipe:1" in c]) == 2 and len([c for c in cmds if "pipe:0" in c]) == 2       # two GPU stages, no plain ffmpeg
    enc2 = [c for c in cmds if "pipe:0" in c][1]
    assert enc2[enc2.index("-pix_fmt") + 1] == "yuv420p" and not any("lut3d" in t for t in enc2)
    assert not master.exists()
    pcts = [e[1] for e in rec.events if e[0] == "progress"]
    assert pcts == sorted(pcts) and pcts[-1] == 100 and any(40 <= p <= 50 for p in pcts)
    got = np.frombuffer(out.read_bytes(), np.uint8)
    fb = w * h * 3 // 2
    assert got.size == nframes * fb
    L = O.OracleLut.from_path(cube)
    for i in (0, 4, 8):
        m = oracle_stage("yuv422p10le", w, h, frames[i], L)                                   # stage 1: the master frame
        Y, U, V = [np.ascontiguousarray(p) for p in synth.split_planes(m, "yuv422p10le", w, h)]
        exp = synth.pack_planes(O.convert_yuv422p10("yuv420p", Y, U, V))                      # stage 2: LUT-less conversion
        assert_bit_exact("yuv420p", w, h, got[i * fb:(i + 1) * fb], exp, f"pro pipeline frame {i}")


@pytest.mark.parametrize("dialect", ["3dl", "csp", "dat"])
def test_other_lut_dialects_through_the_stage(tmp_path, dialect):
    """A task whose lut_path is one of the other files lut3d reads (reference ffmpeg.py:246 passes the path through; lut3d picks the
    reader by extension): DeviceContext.lut_from_file -> lutr_lut_load_file -> the same kernels.  Expected frames: the oracle chain
    on the table / DOMAIN scale the independent reader (oracle/lut_dialects.py) gets from the same file."""
    import torch
    from lut_renderer_b200.stage import ColourStage, DeviceContext
    from oracle import lut_dialects as D
    from oracle import oracle as O
    from test_lut_dialects import write_3dl, write_csp, write_dat
    w, h = 256, 32
    path = tmp_path / f"look.{dialect}"
    if dialect == "3dl":
        write_3dl(path, np.round(synth.look_lut(17) * 4095).astype(np.int64))
    elif dialect == "csp":
        write_csp(path, synth.look_lut(9).astype(np.float32), in_rng=((0, 2), (0, 1), (0, 4)))   # per-channel input range -> DOMAIN scale
    else:
        write_dat(path, synth.look_lut(33).astype(np.float32))
    size, scale, table = D.READERS[dialect](path.read_bytes())
    frames = [synth.natural_frame("yuv422p10le", w, h, t=3), synth.noise_frame("yuv422p10le", w, h, seed=5)]
    st = ColourStage(w, h, "yuv422p10le")
    lut = DeviceContext.get(0).lut_from_file(path)
    assert lut.size == size and DeviceContext.get(0).lut_from_file(path) is lut
    src = torch.from_numpy(np.concatenate(frames)).cuda()
    dst = torch.zeros(len(frames) * st.out_frame_bytes, dtype=torch.uint8, device="cuda")
    st.run(lut, src, dst, len(frames), stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    assert st.kernel_name == "k422_warp"
    out = dst.cpu().numpy()
    L = O.OracleLut.from_table(table, scale)
    for i, f in enumerate(frames):
        assert_bit_exact("yuv422p10le", w, h, out[i * st.out_frame_bytes:(i + 1) * st.out_frame_bytes],
                         oracle_stage("yuv422p10le", w, h, f, L), f".{dialect} frame {i}")
    st.close()


@pytest.mark.parametrize("pipe_mb,batch", [("1", 2), ("64", 2), ("64", 1), ("1", 0)])
def test_pump_output_pages_are_not_reused_while_the_pipe_holds_them(tmp_path, monkeypatch, pipe_mb, batch):
    """lutr_stage_pump lends its pinned output pages to the encoder's pipe (vmsplice) instead of copying them: a batch must not be
    refilled before the (slow) reader has taken every page of it.  Distinct frames, a consumer that reads in small sips with pauses,
    pipes smaller and larger than a batch; the bytes that arrive are the bytes run_host produces for the same frames.
    Reference: the data between the two ffmpeg processes of TaskRunner._run_stage (task_manager.py:134-190)."""
    import subprocess
    import sys
    from lut_renderer_b200.stage import ColourStage
    monkeypatch.setenv("LUTR_PIPE_MB", pipe_mb)
    w, h, nframes = 1920, 1080, 22
    frames = np.concatenate([synth.natural_frame("yuv420p", w, h, t=t) for t in range(nframes)])
    (tmp_path / "in.raw").write_bytes(frames.tobytes())
    st = ColourStage(w, h, "yuv420p")
    lut = st.ctx.lut_from_table(synth.look_lut(33).astype(np.float32))
    expected = np.empty(nframes * st.out_frame_bytes, np.uint8)
    st.run_host(lut, frames, expected, nframes)
    gen = subprocess.Popen([sys.executable, "-c", f"import sys, shutil; shutil.copyfileobj(open({str(tmp_path / 'in.raw')!r}, 'rb'), sys.stdout.buffer)"],
                           stdout=subprocess.PIPE, bufsize=0)
    sip = ("import time\nf = open(0, 'rb', buffering=0); o = open(%r, 'wb'); n = 0\n"
           "while True:\n    b = f.read(1 << 19)\n    if not b: break\n    o.write(b); n += 1\n    if n %% 8 == 0: time.sleep(0.004)\n" % str(tmp_path / "out.raw"))
    sink = subprocess.Popen([sys.executable, "-c", sip], stdin=subprocess.PIPE, bufsize=0)
    stats = st.pump(lut, gen.stdout.fileno(), sink.stdin.fileno(), batch)
    sink.stdin.close()
    assert sink.wait() == 0 and gen.wait() == 0
    assert stats["frames_out"] == nframes
    got = np.frombuffer((tmp_path / "out.raw").read_bytes(), np.uint8)
    assert got.size == expected.size
    bad = np.flatnonzero(got != expected)
    assert bad.size == 0, f"{bad.size} bytes differ, first at frame {bad[0] // st.out_frame_bytes}"
    st.close(), lut.close()
