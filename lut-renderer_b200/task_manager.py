"""Runner of one task: decode -> raw pipe -> GPU colour stage -> raw pipe -> encode.

Replaces ``TaskRunner._run_stage`` of the reference (reference src/lut_renderer/task_manager.py:29-216), which blocks on a
single ffmpeg process whose filter graph contains ``lut3d``.  Signals become plain callbacks (no Qt here); status
strings, progress split over stages, cancellation and clean-up keep the reference's behaviour:

* ``status``: "running" / "completed" / "canceled" / "failed: exit N" / "failed: <exception text>"
* progress: 0..100, pro mode splits 50/50 over its two stages (task_manager.py:86-91)
* cancel(): cooperative flag + terminate() of the child processes (task_manager.py:38-44,180-183)
* intermediates flagged ``cleanup_on_success`` are deleted after a successful run (task_manager.py:117-122)

Stages without a LUT (pro mode stage 2) still run as the reference's single ffmpeg command.
"""
from __future__ import annotations

import ctypes
import shlex
import subprocess
import threading
import time
from typing import Callable, List, Optional

from .ffmpeg import CommandStage, RawPipeUnsupported, build_command, build_gpu_stage, build_pipeline, build_sharded_gpu_stage
from .media_info import VideoInfo, probe_video
from .models import Task, TaskStatus


class TaskCallbacks:
    """Default no-op sinks; the GUI connects its own (these stand in for the reference's Qt signals)."""

    def progress(self, task_id: str, percent: int) -> None: ...
    def status(self, task_id: str, text: str) -> None: ...
    def finished(self, task_id: str, status: str) -> None: ...
    def log(self, task_id: str, message: str) -> None: ...


class TaskRunner:
    def __init__(self, task: Task, ffmpeg_bin: str = "ffmpeg", device: int = 0, callbacks: Optional[TaskCallbacks] = None,
                 frames_per_batch: int = 0, popen: Callable = subprocess.Popen, probe: Callable = probe_video,
                 devices: Optional[List[int]] = None, gpu_stage2: bool = False):
        self.task, self.ffmpeg_bin, self.device = task, ffmpeg_bin, device
        # more than one entry: every LUT stage of the task is cut into GOP-aligned frame ranges, one per listed GPU
        self.devices = list(devices) if devices else [device]
        # pro mode stage 2 has no LUT and by default stays the reference's single ffmpeg command; with gpu_stage2 its pixel-format
        # conversion (yuv422p10le master -> yuv420p / yuv420p10le / yuv422p) runs on the GPU stage between a decoder and an encoder
        self.gpu_stage2 = gpu_stage2
        self.signals = callbacks or TaskCallbacks()
        self.frames_per_batch = max(0, frames_per_batch)   # 0 = the library picks ~64 MB batches
        self._popen, self._probe = popen, probe
        self._procs: List = []
        self._cancel_flags: List = []   # ctypes flags of running lutr_stage_pump calls
        self._cancelled = False
        self._lock = threading.Lock()

    # ------------------------------------------------------------------------------------------------------------
    def cancel(self) -> None:
        self._cancelled = True
        with self._lock:
            procs = list(self._procs)
            for flag in self._cancel_flags:
                flag.value = 1
        for p in procs:
            try:
                if p.poll() is None:
                    p.terminate()
            except Exception:
                pass

    def _log(self, message: str) -> None:
        self.signals.log(self.task.task_id, message)

    def run(self) -> None:
        tid = self.task.task_id
        self.signals.status(tid, TaskStatus.RUNNING.value)
        self._log("开始")
        self.task.started_at = time.time()
        try:
            stages = build_pipeline(self.task, ffmpeg_bin=self.ffmpeg_bin)
            if not stages:
                raise RuntimeError("No ffmpeg stages built.")
            for index, stage in enumerate(stages):
                if self._cancelled:
                    break
                self._log(f"阶段 {index + 1}/{len(stages)}: {stage.name}")
                info = self.task.source_info
                if stage.probe_source:
                    try:
                        info = self._probe(stage.source_path)
                    except Exception as exc:
                        info = None
                        self._log(f"提示: 阶段输入探测失败（将按未知处理）: {exc}")
                span = 100 if len(stages) == 1 else 50
                base = 0 if index == 0 else 100 - span
                final = index == len(stages) - 1
                ok = self._run_stage(stage, info, base, span, final)
                if ok is None:
                    self.signals.status(tid, TaskStatus.CANCELED.value)
                    self._log("已取消")
                    self.signals.finished(tid, TaskStatus.CANCELED.value)
                    return
                if not ok:
                    code = self._last_returncode
                    self.signals.status(tid, f"{TaskStatus.FAILED.value}: exit {code}")
                    self._log(f"失败：退出码 {code}")
                    self.signals.finished(tid, TaskStatus.FAILED.value)
                    return
            for stage in stages:
                if stage.cleanup_on_success and stage.output_path.exists():
                    try:
                        stage.output_path.unlink()
                    except Exception:
                        pass
            self.signals.progress(tid, 100)
            self.signals.status(tid, TaskStatus.COMPLETED.value)
            self._log("完成")
            self.signals.finished(tid, TaskStatus.COMPLETED.value)
        except Exception as exc:
            self.signals.status(tid, f"{TaskStatus.FAILED.value}: {exc}")
            self._log(f"失败：{exc}")
            self.signals.finished(tid, TaskStatus.FAILED.value)

    # ------------------------------------------------------------------------------------------------------------
    _last_returncode = -1

    def _spawn(self, cmd, **kw):
        p = self._popen(cmd, **kw)
        with self._lock:
            self._procs.append(p)
        return p

    def _run_stage(self, stage: CommandStage, info: Optional[VideoInfo], base: int, span: int, final: bool) -> Optional[bool]:
        if not stage.lut_path:
            gpu = None
            if self.gpu_stage2:
                try:
                    gpu = build_gpu_stage(stage, info, self.ffmpeg_bin, convert_without_lut=True)
                except RawPipeUnsupported:
                    gpu = None
            if gpu is None:
                return self._run_plain(stage, info, base, span, final)
        total = int(round(info.duration * info.fps)) if info and info.duration and info.fps else 0
        done_lock = threading.Lock()
        state = {"done": 0, "last": -1}

        def on_frames(n: int) -> None:   # progress over all shards of the stage
            with done_lock:
                state["done"] += n
                if not total:
                    return
                pct = min(int(state["done"] / total * span), span - (0 if final else 1))
                prog = min(base + pct, 100 if final else 99)
                if prog == state["last"]:
                    return
                state["last"] = prog
            self.signals.progress(self.task.task_id, prog)

        try:
            if not stage.lut_path:
                sharded = None     # (gpu: the LUT-less conversion stage built above)
            else:
                sharded = build_sharded_gpu_stage(stage, info, self.devices, self.ffmpeg_bin) if len(self.devices) > 1 else None
                gpu = build_gpu_stage(stage, info, self.ffmpeg_bin) if sharded is None else None
        except RawPipeUnsupported as exc:
            # rawvideo carries no timestamps: this clip keeps the reference's own single-process command (CPU lut3d inside
            # ffmpeg), said out loud in the task log -- never a silent switch
            self._log(f"提示: GPU 色彩阶段不适用于此片段（{exc}），本阶段改用参考实现的单进程 ffmpeg 命令")
            return self._run_plain(stage, info, base, span, final)
        for note in stage.notes:
            self._log(note)
        if sharded is None:
            self._log(f"解码: {shlex.join(gpu.decode_cmd)}")
            self._log(f"编码: {shlex.join(gpu.encode_cmd)}")
            rc, _ = self._pump(gpu.plan, gpu.decode_cmd, gpu.encode_cmd, self.device, on_frames)
        else:
            # one clip over several GPUs: contiguous GOP-aligned frame ranges, one decoder / GPU stage / encoder chain per range,
            # no exchange between them; the parts are stitched by stream copy afterwards
            results: List[Optional[int]] = [None] * len(sharded.shards)

            def work(k: int, sh) -> None:
                self._log(f"分片 {sh.index} (GPU {sh.device}, 帧 {sh.start_frame}..{sh.start_frame + sh.frames - 1}): {shlex.join(sh.decode_cmd)}")
                try:
                    rc_k, frames_k = self._pump(sharded.base.plan, sh.decode_cmd, sh.encode_cmd, sh.device, on_frames)
                    # a part that is short (or, for a closed range, long) would lose / duplicate frames at the cut: refuse to stitch
                    if rc_k == 0 and (frames_k < sh.frames or (frames_k > sh.frames and not sh.open_ended)):
                        self._log(f"分片 {sh.index} 帧数不符: 计划 {sh.frames}，实际 {frames_k}")
                        rc_k = -2
                    results[k] = rc_k
                except BaseException as exc:  # noqa: BLE001 -- reported through the task status
                    self._log(f"分片 {sh.index} 失败: {exc}")
                    results[k] = -1

            threads = [threading.Thread(target=work, args=(k, sh), daemon=True) for k, sh in enumerate(sharded.shards)]
            for th in threads:
                th.start()
            for th in threads:
                th.join()
            rc = next((r for r in results if r not in (0, None)), 0)
            if self._cancelled:
                rc = None
            elif rc == 0:
                sharded.concat_list_path.write_text(sharded.concat_list_text, encoding="utf-8")
                self._log(f"拼接: {shlex.join(sharded.concat_cmd)}")
                proc = self._spawn(sharded.concat_cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                rc = proc.wait()
            for path in [sh.part_path for sh in sharded.shards] + [sharded.concat_list_path]:
                try:
                    if rc == 0 and path.exists():
                        path.unlink()
                except Exception:
                    pass
        if rc is None or self._cancelled:
            return None
        self._last_returncode = rc
        if rc == 0:
            if not final:
                self.signals.progress(self.task.task_id, base + span)
            return True
        return False

    def _pump(self, plan, decode_cmd, encode_cmd, device: int, on_frames: Callable[[int], None]):
        """decoder process -> raw pipe -> GPU colour stage on ``device`` -> raw pipe -> encoder process.

        The data path is one library call (``ColourStage.pump`` -> ``lutr_stage_pump``): a reader thread fills a pinned ring straight
        from the decoder's pipe, the GPU stage runs H2D / kernel / D2H on its three streams, a writer thread drains into the encoder's
        pipe -- no Python on the per-frame path.  This method only owns the two child processes, cancellation and the exit codes.
        Returns (first non-zero exit code of the two processes -- 0 = both fine, None when cancelled --, frames processed)."""
        from .stage import ColourStage  # deferred: loads liblutr.so, fails loudly when it is missing
        colour = ColourStage.from_plan(plan, device=device)
        lut = colour.ctx.lut_from_file(plan.lut_path) if plan.lut_path else None   # None: LUT-less conversion (interp="none")
        dec = self._spawn(decode_cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, bufsize=0)
        enc = self._spawn(encode_cmd, stdin=subprocess.PIPE, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, bufsize=0)
        flag = ctypes.c_int32(1 if self._cancelled else 0)
        with self._lock:
            self._cancel_flags.append(flag)
        seen = [0]

        def progress(total: int) -> None:
            on_frames(total - seen[0])
            seen[0] = total

        error: Optional[BaseException] = None
        frames_done = 0
        try:
            stats = colour.pump(lut, dec.stdout.fileno(), enc.stdin.fileno(), self.frames_per_batch, flag, progress)
            frames_done = int(stats["frames_out"])
        except BaseException as exc:  # noqa: BLE001 -- a failing GPU call / pipe must take the children down, not leave them blocked
            error = exc
        finally:
            with self._lock:
                self._cancel_flags.remove(flag)
            try:
                enc.stdin.close()
            except Exception:
                pass
            if self._cancelled or error is not None:
                for p in (dec, enc):
                    if p.poll() is None:
                        p.kill()
            rc_dec, rc_enc = dec.wait(), enc.wait()
            try:
                dec.stdout.close()
            except Exception:
                pass
            colour.close()
        if self._cancelled:
            return None, frames_done
        if error is not None:
            raise RuntimeError(f"colour stage / raw pipe failed: {error}")
        return (rc_enc or rc_dec), frames_done

    def _run_plain(self, stage: CommandStage, info: Optional[VideoInfo], base: int, span: int, final: bool) -> Optional[bool]:
        """The reference's own single ffmpeg process, progress parsed from its log: stages without a LUT, and LUT stages whose
        timing cannot cross a raw pipe (``RawPipeUnsupported``)."""
        import re
        cmd = build_command(stage.source_path, stage.output_path, stage.params, stage.lut_path, self.ffmpeg_bin, info, stage.notes)
        for note in stage.notes:
            self._log(note)
        self._log(f"命令: {shlex.join(cmd)}")
        dur_re = re.compile(r"Duration: (\d+):(\d+):(\d+(?:\.\d+)?)")
        time_re = re.compile(r"time=(\d+):(\d+):(\d+(?:\.\d+)?)")
        secs = lambda m: int(m.group(1)) * 3600 + int(m.group(2)) * 60 + float(m.group(3))  # noqa: E731
        proc = self._spawn(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, bufsize=1)
        duration, last = None, -1
        for line in proc.stdout:
            if self._cancelled:
                break
            if line.strip():
                self._log(line.strip())
            if duration is None:
                m = dur_re.search(line)
                if m:
                    duration = secs(m)
                    continue
            m = time_re.search(line)
            if m and duration:
                pct = int(secs(m) / duration * span)
                if not final:
                    pct = min(pct, max(0, span - 1))
                prog = min(base + pct, 100 if final else 99)
                if prog != last:
                    last = prog
                    self.signals.progress(self.task.task_id, prog)
        if self._cancelled:
            if proc.poll() is None:
                proc.kill()
            return None
        self._last_returncode = proc.wait()
        if self._last_returncode == 0:
            if not final:
                self.signals.progress(self.task.task_id, base + span)
            return True
        return False


class TaskManager:
    """Queue + bounded concurrency (reference task_manager.py:219-315: QThreadPool with 1-16 runners); tasks are dealt to
    the visible GPUs round-robin so concurrent runners do not share a device unless there are more runners than GPUs."""

    def __init__(self, ffmpeg_bin: str = "ffmpeg", max_concurrency: int = 2, devices: Optional[List[int]] = None,
                 callbacks: Optional[TaskCallbacks] = None, shard_clips: bool = False):
        self.ffmpeg_bin, self.max_concurrency = ffmpeg_bin, max(1, min(16, max_concurrency))
        self.devices = devices or [0]
        # shard_clips: every task is cut into frame ranges over ALL devices (long clips, few tasks) instead of dealing whole
        # tasks to devices (many clips: BASELINE cfg 5)
        self.shard_clips = shard_clips
        self.callbacks = callbacks or TaskCallbacks()
        self.tasks: List[Task] = []
        self._runners = {}

    def add_tasks(self, tasks: List[Task]) -> None:
        self.tasks.extend(tasks)

    def start_all(self) -> None:
        pending = [t for t in self.tasks if t.status == TaskStatus.PENDING]
        sem = threading.Semaphore(self.max_concurrency)
        threads = []

        def work(task: Task, device: int):
            with sem:
                runner = TaskRunner(task, self.ffmpeg_bin, device, self.callbacks,
                                    devices=self.devices if self.shard_clips else None)
                self._runners[task.task_id] = runner
                runner.run()

        for i, t in enumerate(pending):
            th = threading.Thread(target=work, args=(t, self.devices[i % len(self.devices)]), daemon=True)
            th.start()
            threads.append(th)
        for th in threads:
            th.join()

    def cancel_task(self, task_id: str) -> None:
        r = self._runners.get(task_id)
        if r:
            r.cancel()
