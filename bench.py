#!/usr/bin/env python
"""Benchmark of the fused colour stage (BASELINE.json metric: 4K 10-bit frames/s and Gpix/s per GPU, % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg3] [--content natural|noise]

A step = one pass of the hot path over one batch of `frames_per_step` distinct synthetic frames that are already
resident in HBM (working set >> L2, so no L2 flush is needed between steps).  Prints ONE JSON line (rank 0).
For N > 1 launch under torchrun: every rank owns one GPU and a contiguous frame range (no collective on the path).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CONFIGS = {
    # BASELINE.json configs[0..3]; cfg3 is the one the metric is quoted on
    "cfg1": dict(pix_fmt="yuv420p", w=1920, h=1080, lut=33, interp="tetrahedral", in_range="tv", norm_range="tv",
                 frames_per_step=256, desc="1080p yuv420p 8-bit, 33-point look .cube, tetrahedral, bt709 tv"),
    "cfg2": dict(pix_fmt="yuv420p", w=1920, h=1080, lut=33, interp="trilinear", in_range="pc", norm_range="tv",
                 frames_per_step=256, desc="1080p yuv420p full-range source, pc->tv pass, bt709 forced, 33-point, trilinear"),
    "cfg3": dict(pix_fmt="yuv422p10le", w=3840, h=2160, lut=65, interp="tetrahedral", in_range="tv", norm_range="tv",
                 frames_per_step=64, desc="4K UHD yuv422p10le, 65-point look .cube, tetrahedral, bt709 tv->tv"),
    # not a BASELINE config: pro-mode stage 1 on an 8-bit camera file (1.5 + 4 = 5.5 B/px)
    "master8": dict(pix_fmt="yuv420p", out_pix_fmt="yuv422p10le", w=3840, h=2160, lut=65, interp="tetrahedral", in_range="tv",
                    norm_range="tv", frames_per_step=64,
                    desc="4K UHD yuv420p 8-bit source -> 65-point look .cube, tetrahedral -> yuv422p10le master"),
    "master10": dict(pix_fmt="yuv420p10le", out_pix_fmt="yuv422p10le", w=3840, h=2160, lut=65, interp="tetrahedral", in_range="tv",
                     norm_range="tv", frames_per_step=64,
                     desc="4K UHD yuv420p10le 10-bit 4:2:0 source -> 65-point look .cube, tetrahedral -> yuv422p10le master"),
    "hevc10": dict(pix_fmt="yuv420p10le", out_pix_fmt="yuv420p10le", w=3840, h=2160, lut=65, interp="tetrahedral", in_range="tv",
                   norm_range="tv", frames_per_step=64,
                   desc="4K UHD yuv420p10le source -> 65-point look .cube, tetrahedral -> yuv420p10le (libx265 preserve)"),
    "hevc10_422": dict(pix_fmt="yuv422p10le", out_pix_fmt="yuv420p10le", w=3840, h=2160, lut=65, interp="tetrahedral", in_range="tv",
                       norm_range="tv", frames_per_step=64,
                       desc="4K UHD yuv422p10le source -> 65-point look .cube, tetrahedral -> yuv420p10le (libx265 preserve)"),
    "deliver8": dict(pix_fmt="yuv420p10le", out_pix_fmt="yuv420p", w=3840, h=2160, lut=65, interp="tetrahedral", in_range="tv",
                     norm_range="tv", frames_per_step=64,
                     desc="4K UHD yuv420p10le source -> 65-point look .cube, tetrahedral -> yuv420p 8-bit delivery (ordered dither)"),
    "cfg4": dict(pix_fmt="yuv422p10le", w=7680, h=4320, lut=65, interp="tetrahedral", in_range="tv", norm_range="tv",
                 frames_per_step=16, desc="8K yuv422p10le, 65-point look .cube, tetrahedral, frame-range sharded"),
}
# BASELINE.json configs[4]: batch queue of 64 mixed clips with per-clip LUTs, sharded by clip (handled by run_cfg5)
CFG5 = dict(clips=64, frames_per_clip=24,
            desc="batch queue of 64 mixed 1080p yuv420p (33-point) / 4K yuv422p10le (65-point) clips, per-clip .cube LUTs, sharded by clip")
BYTES_PER_PIXEL = {"yuv420p": 1.5, "yuv422p": 2.0, "yuv444p": 3.0, "yuv420p10le": 3.0, "yuv422p10le": 4.0, "yuv444p10le": 6.0}  # per side; algorithmic bytes/frame = (in + out) * w * h


def config_dict(cfg, args, world):
    """The `config` object of the JSON line: what is measured, identical for the repo arm and the reference arm."""
    F = cfg["frames_per_step"]
    bpp_in, bpp_out = BYTES_PER_PIXEL[cfg["pix_fmt"]], BYTES_PER_PIXEL[cfg.get("out_pix_fmt") or cfg["pix_fmt"]]
    return {"workload": cfg["desc"], "content": args.content, "frames_per_launch": F, "float_mode": args.float_mode, "options": list(args.opt),
            "l2": f"inputs larger than L2 ({F * bpp_in * cfg['w'] * cfg['h'] / 1e6:.0f} MB read + {F * bpp_out * cfg['w'] * cfg['h'] / 1e6:.0f} MB written per launch)",
            "sharding": "contiguous frame range per GPU, no collective"}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def ncu_traffic(config_name, frames_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, from the committed `ncu --set full` capture of
    this config (profiles/traffic.json holds bytes per frame of that capture), scaled to this run's frames per launch."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            return int(json.load(f)[config_name]["dram_bytes_per_frame"] * frames_per_launch)
    except Exception:
        return None


def pcie_simultaneous_gbs(device, nbytes=256 << 20, reps=6):
    """Pinned host <-> device copy rate per direction with both directions busy (tools/pcie_peak.py, shortened)."""
    import torch
    h_in, h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory(), torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_in, d_out = torch.empty(nbytes, dtype=torch.uint8, device=device), torch.empty(nbytes, dtype=torch.uint8, device=device)
    s1, s2 = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)

    def run(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        return nbytes * n / (time.perf_counter() - t0) / 1e9

    run(2)
    return run(reps)


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _loop(self):
        nv = self._nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self._nv:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join()
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def make_lut_table(n):
    from lut_renderer_b200 import synth
    return synth.look_lut(n).astype(np.float32)


def make_host_frames(cfg, content, count):
    from lut_renderer_b200 import synth
    if content == "noise":
        return [synth.noise_frame(cfg["pix_fmt"], cfg["w"], cfg["h"], seed=1 + i) for i in range(count)]
    if content == "flat":  # one colour per frame: every lane of a warp gathers the same LUT nodes (lower bound of the gather cost)
        shapes, dt = synth.plane_shapes(cfg["pix_fmt"], cfg["w"], cfg["h"])
        mid = 128 << (np.dtype(dt).itemsize - 1) * 2
        return [synth.pack_planes([np.full(sh, mid + 37 * i + 11 * p, dt) for p, sh in enumerate(shapes)]) for i in range(count)]
    if content.startswith("grain"):  # "grain0.02": the natural scene with that much per-channel noise (fraction of full scale)
        g = float(content[5:])
        return [synth.natural_frame(cfg["pix_fmt"], cfg["w"], cfg["h"], t=7 * i, seed=0, full_range=cfg["in_range"] == "pc", grain=g,
                                    chroma_grain=g) for i in range(count)]
    return [synth.natural_frame(cfg["pix_fmt"], cfg["w"], cfg["h"], t=7 * i, seed=0, full_range=cfg["in_range"] == "pc")
            for i in range(count)]


def device_frames(cfg, content, nframes, device, rank):
    """`nframes` distinct packed frames in one device buffer: a few CPU-generated base frames, the rest derived from
    them on the GPU by rolling the planes (keeps the content class, makes every frame's bytes distinct)."""
    import torch
    from lut_renderer_b200 import synth
    base = make_host_frames(cfg, content, 4)
    shapes, dt = synth.plane_shapes(cfg["pix_fmt"], cfg["w"], cfg["h"])
    item = np.dtype(dt).itemsize
    fb = base[0].nbytes
    buf = torch.empty(nframes * fb, dtype=torch.uint8, device=device)
    dev_base = [torch.from_numpy(b).to(device) for b in base]
    for i in range(nframes):
        src = dev_base[i % 4]
        k = i // 4 + 13 * rank
        off = 0
        for p, (r, c) in enumerate(shapes):
            n = r * c * item
            plane = src[off:off + n].view(r, c * item)
            # roll by an even number of samples / rows so chroma siting stays consistent
            sx = (8 * k if p == 0 else 4 * k) * item
            sy = 2 * k if (p == 0 or cfg["pix_fmt"] != "yuv420p") else k
            buf[i * fb + off:i * fb + off + n].view(r, c * item).copy_(torch.roll(plane, shifts=(sy % r, sx % (c * item)), dims=(0, 1)))
            off += n
    return buf, base


def run_reference(args, cfg, rank, world):
    """--impl reference: the reference's CPU path (see oracle/cpu_baseline.py) on the host cores, same metric/config."""
    if rank != 0:
        return
    from oracle import cpu_baseline as C
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    lut = O.OracleLut.from_table(make_lut_table(cfg["lut"]))
    frames = make_host_frames(cfg, args.content, 4)
    # a step = a bounded sample: every host thread converts frames for `sample_s` seconds
    sample_s = max(1.0, min(8.0, 60.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        C.time_cpu_path(cfg["pix_fmt"], cfg["w"], cfg["h"], lut, frames, seconds=min(sample_s, 1.0), interp=cfg["interp"],
                        out_pix_fmt=cfg.get("out_pix_fmt"))
    total_frames, total_s, legs = 0, 0.0, ""
    for _ in range(args.steps):
        r = C.time_cpu_path(cfg["pix_fmt"], cfg["w"], cfg["h"], lut, frames, seconds=sample_s, interp=cfg["interp"],
                            out_pix_fmt=cfg.get("out_pix_fmt"))
        total_frames += r["frames"]
        total_s += r["seconds"]
        legs = r["legs"]
    fps = total_frames / total_s
    mpix = cfg["w"] * cfg["h"] / 1e6
    line = {
        "impl": "reference", "metric": "4K 10-bit frames/s" if args.config == "cfg3" else f"{args.config} frames/s",
        "value": round(fps, 3), "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1e3 * total_s / max(1, args.steps), 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u16+f32" if "10" in cfg["pix_fmt"] else "u8+f32", "data": "synthetic",
        "gpix_per_s": round(fps * mpix / 1e3, 4),
        "config": config_dict(cfg, args, world),
        "cpu_baseline": {"value": round(fps, 3), "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{legs}; {cores} frame-parallel threads x {sample_s:.1f}s per step, {total_frames} frames total "
                                   f"({total_frames / max(1, args.steps):.1f} per step)"},
        "e2e": {"value": round(fps, 3), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_cfg5(args):
    """BASELINE cfg 5: 64 clips (1080p 8-bit / 4K 10-bit alternating), every clip with its own .cube LUT, whole clips dealt to
    the ranks LPT-first (lut_renderer_b200.sharding.assign_clips), no collective on the data path.  A step = the whole queue.
    `value` = frames/s with the clips resident in HBM and the LUT cache warm (each distinct LUT uploaded once per device);
    `e2e` = the same queue from pinned host memory INCLUDING the per-clip .cube parse + upload (cold LUT cache)."""
    import tempfile
    import torch
    import torch.distributed as dist
    from lut_renderer_b200 import synth
    from lut_renderer_b200.sharding import Clip, assign_clips
    from lut_renderer_b200.stage import ColourStage, DeviceContext

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": "cfg5 is a scheduling config; the CPU chain is timed per format under cfg1 / cfg3"}))
        return
    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    from lut_renderer_b200.sharding import bind_to_device_cpus
    full_affinity = bind_to_device_cpus(local_rank)   # pinned host frames on the GPU's own NUMA node (one process per GPU)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(device))
    nclips = CFG5["clips"]
    fpc = args.frames_per_step or CFG5["frames_per_clip"]
    kinds = {0: dict(pix_fmt="yuv420p", w=1920, h=1080, lut=33), 1: dict(pix_fmt="yuv422p10le", w=3840, h=2160, lut=65)}
    clips = [Clip(f"clip{i:02d}", kinds[i % 2]["w"], kinds[i % 2]["h"], kinds[i % 2]["pix_fmt"], fpc, f"look{i:02d}.cube") for i in range(nclips)]
    mine = assign_clips(clips, world)[rank]
    ctx = DeviceContext.get(local_rank)
    # per-clip LUT files: the analytic look blended with a clip-specific random table (distinct bytes, values inside [0,1])
    tmp = tempfile.mkdtemp(prefix="lutr_cfg5_")
    paths = {}
    for i in mine:
        n = kinds[i % 2]["lut"]
        table = (0.9 * synth.look_lut(n) + 0.1 * synth.random_lut(n, 100 + i)).astype(np.float32)
        paths[i] = synth.write_cube(os.path.join(tmp, clips[i].lut), table)
    stages, src, dst, pins = {}, {}, {}, {}
    for k, cfg in kinds.items():
        c = dict(cfg, in_range="tv")
        st = ColourStage(cfg["w"], cfg["h"], cfg["pix_fmt"], device=local_rank)
        stages[k] = st
        src[k], base = device_frames(c, "natural", fpc, device, rank)
        dst[k] = torch.empty(fpc * st.out_frame_bytes, dtype=torch.uint8, device=device)
        pin_in, pin_out = ctx.pinned_buffer(fpc * st.in_frame_bytes), ctx.pinned_buffer(fpc * st.out_frame_bytes)
        pin_in.array[:] = src[k].cpu().numpy()
        pins[k] = (pin_in, pin_out)
    tstream = torch.cuda.Stream(device=device)
    torch.cuda.set_stream(tstream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    luts = {i: ctx.lut_from_cube(paths[i]) for i in mine}   # warm cache: one upload per distinct LUT and device

    def queue_resident():
        for i in mine:
            stages[i % 2].run(luts[i], src[i % 2], dst[i % 2], fpc, stream=tstream)

    for _ in range(max(args.warmup, 3)):
        queue_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    n0 = sum(st.launch_count for st in stages.values())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    e0.record()
    for _ in range(args.steps):
        queue_resident()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    launches = sum(st.launch_count for st in stages.values()) - n0
    # end to end: cold LUT cache (parse + upload per clip) and host frames
    e2e_steps = max(1, min(args.steps, 2))
    barrier()
    # cold LUT cache: every clip's .cube is read, parsed and uploaded again; the NEXT clip's LUT is loaded on a background thread
    # (DeviceContext.prefetch_lut) while the GPU works on the current clip, as a batch queue would
    parts = {"wait_for_lut": 0.0, "run_host": 0.0, "free": 0.0}
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        nxt = ctx.prefetch_lut(paths[mine[0]], cached=False) if mine else None
        for n_, i in enumerate(mine):
            ta = time.perf_counter()
            lut = nxt.result()
            nxt = ctx.prefetch_lut(paths[mine[n_ + 1]], cached=False) if n_ + 1 < len(mine) else None
            tb = time.perf_counter()
            stages[i % 2].run_host(lut, pins[i % 2][0], pins[i % 2][1], fpc)
            tc = time.perf_counter()
            lut.close()
            td = time.perf_counter()
            parts["wait_for_lut"] += tb - ta; parts["run_host"] += tc - tb; parts["free"] += td - tc
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([ms, dt], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, dt = float(t[0].item()), float(t[1].item())
        lt = torch.tensor([launches], dtype=torch.int64, device=device)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    if rank == 0:
        frames = nclips * fpc
        pix = sum(c.width * c.height * c.frames for c in clips)
        nbytes = sum(2 * BYTES_PER_PIXEL[c.pix_fmt] * c.width * c.height * c.frames for c in clips)
        per_step_ms = ms / args.steps
        peak, peak_src = measured_peak_hbm()
        achieved = nbytes / world / (per_step_ms / 1e3) / 1e9
        line = {
            "metric": "cfg5 frames/s", "value": round(frames * args.steps / (ms / 1e3), 2), "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(per_step_ms, 4), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8/u16+f32", "data": "synthetic",
            "gpix_per_s": round(pix * args.steps / (ms / 1e3) / 1e9, 3),
            "config": {"workload": CFG5["desc"], "content": "natural", "clips": nclips, "frames_per_clip": fpc,
                       "distinct_luts": nclips, "assignment": "whole clips, longest-processing-time first",
                       "l2": "every clip's frames exceed L2", "sharding": "clips per GPU, no collective"},
            "roofline": {"bound": "hbm", "kernel": "k422_warp + k420_warp", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src, "per": "GPU"},
            "cpu_baseline": None,
            "e2e": {"value": round(frames * e2e_steps / dt, 2), "unit": "frames/s",
                    "h2d_bytes_per_step": int(nbytes / 2), "d2h_bytes_per_step": int(nbytes / 2), "steps": e2e_steps,
                    "api": "per clip: .cube read + parse + upload (cold cache, prefetched one clip ahead) + ColourStage.run_host (pinned host buffers)",
                    "seconds_by_part_rank0": {k: round(v, 3) for k, v in parts.items()}},
            "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    for st in stages.values():
        st.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS) + ["cfg5"])
    ap.add_argument("--content", default="natural", help="natural | noise | flat | grain<sigma> (e.g. grain0.02)")
    ap.add_argument("--float-mode", default="c", choices=["c", "fma"])
    ap.add_argument("--frames-per-step", type=int, default=0)
    ap.add_argument("--rows-per-cta", type=int, default=0)
    ap.add_argument("--min-seconds", type=float, default=1.05, help="a step repeats the launch so that the K timed steps last at least this long")
    ap.add_argument("--no-classes", action="store_true", help="skip the second content class (content_classes)")
    ap.add_argument("--opt", action="append", default=[], help="stage tunable name=value (see lutr_stage_set_option)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--in-fmt", default="", help="override the source layout of --config (any layout of include/lutr.h)")
    ap.add_argument("--out-fmt", default="", help="override the output layout of --config")
    ap.add_argument("--interp", default="", help="override the interpolation mode of --config")
    args = ap.parse_args()
    if args.config == "cfg5":
        return run_cfg5(args)
    cfg = dict(CONFIGS[args.config])
    if args.in_fmt or args.out_fmt or args.interp:  # other cells of the format matrix at this config's geometry
        cfg["pix_fmt"] = args.in_fmt or cfg["pix_fmt"]
        cfg["out_pix_fmt"] = args.out_fmt or cfg.get("out_pix_fmt", cfg["pix_fmt"])
        cfg["interp"] = args.interp or cfg["interp"]
        cfg["desc"] = f"{cfg['w']}x{cfg['h']} {cfg['pix_fmt']} -> {cfg['lut']}-point look .cube, {cfg['interp']} -> {cfg['out_pix_fmt']}"
        args.config = f"{args.config}[{cfg['pix_fmt']}->{cfg['out_pix_fmt']},{cfg['interp']}]"
    if args.frames_per_step:
        cfg["frames_per_step"] = args.frames_per_step
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return

    import torch
    import torch.distributed as dist
    from lut_renderer_b200.stage import ColourStage

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    from lut_renderer_b200.sharding import bind_to_device_cpus
    full_affinity = bind_to_device_cpus(local_rank)   # pinned host frames on the GPU's own NUMA node (one process per GPU)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(device))

    F = cfg["frames_per_step"]
    stage = ColourStage(cfg["w"], cfg["h"], cfg["pix_fmt"], out_pix_fmt=cfg.get("out_pix_fmt"), in_range=cfg["in_range"], norm_range=cfg["norm_range"],
                        interp=cfg["interp"], float_mode=args.float_mode, device=local_rank)
    if args.rows_per_cta:
        stage.set_option("rows_per_cta", args.rows_per_cta)
    for kv in args.opt:
        name, _, value = kv.partition("=")
        stage.set_option(name, int(value))
    table = make_lut_table(cfg["lut"])
    lut = stage.ctx.lut_from_table(table)
    # this rank's contiguous frame range of the (world * F)-frame clip, resident in HBM
    src, base_frames = device_frames(cfg, args.content, F, device, rank)
    dst = torch.empty(F * stage.out_frame_bytes, dtype=torch.uint8, device=device)
    # launch on a torch-visible stream so the CUDA events below bracket the kernels
    tstream = torch.cuda.Stream(device=device)
    torch.cuda.set_stream(tstream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(buf, steps, launches_per_step):
        """`steps` steps of `launches_per_step` launches over the frames in `buf`: (ms over all of them, launches, clocks)."""
        barrier()
        sampler = ClockSampler(local_rank)
        n0 = stage.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler.start()
        e0.record()
        for _ in range(steps * launches_per_step):
            stage.run(lut, buf, dst, F, stream=tstream)
        e1.record()
        barrier()
        clocks = sampler.stop()
        return reduce_max(e0.elapsed_time(e1)), stage.launch_count - n0, clocks

    # warm-up, then size a step: one launch processes F frames; a step repeats it so that the K timed steps last >= --min-seconds
    # (a 2.4 ms launch times 100 steps would give NVML three clock samples)
    for _ in range(args.warmup):
        stage.run(lut, src, dst, F, stream=tstream)
    ms_probe, _, _ = timed(src, 1, 3)
    per_launch_probe = max(ms_probe / 3, 1e-3)
    L = max(1, int(np.ceil(args.min_seconds * 1e3 / (args.steps * per_launch_probe))))
    if world > 1:
        lt = torch.tensor([L], dtype=torch.int64, device=device)
        dist.all_reduce(lt, op=dist.ReduceOp.MAX)
        L = int(lt.item())
    ms, launches, clocks = timed(src, args.steps, L)
    if world > 1:
        lt = torch.tensor([launches], dtype=torch.int64, device=device)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    per_launch_ms = ms / (args.steps * L)

    # spot check: the timed outputs are the right pixels (first frame of this rank against the oracle)
    parity, parity_planes = None, {}
    if rank == 0:
        try:
            from oracle import oracle as O
            from lut_renderer_b200 import synth
            got = dst[:stage.out_frame_bytes].cpu().numpy()
            first = src[:stage.in_frame_bytes].cpu().numpy()
            L_or = O.OracleLut.from_table(table)
            Y, U, V = [np.ascontiguousarray(p) for p in synth.split_planes(first, cfg["pix_fmt"], cfg["w"], cfg["h"])]
            O.set_float_mode(args.float_mode)   # the oracle evaluates the blend in the same order as the stage was asked to
            try:
                exp = O.stage_any(L_or, cfg["pix_fmt"], cfg.get("out_pix_fmt", cfg["pix_fmt"]), Y, U, V, interp=cfg["interp"],
                                  range_pass=cfg["in_range"] == "pc" and cfg["norm_range"] == "tv")
            finally:
                O.set_float_mode("c")
            parity = "bit-exact" if np.array_equal(got, synth.pack_planes(exp)) else "MISMATCH"
            # north_star's statement per plane: share of identical samples, max and mean absolute error (in code values)
            out_fmt = cfg.get("out_pix_fmt") or cfg["pix_fmt"]
            for name, a, b_ in zip("YUV", synth.split_planes(got, out_fmt, cfg["w"], cfg["h"]), exp):
                dlt = np.abs(a.astype(np.int32) - b_.astype(np.int32))
                parity_planes[name] = {"identical": round(float((dlt == 0).mean()), 6), "max_abs": int(dlt.max()), "mean_abs": round(float(dlt.mean()), 6)}
        except Exception as exc:  # the checker is optional for the measurement
            parity = f"unchecked ({exc})"

    bytes_per_frame = (BYTES_PER_PIXEL[cfg["pix_fmt"]] + BYTES_PER_PIXEL[cfg.get("out_pix_fmt") or cfg["pix_fmt"]]) * cfg["w"] * cfg["h"]
    peak, peak_src = measured_peak_hbm()

    # the other content class (A = natural, B = worst-case gather): same stage, same launch shape, a shorter region
    classes = {}
    if not args.no_classes:
        classes[args.content] = {"value": round(world * F * args.steps * L / (ms / 1e3), 2),
                                 "frac": round(bytes_per_frame * F / (per_launch_ms / 1e3) / 1e9 / peak, 4)}
        for other in ("natural", "noise"):
            if other in classes:
                continue
            src_o, _ = device_frames(cfg, other, F, device, rank)
            for _ in range(3):
                stage.run(lut, src_o, dst, F, stream=tstream)
            n_o = max(3, int(np.ceil(0.4e3 / per_launch_probe / 3)))  # >= 0.4 s even if this class ran as fast as the main one
            ms_o, _, _ = timed(src_o, 1, n_o)
            classes[other] = {"value": round(world * F * n_o / (ms_o / 1e3), 2),
                              "frac": round(bytes_per_frame * F / (ms_o / n_o / 1e3) / 1e9 / peak, 4)}
            del src_o

    # end to end: the call a user makes (ColourStage.run_host) on pinned host frames, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        # the same F frames per call at every N (4.2 GB pinned per rank and direction at 4K 10-bit): fill and drain of the 3-slot
        # device ring stay amortised the same way whatever the number of ranks
        Fe = F
        pin_in = stage.ctx.pinned_buffer(Fe * stage.in_frame_bytes)
        pin_out = stage.ctx.pinned_buffer(Fe * stage.out_frame_bytes)
        for i in range(Fe):
            pin_in.array[i * stage.in_frame_bytes:(i + 1) * stage.in_frame_bytes] = base_frames[i % len(base_frames)]
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            stage.run_host(lut, pin_in, pin_out, Fe)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            stage.run_host(lut, pin_in, pin_out, Fe)   # synchronous: output bytes are in host memory on return
        torch.cuda.synchronize()
        dt = reduce_max(time.perf_counter() - t0)
        e2e = {"value": round(world * Fe * e2e_steps / dt, 2), "unit": "frames/s", "h2d_bytes_per_step": Fe * stage.in_frame_bytes,
               "d2h_bytes_per_step": Fe * stage.out_frame_bytes, "frames_per_step": Fe, "steps": e2e_steps,
               "api": "ColourStage.run_host -> lutr_stage_run_host (pinned host buffers)",
               "cpu_affinity": "cores local to the GPU (NVML)" if full_affinity else "unchanged"}
        # what the bus allows: pinned copies in both directions at once, ALL ranks copying at the same time (barrier first)
        try:
            barrier()
            pcie = pcie_simultaneous_gbs(device)
            moved = max(stage.in_frame_bytes, stage.out_frame_bytes) * Fe * e2e_steps / dt / 1e9
            e2e["pcie_each_direction_gbs"] = round(pcie, 1)
            e2e["frac_of_pcie"] = round(moved / pcie, 3)
        except Exception:
            pass
        pin_in.close()
        pin_out.close()

    if rank == 0:
        fps = world * F * args.steps * L / (ms / 1e3)
        mpix = cfg["w"] * cfg["h"] / 1e6
        achieved = bytes_per_frame * F / (per_launch_ms / 1e3) / 1e9
        kernel = stage.kernel_name
        cpu = {"value": None, "unit": "frames/s", "cores": 0, "kind": "port", "sample": "measured at N=1 only (rank 0 of a one-GPU run)"}
        if args.no_cpu_baseline:
            cpu = None
        elif world == 1:
            from oracle import cpu_baseline as C
            from oracle import oracle as O
            if full_affinity:
                os.sched_setaffinity(0, full_affinity)   # the CPU chain gets every host core again
            r = C.time_cpu_path(cfg["pix_fmt"], cfg["w"], cfg["h"], O.OracleLut.from_table(table), base_frames,
                                seconds=args.cpu_seconds, interp=cfg["interp"], out_pix_fmt=cfg.get("out_pix_fmt"))
            cpu = {"value": round(r["fps"], 3), "unit": "frames/s", "cores": r["threads"], "kind": "port",
                   "sample": f"{r['legs']}; {r['frames']} frames of the same workload in {r['seconds']:.1f}s, "
                             f"{r['threads']} frame-parallel threads"}
        line = {
            "metric": "4K 10-bit frames/s" if args.config == "cfg3" else f"{args.config} frames/s",
            "value": round(fps, 2), "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u16+f32" if "10" in cfg["pix_fmt"] else "u8+f32", "data": "synthetic",
            "gpix_per_s": round(fps * mpix / 1e3, 3),
            "launches_per_step": L, "frames_per_step": F * L, "ms_per_launch": round(per_launch_ms, 4), "timed_region_s": round(ms / 1e3, 3),
            "config": config_dict(cfg, args, world),
            "parity": {"spot_check": parity, "vs_oracle_first_frame": parity_planes},
            "content_classes": classes,
            "roofline": {"bound": "hbm", "kernel": kernel, "achieved": round(achieved, 1),
                         "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": ncu_traffic(args.config, F), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(bytes_per_frame * F), "frac_of_8TBs_nominal": round(achieved / 8000.0, 4), "per": "GPU"},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    stage.close()
    lut.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
