#!/usr/bin/env python
"""bench.py -- 4K frames/s of the extract hot path (letterbox -> YOLOv8s detect+NMS -> ORB/match/RANSAC -> box warp).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dtype fp16|bf16]

One "step" = one batch of 16 synthetic 3840x2160 BGR frames through the whole path (BASELINE.json configs[1]+[2] fused,
the default preset of /root/reference/geotrax/cfg/default.yaml).  Prints ONE JSON line (rank 0).  N > 1 is launched by
torchrun, one rank per GPU; frames are sharded by contiguous range (each rank owns its own frames: weak scaling) and
the per-frame boxes + homographies are gathered to rank 0 over NCCL inside the timed region.

`--impl reference` times the CPU oracle port of the same path (the reference's engines are CPU OpenCV + fp32 PyTorch;
the packages ultralytics/stabilo themselves are not installable offline -- DESIGN.md) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAME_HW = (2160, 3840)
IMGSZ = 1920
BATCH = 16
CONF, IOU = 0.25, 0.7
CLS_BIAS = -4.4           # random-init class bias giving ~300 candidates / frame at conf 0.25 (golden: 127-136 kept / frame)
METRIC = "4K frames/sec detect+stabilize"
WORKLOAD_FUSED = ("detect+stabilize (configs[1]+[2] fused): 16 x 3840x2160 synthetic BGR frames / step, YOLOv8s nc=4 random-init "
                  "imgsz 1920 (1088x1920), conf 0.25 iou 0.7 agnostic, ORB 2000/4000 + Hamming 2-NN + 5000-hyp RANSAC, box warp")
UNIT = "frames/s"
# algorithmic work per 4K frame (SURVEY.md 8d / BASELINE.md section 2)
CONV_GFLOP_PER_FRAME = 145.03


def _conv_traffic():
    """DRAM bytes of the conv launches of one step, from the committed ncu capture (profiles/conv_traffic_r2.json)."""
    p = os.path.join(ROOT, "profiles", "conv_traffic_r2.json")
    try:
        with open(p) as f:
            return float(json.load(f)["traffic_bytes_per_step"])
    except Exception:
        return None


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d.get("hbm_gbs", 6650.0), tf_burst=d.get("bf16_tflops", 1590.0), tf_sust=d.get("bf16_tflops_sustained", 1400.0), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.times, self._halt = index, [], [], threading.Event()
        # In-process NVML (nvidia-ml-py) when it is importable: spawning `nvidia-smi` every 200 ms from every rank costs host time and takes
        # driver locks inside the timed region (measured: a 60-step 2-GPU run was 10 % slower per step than a 20-step one).  Same fields.
        self._nvml = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            h = None
            uuid = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)      # CUDA ordinal -> NVML handle (the orders can differ)
            for u in (uuid, uuid.encode()):
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(u)
                    break
                except Exception:
                    pass
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._nvml = (pynvml, h)
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        nv, h = self._nvml
        sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
        try:
            rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        flag = lambda bit: "Active" if rs & bit else "Not Active"
        return [str(sm), str(mx), f"{pw:.2f}", flag(nv.nvmlClocksEventReasonHwSlowdown), flag(nv.nvmlClocksEventReasonHwThermalSlowdown),
                flag(nv.nvmlClocksEventReasonSwThermalSlowdown), flag(nv.nvmlClocksEventReasonSwPowerCap)]

    def run(self):
        while not self._halt.is_set():
            try:
                if self._nvml is not None:
                    self.rows.append(self._sample_nvml())
                    self.times.append(time.perf_counter())
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([c.strip() for c in out.split(",")])
                        self.times.append(time.perf_counter())
            except Exception:
                pass
            self._halt.wait(0.05 if self._nvml is not None else 0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons), samples=len(self.rows),
                    power_w_median=float(np.median(pw)) if pw else None)

    def windows(self, t0: float, width: float = 1.0):
        """Per-`width`-second medians of SM clock and power since t0 (sustained-run report)."""
        out = {}
        for t, r in zip(self.times, self.rows):
            k = int((t - t0) // width)
            if k < 0 or not r[0].replace(".", "").isdigit():
                continue
            out.setdefault(k, []).append((float(r[0]), float(r[2]) if r[2].replace(".", "").isdigit() else float("nan")))
        return {k: dict(sm_mhz=float(np.median([a for a, _ in v])), power_w=float(np.nanmedian([b for _, b in v]))) for k, v in sorted(out.items())}


# ------------------------------------------------------------------------------------------------------------------------
# CPU oracle port (cpu_baseline leg and --impl reference)
# ------------------------------------------------------------------------------------------------------------------------
class CpuPath:
    def __init__(self, sd, threads: int):
        import cv2
        import torch
        from oracle.stabilo_cv import Stabilizer
        from oracle.yolov8 import YOLOv8

        torch.set_num_threads(threads)
        cv2.setNumThreads(threads)
        self.model = YOLOv8(4, "detect").eval()
        self.model.load_state_dict(sd, strict=False)
        self.stab = Stabilizer()
        self.have_ref = False
        self.torch = torch

    def frame(self, frame):
        """One iteration of /root/reference/geotrax/extract.py:145-197 minus decode and tracker."""
        from oracle import prepost
        x = prepost.preprocess([frame], IMGSZ)
        with self.torch.no_grad():
            dec, _ = self.model(x)
        det = prepost.postprocess_detect(dec, x.shape[2:], frame.shape[:2], CONF, IOU, [0, 1, 2, 3], True, 1000)[0]
        xywh = prepost.xyxy2xywh(det[:, :4]).numpy() if len(det) else None
        if not self.have_ref:
            self.stab.set_ref_frame(frame, xywh)
            self.have_ref = True
            return det, np.eye(3)
        self.stab.stabilize(frame, xywh)
        self.stab.transform_cur_boxes()
        return det, self.stab.get_cur_trans_matrix()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from geotrax_b200 import synth, weights
    cores = os.cpu_count() or 1
    sd = weights.random_state_dict(4, "detect", seed=0, frame_hw=FRAME_HW, imgsz=IMGSZ, cls_bias=CLS_BIAS)
    frames = synth.make_flight(3, FRAME_HW[0], FRAME_HW[1], seed=100)[0]
    cpu = CpuPath(sd, cores)
    cpu.frame(frames[0])                       # reference frame
    per_step = 2                               # bounded sample: 2 of the 16 frames of a step
    for _ in range(args.warmup):
        cpu.frame(frames[1])
    t0 = time.perf_counter()
    for s in range(args.steps):
        for j in range(per_step):
            cpu.frame(frames[1 + (s + j) % 2])
    dt = time.perf_counter() - t0
    fps = args.steps * per_step / dt
    # same `config` keys as the GPU arm's line (workload, frames_per_step, parallelism): the step is the same 16-frame workload, of which
    # this arm processes a bounded sample (cpu_baseline.sample) -- the CPU path is batch 1 like extract.py, frames/s is what compares
    line = dict(metric=METRIC, value=fps, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=1000 * dt / args.steps,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=WORKLOAD_FUSED, frames_per_step=BATCH, parallelism=f"frame-range shard x{args.gpus}"),
                cpu_baseline=dict(value=fps, unit=UNIT, cores=cores, kind="port",
                                  sample=f"{per_step} of the step's {BATCH} frames per step x {args.steps} steps, batch 1 like extract.py; "
                                         "CPU oracle port (fp32 PyTorch YOLOv8s + OpenCV ORB / BFMatcher / findHomography(USAC_MAGSAC)) on all host cores"),
                e2e=dict(value=fps, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import geotrax_b200
    from geotrax_b200 import pipeline, synth, weights

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints its version banner on fd 1 at the first communicator init; keep stdout to the one JSON line
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    dev = torch.device("cuda", local)

    task = "obb" if args.workload == "obb" else "detect"
    eng = geotrax_b200.Engine(frame_hw=FRAME_HW, imgsz=IMGSZ, nc=4, task=task, max_batch=BATCH, device=local, act_dtype=args.dtype)
    sd = weights.random_state_dict(4, task, seed=0, frame_hw=FRAME_HW, imgsz=IMGSZ, cls_bias=CLS_BIAS)
    eng.load_weights(weights.fold(sd, 4, task))
    # this rank's contiguous frame range of the synthetic flight: BATCH distinct frames (cycled), + the shared reference frame
    flight = synth.make_flight(BATCH, FRAME_HW[0], FRAME_HW[1], seed=100 + rank)
    frames_np = np.stack(flight[0])
    ref_np = frames_np[:1].copy()
    # vehicle masks = the generator's 132 golden-like boxes per frame ("dense vehicle masks", configs[2]); in the reference
    # they are the tracker's boxes (extract.py:166,181) -- a random-init detector's own boxes are meaningless as masks
    # (tests/test_gpu_round2.py::test_mask_source_moves_centres_less_than_half_a_pixel bounds what the mask source changes)
    mask = eng.pack_boxes(flight[1])
    mask_list = flight[1]
    mask_dev = (torch.from_numpy(mask[0]).to(dev), torch.from_numpy(mask[1]).to(dev))
    mask_ref = eng.pack_boxes(flight[1][:1])
    frames_dev = torch.from_numpy(frames_np).to(dev)
    # two pinned host batches (the second is the first rotated by one frame) so that consecutive e2e steps copy different bytes
    host_batches = [frames_np, np.roll(frames_np, 1, axis=0).copy()]
    if args.ingest == "nv12":   # the e2e leg ships decoder-format frames; the resident leg keeps BGR24 frames in HBM
        host_batches = [np.stack([synth.bgr_to_nv12(f) for f in hb]) for hb in host_batches]
    frames_pin = [torch.from_numpy(hb).pin_memory() for hb in host_batches]
    mask_roll = eng.pack_boxes([flight[1][(i - 1) % BATCH] for i in range(BATCH)])
    pin = lambda pair: tuple(torch.from_numpy(a).pin_memory() for a in pair)
    mask_pin, mask_roll_pin = pin(mask), pin(mask_roll)
    out = eng.alloc_outputs(pinned=True)
    tstream = torch.cuda.Stream(device=dev, priority=-1)     # every kernel / copy of the path is issued on this (high-priority) stream; events time it
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    eng.extract_batch(torch.from_numpy(ref_np).to(dev), first_is_reference=True, conf=CONF, iou=IOU, classes=[0, 1, 2, 3], out=out, stream=stream,
                      mask_boxes=mask_ref)
    engines = [eng]
    for _ in range(max(args.engines, 1) - 1):   # optional: further handles on the same GPU taking the batches round-robin (pipeline.run_range)
        e2 = geotrax_b200.Engine(frame_hw=FRAME_HW, imgsz=IMGSZ, nc=4, task=task, max_batch=BATCH, device=local, act_dtype=args.dtype)
        e2.load_weights(weights.fold(sd, 4, task))
        e2.extract_batch(torch.from_numpy(ref_np).to(dev), first_is_reference=True, conf=CONF, iou=IOU, classes=[0, 1, 2, 3], stream=stream, mask_boxes=mask_ref)
        engines.append(e2)
    fused = args.workload in ("fused", "obb", "flight")
    pipelined = fused and not args.no_pipeline
    # one engine: everything on the timed torch stream; several engines: each on its own (high-priority) stream so that they overlap --
    # every gt_wait has returned before the closing event is recorded, so the event pair still brackets all of the work
    det_kw = dict(conf=CONF, iou=IOU, agnostic=True, classes=[0, 1, 2, 3], stream=stream if len(engines) == 1 else None)

    def single_stage_step(src):
        """configs[1] / configs[2] alone (N = 1 diagnostics; not the sharded driver)"""
        if args.workload == "detect":        # letterbox + detector + decode/NMS, boxes read back
            eng.preprocess(src, stream=stream)
            bx, cnt = eng.detect(int(src.shape[0]), conf=CONF, iou=IOU, agnostic=True, classes=[0, 1, 2, 3], stream=stream)
            out["boxes"][...] = bx
            out["counts"][...] = cnt
        else:                                # gray/half-res + ORB + match + RANSAC against the reference frame, H read back
            eng.preprocess(src, stream=stream)
            Hm, st_, stats = eng.stabilize(int(src.shape[0]), mask_list, stream=stream)
            out["H"][...] = Hm.reshape(-1, 9)
            out["status"][...] = st_
            out["stats"][...] = stats

    phase = [0]

    def timed(host: bool, steps: int, seconds: float = 0.0):
        """`steps` batches of BATCH frames through the PRODUCT's sharded driver: pipeline.run_range (two batches in flight, prefetched
        H2D for host frames) on this rank's frames, then pipeline.gather_records to rank 0 (NCCL) -- all inside the timed region."""
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = sum(e.launch_count() for e in engines)
        stage, conv, marks, gather_s = np.zeros(4), [0.0], [], [0.0]

        def account(b0=0, b1=0):
            st = eng.stage_times()
            stage[:] += [st["preprocess"], st["inference"], st["postprocess"], st["stabilize"]]
            conv[0] += eng.conv_stack_stats()[0]
            marks.append(time.perf_counter())

        if host:   # the two pinned batches alternate across rounds too, so that a round's first batch is the one the previous round prefetched
            get_frames = lambda a, b: frames_pin[(a // BATCH + phase[0]) % 2]
            get_masks = lambda a, b: (mask_pin if (a // BATCH + phase[0]) % 2 == 0 else mask_roll_pin)
        else:
            get_frames = lambda a, b: frames_dev
            get_masks = lambda a, b: mask_dev
        t0 = time.perf_counter()
        e0.record()
        done = 0
        while True:
            if fused:
                # steady-state ingest: every step starts the H2D copy of the NEXT batch (the last one that of the following round's first
                # batch), so the timed region holds exactly `steps` copies of 398 MB; the first batch was put in flight by the previous round
                rec_local = pipeline.run_range(engines if len(engines) > 1 else eng, get_frames, 0, steps * BATCH, 0, batch=BATCH, get_masks=get_masks, set_reference=False,
                                               pipelined=pipelined, on_batch=account,
                                               next_range_frames=get_frames(steps * BATCH, (steps + 1) * BATCH) if host else None, **det_kw)
                tg = time.perf_counter()
                pipeline.gather_records(rec_local, rank, world, dev)
                gather_s[0] += time.perf_counter() - tg
            else:
                for i in range(steps):
                    single_stage_step(frames_pin[i % 2] if host else frames_dev)
                    account()
            done += steps
            if host:
                phase[0] += steps
            if time.perf_counter() - t0 >= seconds:
                break
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        ms_min = ms
        if world > 1:
            t = torch.tensor([ms, wall * 1000, gather_s[0] * 1000, -ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall, gather_s[0], ms_min = float(t[0]), float(t[1]) / 1000, float(t[2]) / 1000, -float(t[3])
        return dict(ms=ms, wall=wall, gather_ms=gather_s[0] * 1000, ms_fastest_rank=ms_min, stage=stage / done, conv_ms=conv[0] / done, launches=sum(e.launch_count() for e in engines) - l0, steps=done, t0=t0, marks=marks)

    timed(False, max(args.warmup, 3))
    sampler = ClockSampler(local)
    sampler.start()
    res = timed(False, args.steps, args.seconds if args.workload == "flight" else 0.0)          # inputs resident in HBM
    clocks = sampler.stop()
    steps_done = res["steps"]
    ms, stage, conv_ms, launches = res["ms"], res["stage"], res["conv_ms"], res["launches"]
    flight_windows = None
    if args.workload == "flight":       # sustained run: frames/s, SM clock and power per one-second window
        win = sampler.windows(res["t0"])
        per = {}
        for t in res["marks"]:
            per[int(t - res["t0"])] = per.get(int(t - res["t0"]), 0) + BATCH
        flight_windows = [dict(second=k, frames_per_s=per.get(k, 0), **win.get(k, {})) for k in sorted(per)]
    for e in engines:
        e.set_input_format(args.ingest)
    timed(True, 2)
    res2 = timed(True, args.steps)               # pinned host frames: H2D inside the timed region
    ms_e2e, wall_e2e = res2["ms"], res2["wall"]
    if args.workload == "detect":
        stage[3] = 0.0
    elif args.workload == "stabilize":
        stage[1] = stage[2] = 0.0
        conv_ms = 0.0
    det_counts = out["counts"].copy() if not fused else None
    health = eng.health()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = _peaks()
    total_frames = world * steps_done * BATCH
    value = total_frames / (ms / 1000)
    e2e = world * args.steps * BATCH / max(wall_e2e, ms_e2e / 1000)
    gflop_frame = 150.04 if args.workload == "obb" else CONV_GFLOP_PER_FRAME   # +cv4 branch for OBB
    conv_tflops = gflop_frame * BATCH / conv_ms if conv_ms > 0 else 0.0  # GFLOP / ms = TFLOP/s
    wl_name = {"fused": "detect+stabilize (configs[1]+[2] fused)", "detect": "detect+NMS only (configs[1])",
               "stabilize": "ORB+match+RANSAC homography only (configs[2])", "obb": "YOLOv8s-OBB rotated NMS + stabilization (configs[3])",
               "flight": "sustained detect+stabilize flight (configs[4] per-GPU share)"}[args.workload]
    metric = METRIC if args.workload in ("fused", "flight") else f"4K frames/sec {args.workload}"
    h2d = int(host_batches[0].nbytes)
    d2h = int(sum(v.nbytes for v in out.values()))
    h2d += int(mask[0].nbytes + mask[1].nbytes)
    cpu_base = None
    if world == 1 and not args.no_cpu_baseline and args.workload == "fused":
        cores = os.cpu_count() or 1
        cpu = CpuPath(sd, cores)
        cpu.frame(ref_np[0])
        cpu.frame(frames_np[0])
        n = 6
        t0 = time.perf_counter()
        for i in range(n):
            cpu.frame(frames_np[1 + i % (BATCH - 1)])
        cdt = time.perf_counter() - t0
        cpu_base = dict(value=n / cdt, unit=UNIT, cores=cores, kind="port", sample=f"{n} of the step's {BATCH} frames, batch 1, fp32 PyTorch + OpenCV")
    # per-stage roofline fractions: algorithmic bytes / flops per frame from SURVEY.md 8d (preprocess 39.49 MB = 24.88 read + 12.53 net input
    # + 2.07 gray) and DESIGN.md 4; peaks measured (MEASURED_PEAKS.json)
    per_frame = dict(preprocess=("hbm", 39.49e6), inference=("tensor", gflop_frame * 1e9), postprocess=("hbm", 5.83e6), stabilize=("hbm", 30.0e6))
    stages = {}
    for name, ms_stage in zip(("preprocess", "inference", "postprocess", "stabilize"), stage):
        kind, work = per_frame[name]
        if ms_stage <= 0:
            continue
        rate = work * BATCH / (ms_stage * 1e-3)
        pk = peaks["tf_sust"] * 1e12 if kind == "tensor" else peaks["hbm"] * 1e9
        stages[name] = dict(ms_per_step=float(ms_stage), bound=kind, achieved=rate / (1e12 if kind == "tensor" else 1e9),
                            unit="TFLOP/s" if kind == "tensor" else "GB/s", frac=rate / pk)
        if kind == "tensor":
            stages[name]["frac_vs_burst_peak"] = rate / (peaks["tf_burst"] * 1e12)
    # which peak applies: the burst figure for a kernel timed alone / a short run at full clocks, the sustained one inside a long step
    sm = clocks.get("sm_mhz") or 0.0
    use_burst = sm >= 0.95 * (clocks.get("sm_max_mhz") or 1e9)
    peak_tf = peaks["tf_burst"] if use_burst else peaks["tf_sust"]
    roof = dict(bound="tensor", achieved=conv_tflops, peak=peak_tf, unit="TFLOP/s", frac=conv_tflops / peak_tf, traffic=_conv_traffic(),
                frac_vs_sustained_peak=conv_tflops / peaks["tf_sust"], frac_vs_burst_peak=conv_tflops / peaks["tf_burst"],
                kernel="conv_tc_kernel / conv_sw_kernel launch set (the conv stack of one 16-frame step; traffic = dram bytes of that set, ncu)",
                peak_source=peaks["src"] + (" bf16_tflops (burst: the run sat at max SM clock)" if use_burst else " bf16_tflops_sustained (SM clock below max)"),
                algorithmic_flops_per_launch_set=gflop_frame * BATCH * 1e9)
    if args.workload == "stabilize":
        st = stages.get("stabilize", dict(achieved=0.0, frac=0.0))
        roof = dict(bound="hbm", achieved=st["achieved"], peak=peaks["hbm"], unit="GB/s", frac=st["frac"], traffic=None,
                    kernel="ORB pyramid + FAST + select + describe + match + RANSAC launch set of one 16-frame step",
                    peak_source=peaks["src"] + " hbm_gbs", algorithmic_bytes_per_launch_set=30.0e6 * BATCH)
    cfg = dict(workload=WORKLOAD_FUSED if args.workload == "fused" else
               wl_name + ": 16 x 3840x2160 synthetic BGR frames / step, YOLOv8s nc=4 random-init imgsz 1920 (1088x1920), conf 0.25 iou 0.7 agnostic, "
                         "ORB 2000/4000 + Hamming 2-NN + 5000-hyp RANSAC, box warp",
               frames_per_step=BATCH, parallelism=f"frame-range shard x{world}",
               driver="pipeline.run_range + gather_records (the product's sharded driver)" if fused else "single-stage loop", engines_per_gpu=len(engines),
               pipeline=("two batches in flight (gt_extract_batch_async)" if pipelined else "synchronous"), l2="inputs (398 MB / step) larger than the 126 MB L2",
               stage_ms_per_step=dict(preprocess=stage[0], inference=stage[1], postprocess=stage[2], stabilize=stage[3]),
               mask_boxes_per_frame=float(mask[1].mean()), storage_dtype=args.dtype, nonfinite_head_rows=health,
               conv_launches_per_step=eng.conv_kernel_info()[0], convs_on_swapped_kernel=eng.conv_kernel_info()[1],
               conv_variant_choice="fixed table / rule keyed by layer signature (identical in every process)")
    if det_counts is not None:
        cfg["detections_per_frame"] = float(det_counts.mean())
    if world > 1:   # where the timed region of a multi-rank run goes besides the ranks' own frames (value uses the slowest rank, as the contract says)
        cfg["multi_rank"] = dict(gather_ms_per_flight=res["gather_ms"], ms_per_step_fastest_rank=res["ms_fastest_rank"] / steps_done,
                                 ms_per_step_slowest_rank=ms / steps_done)
    line = dict(metric=metric, value=value, unit=UNIT, n_gpus=world, steps=steps_done, warmup=max(args.warmup, 3), ms_per_step=ms / steps_done,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype=args.dtype, data="synthetic", config=cfg,
                roofline=roof, stages=stages, cpu_baseline=cpu_base,
                e2e=dict(value=e2e, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, ms_per_step=1000 * wall_e2e / args.steps,
                         ingest=args.ingest),
                gpu_launches=int(launches), clocks=clocks)
    if flight_windows is not None:
        line["flight"] = dict(seconds=res["wall"], frames=total_frames, windows=flight_windows)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="synchronous gt_extract_batch per step instead of the two-deep gt_extract_batch_async pipeline")
    ap.add_argument("--ingest", default="bgr24", choices=["bgr24", "nv12"],
                    help="host frame format of the end-to-end leg: bgr24 = what the reference's reader delivers (default, the headline); "
                         "nv12 = decoder format (SURVEY 8f rank 1), half the PCIe bytes, converted on the device")
    ap.add_argument("--engines", type=int, default=2, help="handles per GPU taking the batches round-robin (pipeline.run_range): while one batch is in its low-occupancy "
                    "stabiliser tail (selection, matching, RANSAC: 16-128 blocks) the other handle's conv CTAs fill the SMs: +6.8 %% frames/s with two (r2 final build; "
                    "the conv-stack time the roofline uses is unchanged: 3.27 ms either way).  1 = a single handle; stage_ms_per_step is only meaningful then")
    ap.add_argument("--seconds", type=float, default=15.0, help="--workload flight: keep running whole `--steps` rounds until this much wall time has passed")
    ap.add_argument("--workload", default="fused", choices=["fused", "detect", "stabilize", "obb", "flight"],
                    help="fused = BASELINE configs[1]+[2] (default, the headline); detect = configs[1] (YOLOv8s detect+NMS only); "
                         "stabilize = configs[2] (ORB + match + RANSAC + warp only); obb = configs[3] (YOLOv8s-OBB rotated NMS + stabilization); "
                         "flight = the fused path sustained for --seconds with per-second frames/s, SM clock and power (configs[4] per-GPU share)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
