#!/usr/bin/env python
"""bench.py -- disparity volumes/s of the STEREO_GIF hot path (CVC -> CVF -> WTA) on B200.

Contract (see the task brief): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line
on rank 0.  A "step" is one stereo frame through CostConst_GPU + CostFilter_GPU + DispSelect_GPU
producing BOTH disparity maps.  Workload at every N: BASELINE config C4, synthetic 1920x1080, D=128,
fp32.  N>1 (launched by torchrun, one rank per GPU): the disparity axis is sharded D/N slices per
rank, one NCCL all-gather of packed per-pixel (cost,d) minima per view, then the final min -> u8 maps
(strong scaling: the frame is fixed, `value` = frames/s of the whole job).

  value : frames/s with the interleaved f32 images already resident in HBM (device-timed, CUDA events)
  e2e   : same metric through the host-facing C-ABI calls: images in pinned HOST memory, H2D of both
          images and D2H of both u8 maps inside the timed region
  roofline : the fused CVF kernel's algorithmic bytes / its event-timed duration vs the measured HBM peak
  cpu_baseline / --impl reference : the CPU oracle (C restatement of the reference's pthreads path;
          the reference itself cannot be built here: no OpenCV C++ headers / CL/cl.h) on the host cores
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {  # BASELINE.json configs; C4 is the one the metric is quoted on and the default at every N
    "C3": (1280, 720, 64), "C4": (1920, 1080, 128), "C5": (1920, 1080, 256),
}
W, H, D = WORKLOADS["C4"]
WORKLOAD = "C4 synthetic 1920x1080 D=128 fp32, both views (lDisMap+rDisMap)"
METRIC = "disparity_volumes_per_s"
UNIT = "volumes/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 7 and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def oracle_frame(l, r, threads, sample_slices=None):
    """One frame (or a slice sample of it) through the CPU oracle -> seconds per FULL frame."""
    from oracle import oracle as O
    Ds = D if sample_slices is None else sample_slices
    t0 = time.perf_counter()
    res = O.pipeline(l, r, Ds, threads=threads)
    dt = time.perf_counter() - t0
    # per-slice cost is uniform (every slice is the same CVC + GIF work; WTA is 1% of the frame)
    return dt * (D / Ds), res["times_ms"]


def run_reference(args, rank):
    """--impl reference: the reference's CPU path (oracle port) on the host cores, rank 0 only."""
    if rank != 0:
        return
    from primestereomatch_b200 import synth
    l, r, _ = synth.stereo_pair_f32(W, H, D)
    cores = os.cpu_count() or 1
    threads = min(cores, D)
    n = args.steps + args.warmup
    # bounded sample: the whole frame when the run is short, else a D/8-slice sample of it
    sample = None if n <= 8 else max(8, D // 8)
    for _ in range(args.warmup):
        oracle_frame(l, r, threads, sample)
    ts = []
    for _ in range(args.steps):
        t, _ = oracle_frame(l, r, threads, sample)
        ts.append(t)
    sec = float(np.mean(ts))
    val = 1.0 / sec
    what = "full frame" if sample is None else f"{sample} of {D} disparity slices of both views, scaled x{D / sample:g}"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "threads": threads,
                   "note": "CPU oracle = C restatement of the reference pthreads path (reference needs OpenCV+OpenCL headers: unbuildable here); "
                           "threads = all host cores (the reference itself caps at MAX_CPU_THREADS=8)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": what},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cvf-mode", type=int, default=0)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="C4", choices=sorted(WORKLOADS),
                    help="C4 (default) is the benchmark; C3 / C5 are the other synthetic BASELINE configs")
    ap.add_argument("--seg-rows", type=int, default=0)
    ap.add_argument("--extra-smem", type=int, default=0)
    ap.add_argument("--cta-threads", type=int, default=0)
    ap.add_argument("--remap", type=int, default=0)
    ap.add_argument("--upload", default="banded", choices=["banded", "replicated"],
                    help="N>1, e2e: banded = every rank uploads H/N rows of both images and the bands are "
                         "all-gathered over NVLink; replicated = every rank uploads both full images over PCIe")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: p2p = WTA kernel stores its minima into every rank's buffer over NVLink (fused "
                         "compute+exchange); nccl = local WTA then ncclAllGather")
    args = ap.parse_args()
    global W, H, D, WORKLOAD
    W, H, D = WORKLOADS[args.workload]
    WORKLOAD = f"{args.workload} synthetic {W}x{H} D={D} fp32, both views (lDisMap+rDisMap)"

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return 0

    import torch
    import torch.distributed as dist
    from primestereomatch_b200 import DispEst, capi, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert world == args.gpus or world == 1, (world, args.gpus)
    warm = max(3, args.warmup)

    L = capi.lib()
    l8, r8, _ = synth.stereo_pair_u8(W, H, D)
    l, r = synth.to_f32(l8), synth.to_f32(r8)
    lp8 = torch.from_numpy(l8).pin_memory()
    rp8 = torch.from_numpy(r8).pin_memory()
    lp = torch.from_numpy(l).pin_memory()
    rp = torch.from_numpy(r).pin_memory()
    ld_dev, rd_dev = lp.cuda(), rp.cuda()
    torch.cuda.synchronize()

    d_count = D // world
    d_begin = rank * d_count
    if rank == world - 1:
        d_count = D - d_begin
    de = DispEst(l, r, D, 8, True, device=local_rank, d_begin=d_begin, d_count=d_count)
    de.set_option(capi.PSM_OPT_CVF_MODE, args.cvf_mode)
    de.set_option(capi.PSM_OPT_VARIANT, args.variant)
    de.set_option(101, args.seg_rows)
    de.set_option(102, args.extra_smem)
    de.set_option(103, args.cta_threads)
    de.set_option(104, args.remap)
    stream = torch.cuda.Stream()  # a real (non-default) stream: handle 0 would mean "context's own stream"
    torch.cuda.set_stream(stream)
    capi.check(L.psm_set_stream(de.handle, C.c_void_p(stream.cuda_stream)), de.handle)

    npix = W * H
    lmap = torch.empty((H, W), dtype=torch.uint8).pin_memory()
    rmap = torch.empty((H, W), dtype=torch.uint8).pin_memory()
    p2p = None
    if world > 1 and args.exchange == "nccl":
        keys = torch.empty((2, npix), dtype=torch.int64, device="cuda")
        gathered = torch.empty((2, world, npix), dtype=torch.int64, device="cuda")
    elif world > 1:
        from primestereomatch_b200.sharding import P2PExchange
        p2p = P2PExchange(de, world, rank)
    step_bytes = W * 3 * 4
    banded = world > 1 and args.upload == "banded"
    if banded:  # row band of this rank (pinned host views) + device buffers for the band and the gathered images
        rows = (H + world - 1) // world
        r0, r1 = min(H, rank * rows), min(H, (rank + 1) * rows)
        band_l = torch.zeros((rows, W, 3), dtype=torch.float32, device="cuda")
        band_r = torch.zeros((rows, W, 3), dtype=torch.float32, device="cuda")
        full_l = torch.empty((world * rows, W, 3), dtype=torch.float32, device="cuda")
        full_r = torch.empty((world * rows, W, 3), dtype=torch.float32, device="cuda")

    def step(e2e):
        if e2e == "u8":   # caller keeps 8-bit frames: StereoMatch.cpp:193-197's convertTo runs on the device
            capi.check(L.psm_set_images_u8(de.handle, lp8.data_ptr(), W * 3, rp8.data_ptr(), W * 3), de.handle)
        elif e2e and banded:
            # each rank moves only its band over PCIe; NVLink all-gather completes the images on every GPU
            band_l[: r1 - r0].copy_(lp[r0:r1], non_blocking=True)
            band_r[: r1 - r0].copy_(rp[r0:r1], non_blocking=True)
            dist.all_gather_into_tensor(full_l.view(-1), band_l.view(-1))
            dist.all_gather_into_tensor(full_r.view(-1), band_r.view(-1))
            capi.check(L.psm_set_images_device(de.handle, full_l.data_ptr(), step_bytes, full_r.data_ptr(), step_bytes), de.handle)
        elif e2e:
            capi.check(L.psm_set_images(de.handle, lp.data_ptr(), step_bytes, rp.data_ptr(), step_bytes), de.handle)
        else:
            capi.check(L.psm_set_images_device(de.handle, ld_dev.data_ptr(), step_bytes, rd_dev.data_ptr(), step_bytes), de.handle)
        capi.check(L.psm_cost_const(de.handle), de.handle)
        capi.check(L.psm_cost_filter(de.handle), de.handle)
        if world == 1:
            if e2e:
                capi.check(L.psm_disp_select(de.handle, lmap.data_ptr(), W, rmap.data_ptr(), W), de.handle)
            else:
                capi.check(L.psm_disp_select_device(de.handle), de.handle)
        elif p2p is not None:
            # WTA+scatter kernel, barrier, chunk-reduce+gather kernel, barrier (all over NVLink peer memory)
            p2p.frame(lmap.data_ptr() if e2e else None, rmap.data_ptr() if e2e else None)
        else:
            capi.check(L.psm_disp_select_keys(de.handle, keys[0].data_ptr(), keys[1].data_ptr()), de.handle)
            dist.all_gather_into_tensor(gathered[0].view(-1), keys[0])
            dist.all_gather_into_tensor(gathered[1].view(-1), keys[1])
            out_l = lmap.data_ptr() if e2e else None
            out_r = rmap.data_ptr() if e2e else None
            capi.check(L.psm_disp_reduce_keys(de.handle, gathered[0].data_ptr(), gathered[1].data_ptr(), world,
                                              out_l, W, out_r, W), de.handle)

    def timed(e2e, steps):
        """K steps bracketed by barrier + synchronize on both sides, CUDA events on the launching
        stream, max over ranks.  No host synchronisation inside the region."""
        for _ in range(warm):
            step(e2e)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step(e2e)
        e1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def kernel_times(steps):
        """Duration of the fused CVF kernel alone (cudaEvent pair around that launch, on the launching
        stream) for `steps` further steps of the same workload; reading an event pair needs a host
        sync per step, which is why this is a separate loop from the throughput measurement."""
        out = []
        for _ in range(steps):
            step(False)
            out.append(de.stage_ms(4))
        return out

    launches0 = de.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms = timed(False, args.steps)
    launches_per_step = (de.launch_count() - launches0) // (warm + args.steps)
    kms = kernel_times(args.steps)
    clocks = sampler.stop() if rank == 0 else None
    stage = {n: de.stage_ms(i) for i, n in enumerate(["ingest", "cvc", "cvf", "wta", "cvf_kernel"])}
    e2e_ms = timed(True, args.steps)
    e2e_u8_ms = timed("u8", args.steps)

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = 1e3 / ms_per_step
        e2e_value = 1e3 / (e2e_ms / args.steps)
        peak, peak_src = measured_peaks()
        V = W * H * d_count
        algo_bytes = 2 * (8 * V + 48 * W * H)  # one launch filters both views: read p + write q + 12 guide floats/px
        kern_ms = float(np.mean(kms))
        achieved = algo_bytes / (kern_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "cvf_traffic.json")
        if world == 1 and os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "parallelism": (f"disparity-sharded x{world}, exchange={args.exchange}") if world > 1 else "single GPU",
                       "cvf_mode": ["exact", "mixed", "naive"][args.cvf_mode], "variant": args.variant,
                       "l2": "inputs larger than L2: each step streams 4 x 1.06 GB volumes (raw+filtered, 2 views), no flush needed",
                       "stage_ms_last_step": stage},
            "roofline": {"bound": "hbm", "kernel": "cvf_stream_kernel (fused guided filter, both views per launch)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": kern_ms,
                         "peak_source": peak_src},
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": 2 * W * H * 3 * 4 * (1 if banded else world),
                    "upload": ("banded: each rank uploads H/N rows, NCCL all-gather over NVLink" if banded else "every rank uploads both images"),
                    "d2h_bytes_per_step": 2 * W * H * world, "ms_per_step": e2e_ms / args.steps},
            "e2e_u8": {"value": 1e3 / (e2e_u8_ms / args.steps), "unit": UNIT, "h2d_bytes_per_step": 2 * W * H * 3 * world,
                       "d2h_bytes_per_step": 2 * W * H * world, "ms_per_step": e2e_u8_ms / args.steps,
                       "note": "same as e2e but the host frames are 8-bit (psm_set_images_u8)"},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            threads = min(cores, D)
            sec, tms = oracle_frame(l, r, threads)
            line["cpu_baseline"] = {"value": 1.0 / sec, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "1 full C4 frame (both views) through the CPU oracle, all host cores",
                                    "stage_ms": {"cvc": tms[0], "cvf": tms[1], "wta": tms[2]}}
        print(json.dumps(line), flush=True)
    de.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
