#!/usr/bin/env python
"""Headline benchmark: 1080p ExtractSift frames/s (BASELINE.json metric) on N B200s.

  python bench.py --gpus N --steps K --warmup W            product arm (one rank per GPU)
  python bench.py --impl reference --gpus N ...            the unmodified reference (oracle/_ref)

Workload (BASELINE config 4 / SURVEY.md 8d "C4"): a pool of synthetic 1920x1080 fp32
frames synth(1920,1080,1000+i) resident in HBM, logical frame f uses pool[f % pool],
5 octaves, peakThresh 1.0, edgeThresh 10, maxPts 16384 (~7.6 k keypoints / frame).
A step is one pass of the hot path over `frames_per_step` frames per GPU (512 = C4's
4096 frames / 8 GPUs); frames are sharded cyclically over ranks, no data-path
collective (weak scaling: per-GPU work is fixed).

  value : frames/s, frame already in HBM -> SiftPoint array in pinned HOST memory and numPts
          known on the host: the legacy ExtractSift contract with host=true (main.cpp:317-328,
          cuSIFT.cu:113) and SURVEY.md 8d's timed region
  value_hbm_results : the same with the SiftPoint arrays left in HBM and only the count read back
          (SiftData(dev=true, host=false): the reference downloads only when h_data != NULL,
          cuSIFT.cu:55-58) - what the GPU does when the host link is not in the way
  e2e   : the same through csb_extract_batch with HOST frames: pinned-host upload of
          every frame and result download inside the timed region
          (HEAD SiftData::Extract(float*) contract, cuSIFT.cu:61-120).  8.3 MB up and
          ~4.5 MB down per frame: bound by the box's host<->GPU bandwidth (profiles/r02_pcie_probe.json)
  e2e_u8: extension - frames uploaded as 8-bit and converted on the device
Extra keys (N = 1): per-kernel CUDA-event times and roofline fractions (`kernels`, `k1_pyramid_all`,
`roofline_match`, `k3_orient_desc`), BASELINE configs 1-3 (`c1_demo_pipeline`, `c2` = latency,
`c3_4k_rootsift`), config 5 sample / full (`allpairs_c5_*`).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

W, H = 1920, 1080
N_OCT, PEAK, EDGE, MAXPTS = 5, 1.0, 10.0, 16384
# octave sizes of a 1080p frame: sum = 2 762 040 octave pixels
OCT_PX = [1920 * 1080, 960 * 540, 480 * 270, 240 * 135, 120 * 67]
# algorithmic HBM bytes per octave pixel: pyramid 4 read + 28 DoG written
# (+1 next-octave base), extrema 28 read
BYTES_PYR, BYTES_PYR_LAST, BYTES_EXT = 33.0, 32.0, 28.0


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def tensor_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["bf16_tflops"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured (MEASURED_PEAKS.json)"
    return 1590.0, 1400.0, "fallback (B200_PROFILING.md)"


def tex_peak():
    """Bilinear tex2D<float> fetch peak measured with tools/ubench_tex.cu on this pool's B200 (committed record)."""
    p = ROOT / "profiles" / "r02_tex_peak.json"
    try:
        return float(json.loads(p.read_text())["gfetch_per_s"])
    except (OSError, ValueError, KeyError):
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def reference_arm(args, rank: int, world: int):
    """The unmodified reference through its own API (SiftData::Extract, cuSIFT.cu:61-120) on the same
    config.  The reference is a CUDA library with no CPU implementation, so its arm runs on the GPU of
    the box (rank 0 only); if oracle/_ref is absent the CPU oracle port is timed instead."""
    if rank != 0:
        return
    import cusift_b200 as csb
    from oracle import oracle as O
    frames = max(4, min(args.frames_per_step, 24))      # bounded sample per step
    img = csb.synth(W, H, 1000)
    work = ROOT / "gpurun_out"
    work.mkdir(exist_ok=True)
    line = {"impl": "reference", "metric": "1080p ExtractSift frames/s", "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C4: 1080p ExtractSift, synth(1920,1080,1000), 5 octaves, peakThresh 1.0, "
                                   "edgeThresh 10, maxPts 16384", "frames_per_step": frames}}
    if O.ref_available():
        raw = work / "bench_ref_frame.f32"
        np.ascontiguousarray(img, np.float32).tofile(raw)
        iters = frames * args.steps
        cmd = [O.REF_DRIVER, "bench", raw, W, H, N_OCT, 0.0, PEAK, EDGE, 0.0, MAXPTS, max(3, args.warmup), iters, 0]
        out = subprocess.run([str(x) for x in cmd], capture_output=True, text=True, timeout=1200)
        raw.unlink(missing_ok=True)
        if out.returncode != 0:
            print(json.dumps({"impl": "reference", "unavailable": f"ref_driver failed rc={out.returncode}: "
                                                                  f"{out.stderr[-200:]}"}))
            return
        res = json.loads(out.stdout.strip().splitlines()[-1])
        fps = 1000.0 / res["ms_per_frame"]
        line.update({"value": fps, "ms_per_step": res["ms_per_frame"] * frames,
                     "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": 1, "kind": "reference",
                                      "sample": f"{iters} x SiftData::Extract on one 1080p frame, unmodified reference "
                                                "(CUDA library: 1 host thread + the box's GPU 0)"},
                     "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "keypoints_per_frame": res["keypoints"]})
        # the wall-clock number above is dominated by the reference's per-frame cudaMallocPitch / cudaFree / symbol
        # copies and swings several-fold between boxes; its KERNELS take a stable ~0.5 ms per frame (committed ncu
        # launch list of this same command on a B200)
        try:
            rk = json.loads((ROOT / "profiles" / "r02_reference_kernels.json").read_text())["extract_1080p_frame"]
            line["reference_device_time"] = {"us_per_frame": rk["device_us_per_frame"], "launches_per_frame": rk["launches_per_frame"],
                                             "frames_per_s_if_only_kernels": 1e6 / rk["device_us_per_frame"],
                                             "source": "profiles/r02_reference_kernels.json (ncu launch list, not measured in this run)"}
        except (OSError, ValueError, KeyError):
            pass
    else:
        t0 = time.perf_counter()
        n = 0
        for _ in range(2):
            O.extract(img, N_OCT, 0.0, PEAK, EDGE, 0.0, False, MAXPTS)
            n += 1
        fps = n / (time.perf_counter() - t0)
        line.update({"value": fps, "ms_per_step": 1000.0 * frames / fps,
                     "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                                      "sample": f"{n} frames, oracle/oracle.c (OpenMP)"},
                     "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line))


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the committed ncu --set full
    capture (profiles/*_traffic.json, written by tools/summarize_ncu.py); None if there is no capture."""
    best = None
    for f in sorted((ROOT / "profiles").glob("*_traffic.json")):
        try:
            d = json.loads(f.read_text())
        except (OSError, ValueError):
            continue
        if kernel in d:
            best = d[kernel].get("dram_bytes")
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=512)
    ap.add_argument("--pool", type=int, default=32)
    ap.add_argument("--slots", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="skip the extra N=1 arms (matcher roofline, configs 1 and 3)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import cusift_b200 as csb

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = csb.Context(local_rank, args.slots)
    prm = csb.make_params(N_OCT, 0.0, PEAK, EDGE, 0.0)
    F = args.frames_per_step
    # logical frames of this rank: global frame g = rank + world*i (cyclic shard), pool image g % pool
    my_frames = csb.shard_frames(F * world, rank, world)
    pool_ids = sorted({g % args.pool for g in my_frames})
    pool_host = {i: csb.synth(W, H, 1000 + i) for i in pool_ids}
    pool_dev = {}
    pitch = csb.align_up(W, 128)
    for i, im in pool_host.items():
        pool_dev[i], pitch = ctx.upload_image(im)
    # pinned host copies of the pool for the e2e arm
    pins_img = {}
    for i, im in pool_host.items():
        pa = csb.PinnedArray(W * H, np.float32)
        pa.array[:] = im.ravel()
        pins_img[i] = pa
    nbuf = max(2 * args.slots, 8)
    d_sifts = [ctx.alloc(588 * MAXPTS) for _ in range(nbuf)]
    pins = [csb.PinnedArray(MAXPTS) for _ in range(nbuf)]

    dev_list = [pool_dev[g % args.pool] for g in my_frames]
    host_list = [pins_img[g % args.pool].ptr for g in my_frames]
    ds_list = [d_sifts[k % nbuf] for k in range(len(my_frames))]
    hs_list = [pins[k % nbuf].ptr for k in range(len(my_frames))]

    def step_device():            # frames in HBM -> SiftPoint arrays in HBM, counts on the host
        return ctx.extract_batch(dev_list, W, H, pitch, prm, ds_list, None, MAXPTS)

    def step_device_host_results():   # same, SiftPoint arrays downloaded into pinned host memory
        return ctx.extract_batch(dev_list, W, H, pitch, prm, ds_list, hs_list, MAXPTS)

    def step_host():
        return ctx.extract_batch(host_list, W, H, W, prm, ds_list, hs_list, MAXPTS, on_host=True)

    # 8-bit ingest arm: the same frames rounded to uint8, uploaded as bytes, converted on the device
    pins_u8 = {}
    for i, im in pool_host.items():
        pa = csb.PinnedArray(W * H, np.uint8)
        pa.array[:] = np.clip(np.rint(im), 0, 255).astype(np.uint8).ravel()
        pins_u8[i] = pa
    u8_list = [pins_u8[g % args.pool].ptr for g in my_frames]

    def step_host_u8():
        return ctx.extract_batch_u8(u8_list, W, H, W, False, prm, ds_list, hs_list, MAXPTS)

    # opt-in compact result records (288 B instead of 588 B per keypoint: fp16 descriptor), csb_extract_batch_compact
    pins_c = [csb.PinnedArray(MAXPTS, csb.COMPACT_DTYPE) for _ in range(nbuf)]
    hc_list = [pins_c[k % nbuf].ptr for k in range(len(my_frames))]

    def step_device_compact():
        return ctx.extract_batch_compact(dev_list, W, H, pitch, prm, ds_list, hc_list, MAXPTS, source="device")

    def step_host_u8_compact():
        return ctx.extract_batch_compact(u8_list, W, H, W, prm, ds_list, hc_list, MAXPTS, source="host_u8")

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize; device-clock time via CUDA events recorded on an idle
        stream right after each synchronisation; max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launch_count()
        e0.record()
        t0 = time.perf_counter()
        counts = None
        for _ in range(steps):
            counts = fn()
        torch.cuda.synchronize()
        e1.record()
        e1.synchronize()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        launches = ctx.launch_count() - l0
        t = torch.tensor([ms, wall * 1e3], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return float(t[0]), float(t[1]), launches, counts

    for _ in range(args.warmup):
        step_device_host_results()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, wall_ms, launches, counts = timed(step_device_host_results, args.steps)
    clocks = sampler.stop()
    frames_total = F * world * args.steps
    value = frames_total / (ms / 1e3)

    for _ in range(args.warmup):
        step_device()
    ms_h, _, _, _ = timed(step_device, args.steps)
    value_hbm_results = frames_total / (ms_h / 1e3)

    for _ in range(args.warmup):
        step_host()
    ms_e, wall_e, launches_e, counts_e = timed(step_host, args.steps)
    e2e_value = frames_total / (ms_e / 1e3)
    for _ in range(args.warmup):
        step_host_u8()
    ms_u8, _, _, counts_u8 = timed(step_host_u8, args.steps)
    e2e_u8 = {"value": frames_total / (ms_u8 / 1e3), "unit": "frames/s", "h2d_bytes_per_step": len(my_frames) * W * H * world,
              "d2h_bytes_per_step": (int(np.sum(counts_u8)) * 588 + len(my_frames) * 4) * world, "ms_per_step": ms_u8 / args.steps,
              "note": "csb_extract_batch_u8: frames uploaded as 8-bit, converted to fp32 on the device (extension, SURVEY 8f-4)"}
    for _ in range(args.warmup):
        step_device_compact()
    ms_vc, _, _, counts_vc = timed(step_device_compact, args.steps)
    for _ in range(args.warmup):
        step_host_u8_compact()
    ms_uc, _, _, counts_uc = timed(step_host_u8_compact, args.steps)
    compact = {"record_bytes": 288, "note": "opt-in csb_extract_batch_compact: header + fp16 descriptor instead of the 588-byte "
                                            "SiftPoint; the default entry points keep the reference layout",
               "value_compact_results": frames_total / (ms_vc / 1e3),
               "e2e_u8_compact": {"value": frames_total / (ms_uc / 1e3), "unit": "frames/s",
                                  "h2d_bytes_per_step": len(my_frames) * W * H * world,
                                  "d2h_bytes_per_step": (int(np.sum(counts_uc)) * 288 + len(my_frames) * 4) * world,
                                  "ms_per_step": ms_uc / args.steps}}
    kp_sum = int(np.sum(counts_e))
    h2d = len(my_frames) * W * H * 4
    d2h = kp_sum * 588 + len(my_frames) * 8

    # single-frame latency (BASELINE config 2): median of 200 synchronous csb_extract calls
    lat = []
    for i in range(220):
        t0 = time.perf_counter()
        ctx.extract_batch(dev_list[:1], W, H, pitch, prm, ds_list[:1], hs_list[:1], MAXPTS)
        if i >= 20:
            lat.append((time.perf_counter() - t0) * 1e3)

    # per-kernel device time, CUDA events on each launching stream, same workload (profiled pass)
    ctx.profile(True)
    ctx.profile_reset()
    nprof = min(len(my_frames), 64)
    for k in range(nprof):     # one frame in flight: kernels are timed alone, not overlapped with other slots
        ctx.extract_batch(dev_list[k:k + 1], W, H, pitch, prm, ds_list[k:k + 1], hs_list[k:k + 1], MAXPTS)
    tab = ctx.profile_table()
    ctx.profile(False)
    peak, peak_src = peaks()
    kernels = {}
    # algorithmic HBM bytes per launch (4 B read + 28 B DoG written (+ 1 B next base) per octave pixel for
    # the pyramid, 28 B read per octave pixel for the extrema scan)
    alg_bytes = {"pyramid_o0": OCT_PX[0] * BYTES_PYR,                      # octave 0 incl. the octave-1 base it writes
                 "pyramid_rest": sum(OCT_PX[1:]) * BYTES_PYR_LAST,         # octaves 1..4: base read + 7 DoG planes
                 "down_chain": OCT_PX[1] * 4.0 + sum(OCT_PX[2:]) * 4.0,    # reads base 1, writes bases 2..4
                 "find_points": sum(OCT_PX) * BYTES_EXT}
    for name, v in tab.items():
        if not v["launches"]:
            continue
        avg_ms = v["total_ms"] / v["launches"]
        ent = {"avg_us": round(avg_ms * 1e3, 2), "launches_per_frame": v["launches"] / nprof,
               "us_per_frame": round(v["total_ms"] * 1e3 / nprof, 2)}
        alg = alg_bytes.get(name)
        if alg:
            ent["alg_bytes"] = alg
            ent["achieved_gbs"] = round(alg / (avg_ms * 1e-3) / 1e9, 1)
            ent["frac"] = round(ent["achieved_gbs"] / peak, 4)
        kernels[name] = ent
    hbm_kernels = {k: v for k, v in kernels.items() if k in ("pyramid_o0", "find_points")}
    dom = max(hbm_kernels, key=lambda k: hbm_kernels[k]["us_per_frame"]) if hbm_kernels else None
    roofline = None
    if dom:
        d = hbm_kernels[dom]
        roofline = {"kernel": dom, "bound": "hbm", "achieved": d["achieved_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": d["frac"], "traffic": ncu_traffic(dom), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": d["alg_bytes"], "avg_launch_us": d["avg_us"],
                    "timing": "CUDA events around every launch on its own stream, one frame in flight, the stream held by a "
                              "150 us spin kernel while the frame's launches are queued (no host enqueue gaps inside the brackets)"}
    # the whole pyramid (K1 = 91.1 MB per 1080p frame) = three launches
    k1_names = [k for k in ("pyramid_o0", "down_chain", "pyramid_rest") if k in kernels]
    k1_all = None
    if k1_names:
        us = sum(kernels[k]["us_per_frame"] for k in k1_names)
        alg = OCT_PX[0] * BYTES_PYR + sum(OCT_PX[1:-1]) * BYTES_PYR + OCT_PX[-1] * BYTES_PYR_LAST
        k1_all = {"launches": k1_names, "us_per_frame": round(us, 2), "alg_bytes": alg,
                  "achieved_gbs": round(alg / (us * 1e-6) / 1e9, 1), "frac": round(alg / (us * 1e-6) / 1e9 / peak, 4)}
    # orientation + descriptor kernel: keypoints/s and fraction of the measured bilinear-fetch peak
    k3 = None
    if "orient_desc" in kernels:
        kp = float(np.mean(counts))
        us = kernels["orient_desc"]["us_per_frame"]
        fetches = 1508.0 * kp                      # 121 x 4 (orientation) + 256 x 4 (descriptor) bilinear fetches per keypoint
        tp = tex_peak()
        k3 = {"us_per_frame": us, "keypoints_per_s": round(kp / (us * 1e-6), 0), "gfetch_per_s": round(fetches / (us * 1e-6) / 1e9, 1),
              "tex_peak_gfetch_per_s": tp, "tex_frac": round(fetches / (us * 1e-6) / 1e9 / tp, 4) if tp else None,
              "bytes_written_per_keypoint": 588}

    # BASELINE config 5: all-pairs MatchSiftData + 1024-hypothesis RANSAC + ImproveHomography over sets of 8192 keypoints
    # through csb_allpairs_distributed (NCCL all-gathers issued from the C++ library, pairs partitioned cyclically,
    # results all-gathered).  At N = 8 this is the FULL configuration (256 sets, 32 640 pairs); smaller N run 8 sets per GPU.
    allpairs = None
    try:
        P, per = 8192, (32 if world == 8 else 8)
        prm5 = csb.make_params(N_OCT, 0.0, 0.5, EDGE, 0.0)
        d_sets, c_sets = [], []
        for k in range(per):
            pts = ctx.extract(csb.synth(W, H, 3000 + rank * per + k), prm5, max_pts=32768)
            order = np.lexsort((pts["scale"], pts["coords2D"][:, 1], pts["coords2D"][:, 0], pts["subsampling"]))
            pts = np.ascontiguousarray(pts[order][:P])
            d_sets.append(ctx.upload_sift(pts))
            c_sets.append(len(pts))

        def bcast(raw):
            if dist is None:
                return raw
            obj = [raw]
            dist.broadcast_object_list(obj, src=0)
            return obj[0]

        comm = ctx.nccl_comm(rank, world, bcast) if dist is not None else None
        kw5 = dict(distance="l2", num_loops=1024, min_score=0.0, max_ambiguity=0.80, thresh=5.0, seed=1, improve_loops=5,
                   improve_thresh=3.0)
        ctx.allpairs_distributed(comm, rank, world, d_sets, c_sets, P, **kw5)          # warm-up: allocations, NCCL channels
        best5, tm5, res5 = 1e30, None, None
        for _ in range(2):
            barrier()
            t0 = time.perf_counter()
            res5 = ctx.allpairs_distributed(comm, rank, world, d_sets, c_sets, P, **kw5)
            tt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
            if dist is not None:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            if float(tt[0]) < best5:
                best5, tm5 = float(tt[0]), res5["timings_ms"].copy()
        cs = torch.tensor(c_sets, device="cuda", dtype=torch.int64)
        if dist is not None:
            allc = [torch.zeros_like(cs) for _ in range(world)]
            dist.all_gather(allc, cs)
            cnts = torch.cat(allc).cpu().numpy()
        else:
            cnts = cs.cpu().numpy()
        n_sets = world * per
        prs = csb.all_pairs(n_sets)
        q = float(sum(int(cnts[i]) for i, _ in prs))
        qc = float(sum(int(cnts[i]) * int(cnts[j]) for i, j in prs))
        allpairs = {"sets": n_sets, "points_per_set": P, "pairs": len(prs), "seconds": best5, "Mmatches_per_s": q / best5 / 1e6,
                    "pairs_per_s": len(prs) / best5, "useful_TFLOP_per_s": 2 * 128 * qc / best5 / 1e12, "ransac_loops": 1024,
                    "improve_loops": 5, "full_config_5": bool(n_sets == 256),
                    "exchange": "ncclAllGather of the SiftPoint arrays issued from C++ (csb_allpairs_distributed)" if dist is not None
                    else "none (1 GPU)",
                    "exchange_bytes_per_rank": per * P * 588,
                    "rank0_timings_ms": {"local_pairs_and_queueing": float(tm5[0]), "wait_for_exchange": float(tm5[1]),
                                         "remaining_pairs": float(tm5[2]), "result_exchange": float(tm5[3])},
                    "mean_inliers": float(res5["inliers"].mean()), "mean_numfit": float(res5["num_fit"].mean())}
        if comm is not None:
            ctx.nccl_comm_destroy(comm)
        for d_ in d_sets:
            ctx.free(d_)
    except Exception as e:  # noqa: BLE001  (informational arm: never fail the headline line)
        allpairs = {"unavailable": repr(e)}

    # ---- extra arms at N = 1: matcher tensor-pipe roofline, BASELINE configs 1 and 3 ---------------------------------
    roofline_match = c1 = c3 = None
    if rank == 0 and world == 1 and not args.quick:
        L = csb.lib()

        def rand_set(n, seed):
            r = np.random.default_rng(seed)
            sset = np.zeros(n, csb.SIFT_DTYPE)
            d = np.abs(r.standard_normal((n, 128))).astype(np.float32)
            sset["data"] = d / np.linalg.norm(d, axis=1, keepdims=True)
            sset["coords2D"] = r.uniform(0, 1000, (n, 2)).astype(np.float32)
            return sset

        try:   # 8192 x 8192 MatchSiftData (BASELINE config 5's pair size), kernels timed with CUDA events
            n = 8192
            sa, sb = rand_set(n, 1), rand_set(n, 2)
            d1, d2 = ctx.upload_sift(sa), ctx.upload_sift(sb)
            for _ in range(3):
                L.csb_match(ctx.h, d1, n, d2, n, 1, None)
            reps = 20
            t0 = time.perf_counter()
            for _ in range(reps):
                L.csb_match(ctx.h, d1, n, d2, n, 1, None)
            wall_us = (time.perf_counter() - t0) / reps * 1e6
            ctx.profile(True)
            ctx.profile_reset()
            for _ in range(reps):
                L.csb_match(ctx.h, d1, n, d2, n, 1, None)
            mt = ctx.profile_table()
            ctx.profile(False)
            us = {k: mt[k]["total_ms"] * 1e3 / reps for k in ("match_pack", "match_tc", "match_rescore", "match_redo") if k in mt}
            flop = 2.0 * 128 * n * n
            burst, sustained, tsrc = tensor_peaks()
            scan = us.get("match_tc", float("nan"))
            scored = scan + us.get("match_rescore", 0.0) + us.get("match_redo", 0.0)
            roofline_match = {"bound": "tensor", "n1": n, "n2": n, "flop": flop, "us": {k: round(v, 2) for k, v in us.items()},
                              "achieved": round(flop / (scan * 1e-6) / 1e12, 1), "peak": burst, "unit": "TFLOP/s",
                              "frac": round(flop / (scan * 1e-6) / 1e12 / burst, 4),
                              "frac_incl_rescoring": round(flop / (scored * 1e-6) / 1e12 / burst, 4),
                              "peak_source": tsrc + " bf16_tflops (burst; the scan is timed alone)",
                              "input_type": "fp16 operands, fp32 accumulation in TMEM, exact fp32 rescoring of the short lists",
                              "call_wall_us": round(wall_us, 1), "Mmatches_per_s_one_pair_at_a_time": round(n / wall_us, 2)}
            # what actually bounds the scan: every score is read once from tensor memory, whose read port delivers
            # 64 B/clk per SM (measured); the scan makes 1.25 passes over the n x n fp32 scores
            tmem_floor_us = 1.25 * n * n * 4.0 / (148 * 64.0 * 1.965e9) * 1e6
            roofline_match["tmem_read"] = {"passes": 1.25, "bytes": 1.25 * n * n * 4.0, "port_bytes_per_clk_per_sm": 64,
                                           "floor_us": round(tmem_floor_us, 2), "frac": round(tmem_floor_us / scan, 4)}
            from oracle import oracle as O
            if O.ref_available():
                work = ROOT / "gpurun_out"
                work.mkdir(exist_ok=True)
                fa, fb = work / "bench_a.sift", work / "bench_b.sift"
                O.write_sift_file(fa, sa)
                O.write_sift_file(fb, sb)
                out = subprocess.run([str(O.REF_DRIVER), "benchmatch", str(fa), str(fb), "2", "5"], capture_output=True,
                                     text=True, timeout=300)
                fa.unlink(missing_ok=True)
                fb.unlink(missing_ok=True)
                if out.returncode == 0:
                    roofline_match["reference_ms_per_pair"] = json.loads(out.stdout.strip().splitlines()[-1])["ms_per_pair"]
            ctx.free(d1)
            ctx.free(d2)
        except Exception as e:  # noqa: BLE001
            roofline_match = {"unavailable": repr(e)}

        try:   # BASELINE config 1: main.cpp's demo on the reference's own image pair (tests/golden/frames.npz)
            import cv2
            z = np.load(ROOT / "tests" / "golden" / "frames.npz")
            g = [cv2.GaussianBlur(z[k].astype(np.float32), (3, 3), 0.5) for k in ("gray1", "gray2")]   # main.cpp:308-309
            p1 = csb.make_params(6, 0.0, 0.1, 10.0, 0.0)
            dimg = [ctx.upload_image(x) for x in g]
            dsf = [ctx.alloc(588 * 4096) for _ in range(2)]
            pin = [csb.PinnedArray(4096) for _ in range(2)]
            H9 = np.zeros(9, np.float32)
            nin, nfit = C.c_int(0), C.c_int(0)
            rngs = np.random.default_rng(1)
            stages = {"extract_x2": [], "match": [], "find_homography_10000": [], "improve_homography_5": []}
            for it in range(12):
                t0 = time.perf_counter()
                cnt = [int(ctx.extract_batch([dimg[k][0]], 640, 480, dimg[k][1], p1, [dsf[k]], [pin[k].ptr], 4096)[0]) for k in range(2)]
                t1 = time.perf_counter()
                L.csb_match(ctx.h, dsf[0], cnt[0], dsf[1], cnt[1], 1, pin[0].ptr)
                t2 = time.perf_counter()
                mm = pin[0].array[: cnt[0]]
                valid = np.nonzero((mm["score"] > 0.0) & (mm["ambiguity"] < 0.80))[0].astype(np.int32)
                rp = np.ascontiguousarray(valid[rngs.integers(0, len(valid), (4, 10000))], np.int32)
                t3 = time.perf_counter()
                ctx._check(L.csb_find_homography(ctx.h, dsf[0], cnt[0], rp.ctypes.data_as(C.POINTER(C.c_int)), 10000, 5.0,
                                                 H9.ctypes.data_as(C.POINTER(C.c_float)), C.byref(nin)), "csb_find_homography")
                t4 = time.perf_counter()
                ctx._check(L.csb_improve_homography(ctx.h, dsf[0], cnt[0], H9.ctypes.data_as(C.POINTER(C.c_float)), 5, 0.0, 0.80, 3.0,
                                                    C.byref(nfit), None), "csb_improve_homography")
                t5 = time.perf_counter()
                if it >= 2:
                    for k, v in zip(stages, (t1 - t0, t2 - t1, t4 - t3, t5 - t4)):
                        stages[k].append(v * 1e3)
            c1 = {"workload": "C1: main.cpp demo on test/data/color1+2 (640x480, 3x3 pre-blur, 6 octaves, thresh 0.1, maxPts 4096), "
                              "MatchSiftData L2, FindHomography(10000, 0.0, 0.80, 5.0), ImproveHomography(5, 0.0, 0.80, 3.0)",
                  "ms_per_stage_median": {k: round(statistics.median(v), 4) for k, v in stages.items()},
                  "ms_total_median": round(sum(statistics.median(v) for v in stages.values()), 4),
                  "num_pts": cnt, "num_matches": int(nin.value), "num_fit": int(nfit.value)}
            for d_, _ in dimg:
                ctx.free(d_)
            for d_ in dsf:
                ctx.free(d_)
        except Exception as e:  # noqa: BLE001
            c1 = {"unavailable": repr(e)}

        try:   # BASELINE config 3: one 3840x2160 frame, ExtractRootSift, thresh 0.1 (~95 k keypoints)
            img4k = csb.synth(3840, 2160, 2000)
            p3 = csb.make_params(5, 0.0, 0.1, 10.0, 0.0, rootsift=True)
            d4k, pitch4k = ctx.upload_image(img4k)
            ds4k = ctx.alloc(588 * 131072)
            pin4k = csb.PinnedArray(131072)
            ts = []
            for it in range(13):
                t0 = time.perf_counter()
                c4k = ctx.extract_batch([d4k], 3840, 2160, pitch4k, p3, [ds4k], [pin4k.ptr], 131072)
                if it >= 3:
                    ts.append((time.perf_counter() - t0) * 1e3)
            ms3 = statistics.median(ts)
            ctx.profile(True)
            ctx.profile_reset()
            for _ in range(4):
                ctx.extract_batch([d4k], 3840, 2160, pitch4k, p3, [ds4k], [pin4k.ptr], 131072)
            t3 = ctx.profile_table()
            ctx.profile(False)
            c3 = {"workload": "C3: synth(3840,2160,2000), ExtractRootSift, 5 octaves, thresh 0.1, maxPts 131072; device frame -> "
                              "SiftPoint array in pinned host memory", "ms_per_frame_median": round(ms3, 3),
                  "keypoints": int(c4k[0]), "keypoints_per_s": round(int(c4k[0]) / (ms3 * 1e-3), 0),
                  "kernel_us": {k: round(v["total_ms"] * 1e3 / 4, 1) for k, v in t3.items() if v["launches"]}}
            ctx.free(d4k)
            ctx.free(ds4k)
            pin4k.free()
        except Exception as e:  # noqa: BLE001
            c3 = {"unavailable": repr(e)}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        img = pool_host[pool_ids[0]]
        t0 = time.perf_counter()
        O.extract(img, N_OCT, 0.0, PEAK, EDGE, 0.0, False, MAXPTS)
        one = time.perf_counter() - t0
        n = int(max(2, min(12, 15.0 / max(one, 1e-3))))
        t0 = time.perf_counter()
        for k in range(n):
            O.extract(pool_host[pool_ids[k % len(pool_ids)]], N_OCT, 0.0, PEAK, EDGE, 0.0, False, MAXPTS)
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": n / dt, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                        "sample": f"{n} frames of the pool through oracle/oracle.c (OpenMP, all host cores)"}
        try:
            import cv2
            cv2.setNumThreads(os.cpu_count())
            sift = cv2.SIFT_create(nOctaveLayers=3, contrastThreshold=0.04, edgeThreshold=10, sigma=1.6)
            u8 = np.clip(img, 0, 255).astype(np.uint8)
            t0 = time.perf_counter()
            kps, _ = sift.detectAndCompute(u8, None)
            cpu_baseline["opencv_sift"] = {"frames_per_s": 1.0 / (time.perf_counter() - t0), "keypoints": len(kps),
                                           "threads": os.cpu_count()}
        except Exception as e:  # noqa: BLE001
            cpu_baseline["opencv_sift"] = {"unavailable": repr(e)}

    if rank == 0:
        line = {
            "metric": "1080p ExtractSift frames/s", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C4: 1080p ExtractSift throughput, frames sharded cyclically over GPUs; pool of "
                                   f"{args.pool} frames synth(1920,1080,1000+i) resident in HBM; 5 octaves, peakThresh 1.0, "
                                   "edgeThresh 10, maxPts 16384",
                       "frames_per_step_per_gpu": F, "frames_per_step": F * world, "slots_per_gpu": args.slots,
                       "l2": f"inputs larger than L2: {len(pool_ids)} x 8.3 MB frame pool + {args.slots} x 90 MB pyramid "
                             "workspaces cycle through HBM between reuses",
                       "timed_region": "device frame -> SiftPoint array in pinned host memory + count on the host (legacy "
                                       "ExtractSift contract, SURVEY.md 8d); value_hbm_results leaves the arrays in HBM "
                                       "(SiftData(dev=true, host=false)); e2e adds the upload of every frame as well"},
            "value_hbm_results": value_hbm_results,
            "ms_per_frame": ms / args.steps / (F * world), "wall_ms_per_step": wall_ms / args.steps,
            "latency_ms_single_frame": {"median": statistics.median(lat), "p99": sorted(lat)[int(len(lat) * 0.99) - 1]},
            "keypoints_per_frame": float(np.mean(counts)),
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world, "ms_per_step": ms_e / args.steps},
            "e2e_u8": e2e_u8, "compact_results": compact,
            "gpu_launches": int(launches) * world, "gpu_launches_e2e": int(launches_e) * world,
            "allpairs_c5": allpairs,
            "clocks": clocks, "roofline": roofline, "kernels": kernels, "k1_pyramid_all": k1_all, "k3_orient_desc": k3,
            "roofline_match": roofline_match, "c1_demo_pipeline": c1, "c3_4k_rootsift": c3, "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line))
    barrier()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
