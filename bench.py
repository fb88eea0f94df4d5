#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's config.

  python bench.py --gpus N --steps K --warmup W            # the CUDA path (this repo)
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

A "step" is one full pass of the hot path over one batch of synthetic input: one `Camera::render` of the
seeded book1 end scene at 1920x1080, 100 spp, max depth 50 (207.36 M camera samples).  For N > 1 the same
image is sharded by interleaved row blocks over the ranks (strong scaling; the only exchange is the NCCL
framebuffer gather).  Timing: CUDA events around every step on the launching stream, L2 flushed between
steps, max over ranks.
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

WORKLOAD = dict(scene="book1", image_width=1920, samples=100, seed=1)
WORKLOAD_NAME = "book1 end scene 1920x1080, 100 spp, max depth 50 (BASELINE configs[0]; the config the metric is quoted on)"


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t_begin=None, t_end=None):
        """Median / reasons over the samples taken in [t_begin, t_end] (the sampler is started before the warm-up so that
        nvidia-smi's own start-up does not eat a short timed region)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for t, r in self.rows if (t_begin is None or t >= t_begin) and (t_end is None or t <= t_end)]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_sample(desc, cam, row_step, threads=0):
    """The CPU path (oracle port of the reference algorithm) on a bounded sample: every row_step-th row of
    the same image at full spp.  Returns (stats dict, per-segment traversal counters)."""
    from oracle import binding as oracle

    orc = oracle.OracleScene(desc)
    threads = threads or (os.cpu_count() or 1)
    _, _, st = orc.render(cam, seed=WORKLOAD["seed"], rows=(0, cam.image_height, row_step), threads=threads, want_rgb8=False)
    return st


def hbm_view(alg_bytes_per_launch, dur_s, traffic):
    peak, src = 6533.5, "fallback"
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        for k in ("hbm_gbs", "hbm_gbps", "hbm_copy_gbps"):
            if k in mp:
                peak, src = float(mp[k]), f"MEASURED_PEAKS.json:{k}"
                break
    except Exception:
        pass
    ach = alg_bytes_per_launch / dur_s / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": src,
            "dram_gbps_measured": (traffic / dur_s / 1e9) if traffic else None}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path, timed on the host cores.
    The reference is a Rust crate and no rustc/cargo exists here, so the timed code is the oracle port
    (kind "port"): same algorithm and arithmetic, world shared read-only, all host threads."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from crucible_b200 import demo_builder

    sc = demo_builder.book1_end_scene(image_width=WORKLOAD["image_width"], samples=WORKLOAD["samples"], seed=WORKLOAD["seed"])
    desc, cam = sc.describe(), sc.scene_cam.to_abi()
    threads = os.cpu_count() or 1
    # size the sample so that one step is roughly 5-10 s of wall time on this box
    t0 = time.time()
    st = oracle_sample(desc, cam, 120, threads)
    rate = st["samples"] / max(time.time() - t0, 1e-6)
    row_step = int(max(1, min(120, round(cam.image_height * cam.image_width * cam.samples / max(rate * 6.0, 1.0)))))
    for _ in range(args.warmup):
        oracle_sample(desc, cam, max(row_step, 60), threads)
    secs, samples, rays = 0.0, 0, 0
    for _ in range(args.steps):
        st = oracle_sample(desc, cam, row_step, threads)
        secs += st["seconds"]
        samples += st["samples"]
        rays += st["rays"]
    val = samples / secs / 1e6
    sample_desc = f"rows j % {row_step} == 0 of the 1920x1080x100spp job ({samples // args.steps} samples per step)"
    line = {"impl": "reference", "metric": "Msamples/s", "value": val, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": WORKLOAD_NAME, "sample": sample_desc},
            "mrays_per_s": rays / secs / 1e6,
            "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": threads, "kind": "port", "sample": sample_desc},
            "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="crucible_b200")
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--pool", type=int, default=0)
    ap.add_argument("--samples", type=int, default=WORKLOAD["samples"], help="debug only: a reduced-spp run is NOT the headline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from crucible_b200 import abi, demo_builder, multigpu
    from crucible_b200.gpu import GpuScene, rows_of_rank

    rank, world, local = dist_env()
    assert torch.cuda.is_available(), "bench.py needs a B200: crucible_b200 has no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    precision = abi.CR_PRECISION_F64 if args.precision == "f64" else abi.CR_PRECISION_F32

    sc = demo_builder.book1_end_scene(image_width=WORKLOAD["image_width"], samples=args.samples, seed=WORKLOAD["seed"])
    desc, cam = sc.describe(), sc.scene_cam.to_abi()
    H, W = cam.image_height, cam.image_width
    gs = GpuScene(desc, local)
    row_block = 8
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step(time_kernels):
        full, full8, st = multigpu.render_sharded(gs, cam, rank, world, seed=WORKLOAD["seed"], precision=precision,
                                                  row_block=row_block, pool_paths=args.pool, time_kernels=time_kernels)
        return st, full8

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step(False)
        flush.fill_(1)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    stats = []
    barrier()
    t_timed0 = time.perf_counter()
    for k in range(args.steps):
        ev[k][0].record()
        st, _ = step(True)
        ev[k][1].record()
        stats.append(st)
        flush.fill_(k)  # L2 flush between timed iterations, outside the events
    barrier()
    clocks = sampler.stop(t_timed0, time.perf_counter()) if rank == 0 else None
    ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    agg = torch.tensor([sum(s["samples"] for s in stats), sum(s["rays"] for s in stats), sum(s["launches"] for s in stats),
                        sum(s["iterations"] for s in stats)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    ms_total = float(t.item())
    samples, rays, launches, iters = (float(x) for x in agg.tolist())

    # ---- e2e: the public API with HOST buffers.  Every step: scene description -> cr_scene_* (host BVH build
    # as Scene::render_image does per frame, scene/mod.rs:333) -> H2D -> render -> gather -> D2H of the image.
    h2d = sum(b[1].nbytes + b[2].nbytes + b[3].nbytes for b in desc.batches) + abi.C.sizeof(abi.CrCamera)
    d2h = H * W * 3 * (8 + 1)
    pinned = torch.empty((H, W, 3), dtype=torch.float64).pin_memory()
    pinned8 = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()

    pinned_np, pinned8_np = pinned.numpy(), pinned8.numpy()

    def e2e_step():
        g = GpuScene(desc, local)
        if world == 1:
            # the reference-facing call itself: cr_render with HOST buffers (H2D of the camera, D2H of both images inside)
            g.render(cam, seed=WORKLOAD["seed"], precision=precision, pool_paths=args.pool, out_rgb=pinned_np, out_rgb8=pinned8_np)
        else:
            full, full8, st = multigpu.render_sharded(g, cam, rank, world, seed=WORKLOAD["seed"], precision=precision,
                                                      row_block=row_block, pool_paths=args.pool)
            if rank == 0:
                pinned.copy_(full, non_blocking=True)
                pinned8.copy_(full8, non_blocking=True)
            torch.cuda.synchronize()
        g.close()

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())

    if rank == 0:
        total_samples_per_step = H * W * cam.samples
        value = samples / (ms_total * 1e-3) / 1e6
        line = {"metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
                "config": {"workload": WORKLOAD_NAME, "parallelism": f"rows x{world} (blocks of {row_block})", "l2": "flushed between steps (256 MiB write)",
                           "seed": WORKLOAD["seed"], "samples_per_step": total_samples_per_step, "spp": cam.samples},
                "mrays_per_s": rays / (ms_total * 1e-3) / 1e6,
                "rays_per_sample": rays / samples,
                "clocks": clocks,
                "e2e": {"value": total_samples_per_step * args.steps / e2e_s / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h},
                "gpu_launches": int(launches)}
        # ---- roofline of the dominant kernel (k_trace): algorithmic flops per ray segment from the oracle's
        # reference-order traversal counters, segments per launch from the run, duration from events
        ms_trace = sum(s["ms_trace"] for s in stats)
        ms_shade = sum(s["ms_shade"] for s in stats)
        ms_gen = sum(s["ms_raygen"] for s in stats)
        my_iters = sum(s["iterations"] for s in stats)
        my_rays = sum(s["rays"] for s in stats)
        f64p, f32p = abi.C.c_double(), abi.C.c_double()
        abi.check(abi.load().cr_measure_fma_peak(local, abi.C.byref(f64p), abi.C.byref(f32p)))
        cpu = None
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            probe = oracle_sample(desc, cam, 120, threads)  # 9 rows: sizes the sample for ~15 s of CPU work
            rate = probe["samples"] / max(probe["seconds"], 1e-6)
            row_step = int(max(1, min(120, round(total_samples_per_step / max(rate * 15.0, 1.0)))))
            ost = oracle_sample(desc, cam, row_step, threads)
            cpu = ost
            line["cpu_baseline"] = {"value": ost["samples"] / ost["seconds"] / 1e6, "unit": "Msamples/s", "cores": threads, "kind": "port",
                                    "sample": f"rows j % {row_step} == 0 of the same 1920x1080x{cam.samples}spp job ({ost['samples']} samples, {ost['seconds']:.1f} s)",
                                    "mrays_per_s": ost["rays"] / ost["seconds"] / 1e6}
        else:
            cpu = oracle_sample(desc, cam, 120, os.cpu_count() or 1)
        per_seg_flops = (24.0 * cpu["node_tests"] + 40.0 * cpu["sphere_tests"] + 50.0 * cpu["tri_tests"]) / cpu["rays"] + 60.0
        per_seg_bytes = (32.0 * cpu["node_tests"] + 16.0 * cpu["sphere_tests"] + 36.0 * cpu["tri_tests"]) / cpu["rays"] + 16.0
        seg_per_launch = my_rays / max(my_iters, 1)
        dur_s = ms_trace * 1e-3 / max(my_iters, 1)
        achieved = per_seg_flops * seg_per_launch / dur_s / 1e12
        peak = f64p.value if args.precision == "f64" else f32p.value
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "trace_r01d_traffic.json")
        if args.precision == "f64" and os.path.exists(tpath):
            tj = json.load(open(tpath))  # DRAM bytes of one ncu --set full capture, scaled to this run's segments per launch
            traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) / tj["rays_in_launch"] * seg_per_launch
            traffic_src = tj["capture"]
        line["roofline"] = {"bound": "fp64" if args.precision == "f64" else "fp32", "kernel": "k_trace", "achieved": achieved, "peak": peak,
                            "unit": "TFLOP/s", "frac": achieved / peak if peak else None, "traffic": traffic, "traffic_source": traffic_src,
                            "algorithmic_bytes_per_launch": per_seg_bytes * seg_per_launch,
                            "peak_source": "cr_measure_fma_peak (register-resident FMA micro-kernel, same run; MEASURED_PEAKS.json holds no FP32/FP64 vector peak)",
                            "flops_per_segment": per_seg_flops, "bytes_per_segment": per_seg_bytes, "segments_per_launch": seg_per_launch,
                            "avg_launch_ms": dur_s * 1e3, "kernel_share_of_step": ms_trace / (ms_total if world == 1 else sum(s["ms_total"] for s in stats)),
                            "ms_trace": ms_trace / args.steps, "ms_shade": ms_shade / args.steps, "ms_raygen": ms_gen / args.steps,
                            "fp64_fma_peak_tflops": f64p.value, "fp32_fma_peak_tflops": f32p.value,
                            # the same launches seen as a memory kernel: SURVEY 8d's algorithmic bytes against the measured HBM
                            # copy peak.  The scene (16 KB of nodes) is cache resident, so this view only shows that HBM is NOT
                            # the bound of this config (the DRAM traffic of the launch is `traffic`); it IS the bound of configs[3].
                            "hbm_view": hbm_view(per_seg_bytes * seg_per_launch, dur_s, traffic)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
