#!/usr/bin/env python
"""bench.py — BASELINE.json's metric (Msamples/s, Mrays/s) on BASELINE.json's configs.

  python bench.py --gpus N --steps K --warmup W [--config book1|cornell|teapot|instanced|walkthrough]
  python bench.py --impl reference ...          # the reference algorithm (oracle port) on the host cores

A "step" is one full pass of the hot path over one batch of synthetic input: one `Camera::render` of the seeded scene
at the config's full resolution / spp / depth (default config: book1 1920x1080, 100 spp, depth 50 = 207.36 M camera
samples; `walkthrough`: 24 consecutive frames of the 240-frame movie).  For N > 1 a still is sharded by interleaved
row blocks over the ranks (strong scaling; the only exchange is the framebuffer: every rank's resolve kernel stores
its rows into rank 0's image over NVLink), a movie by whole frames (no exchange).  Timing: CUDA events around every
step on the launching stream, L2 flushed between steps, max over ranks.
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

# BASELINE.json configs[0..4].  `bound`: what SURVEY 8d names as the roofline of the trace kernel on that config.
CONFIGS = {
    "book1": dict(kw=dict(image_width=1920, samples=100, seed=1), bound="fp64",
                  title="book1 end scene 1920x1080, 100 spp, max depth 50 (BASELINE configs[0]; the config the metric is quoted on)"),
    "cornell": dict(kw=dict(image_width=1024, samples=1000), bound="fp64",
                    title="Cornell box of quads + emissive light 1024x1024, 1000 spp, max depth 50 (BASELINE configs[1])"),
    "teapot": dict(kw=dict(image_width=1920, samples=256), bound="fp64",
                   title="teapot.obj mesh + spherical sky 1920x1080, 256 spp, max depth 50 (BASELINE configs[2])"),
    "instanced": dict(kw=dict(image_width=3840, samples=64), bound="hbm",
                      title="9 998 240 flattened teapot triangles + earthmap textures 3840x2160, 64 spp, max depth 50 (BASELINE configs[3])"),
    "walkthrough": dict(kw=dict(image_width=1920, samples=64, seed=1), bound="fp64", frames_per_step=24,
                        title="book1 walk-through 1920x1080, 64 spp, max depth 50: 24 consecutive frames of the 240-frame movie per step "
                              "(BASELINE configs[4]), whole frames sharded over the ranks"),
}
TRAFFIC_FILES = {"book1": "trace_book1_traffic.json", "instanced": "trace_cfg4_traffic.json", "cornell": "trace_cornell_traffic.json",
                 "teapot": "trace_teapot_traffic.json"}
SEED = 1


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


def build_config(name, samples=0):
    from crucible_b200 import demo_builder

    cfg = CONFIGS[name]
    kw = dict(cfg["kw"])
    if samples:
        kw["samples"] = samples
    sc = demo_builder.CONFIGS[name](**kw)
    return cfg, sc, sc.describe(), sc.scene_cam.to_abi()


_ORACLE_SCENES = {}


def oracle_sample(desc, cam, row_step, threads=0, first_row=0):
    """The CPU path (oracle port of the reference algorithm) on a bounded sample: rows first_row, first_row + row_step, ...
    of the same image at full spp.  Returns the oracle's stats (samples, rays, seconds, traversal counters)."""
    from oracle import binding as oracle

    orc = _ORACLE_SCENES.get(id(desc))
    if orc is None:
        orc = _ORACLE_SCENES[id(desc)] = oracle.OracleScene(desc)  # config 4: the oracle builds its tree on one core (~70 s, untimed)
    threads = threads or (os.cpu_count() or 1)
    _, _, st = orc.render(cam, seed=SEED, rows=(first_row, cam.image_height, row_step), threads=threads, want_rgb8=False)
    return st


def sized_oracle_sample(desc, cam, seconds, threads):
    """Probe with the two rows at 1/3 and 2/3 of the image height, then size an evenly spread row sample for about
    `seconds` of CPU work (at least one row: Cornell's 1000 spp make a single row 1 M samples)."""
    H, W = cam.image_height, cam.image_width
    # the probe needs a row per thread and more, or it measures thread start-up (two rows on 16 threads
    # underestimated the rate 4x and the sized sample then ran 1.7 s instead of 15 s)
    probe_step = max(1, H // max(2 * threads, 8))
    probe = oracle_sample(desc, cam, probe_step, threads, first_row=probe_step // 2)
    rate = probe["samples"] / max(probe["seconds"], 1e-6)
    n_rows = max(1.0, rate * seconds / (W * cam.samples))
    row_step = int(max(1, min(H, round(H / n_rows))))
    return oracle_sample(desc, cam, row_step, threads, first_row=row_step // 2), row_step


def measured_hbm_peak():
    peak, src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        if "hbm_gbs" in mp:
            peak, src = float(mp["hbm_gbs"]), "MEASURED_PEAKS.json:hbm_gbs"
    except Exception:
        pass
    return peak, src


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path, timed on the host cores.
    The reference is a Rust crate and no rustc/cargo exists here, so the timed code is the oracle port
    (kind "port"): same algorithm and arithmetic, world shared read-only, all host threads."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    cfg, sc, desc, cam = build_config(args.config, args.samples)
    threads = os.cpu_count() or 1
    H, W = cam.image_height, cam.image_width
    st, row_step = sized_oracle_sample(desc, cam, 6.0, threads)  # one step = roughly 6 s of wall time on this box
    for _ in range(args.warmup):
        oracle_sample(desc, cam, max(row_step, H // 2), threads, first_row=H // 2)
    secs, samples, rays = 0.0, 0, 0
    for _ in range(args.steps):
        st = oracle_sample(desc, cam, row_step, threads, first_row=row_step // 2)
        secs += st["seconds"]
        samples += st["samples"]
        rays += st["rays"]
    val = samples / secs / 1e6
    sample_desc = f"rows j % {row_step} == {row_step // 2} of the {W}x{H}x{cam.samples}spp job ({samples // args.steps} samples per step)"
    line = {"impl": "reference", "metric": "Msamples/s", "value": val, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": cfg["title"], "sample": sample_desc},
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
    ap.add_argument("--config", default="book1", choices=list(CONFIGS))
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--pool", type=int, default=0)
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="N > 1 framebuffer exchange (crucible_b200.multigpu)")
    ap.add_argument("--samples", type=int, default=0, help="debug only: a reduced-spp run is NOT the headline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from crucible_b200 import abi, multigpu
    from crucible_b200.gpu import GpuScene

    rank, world, local = dist_env()
    assert torch.cuda.is_available(), "bench.py needs a B200: crucible_b200 has no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    precision = abi.CR_PRECISION_F64 if args.precision == "f64" else abi.CR_PRECISION_F32

    cfg, sc, desc, cam = build_config(args.config, args.samples)
    H, W = cam.image_height, cam.image_width
    movie = "frames_per_step" in cfg
    fps = cfg.get("frames_per_step", 1)
    gs = GpuScene(desc, local)
    row_block = 8
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    frame_buf = torch.empty((H, W, 3), dtype=torch.uint8, device=dev) if movie else None
    stream = torch.cuda.current_stream(dev).cuda_stream
    step_no = [0]

    def add_stats(acc, st):
        for k, v in st.items():
            acc[k] = v if k == "trace_engine" else acc.get(k, 0) + v  # an id, not a quantity: the last frame's
        return acc

    def step(time_kernels):
        """One pass of the hot path over one batch, outputs left in device memory (bytes: what Camera::render consumes)."""
        if not movie:
            _, full8, st = multigpu.render_sharded(gs, cam, rank, world, seed=SEED, precision=precision, row_block=row_block,
                                                   pool_paths=args.pool, time_kernels=time_kernels, want_rgb=False, exchange=args.exchange)
            return st
        first = (step_no[0] * fps) % 240
        step_no[0] += 1
        acc = {}
        for f in range(first + rank, first + fps, world):  # whole frames: frame f -> rank f % world within the step
            cam.frame = f
            add_stats(acc, gs.render_device(cam, 0, frame_buf.data_ptr(), stream=stream, seed=SEED, precision=precision,
                                            pool_paths=args.pool, time_kernels=time_kernels, global_rows=True))
        cam.frame = 0
        return acc

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
        st = step(True)
        ev[k][1].record()
        stats.append(st)
        flush.fill_(k)  # L2 flush between timed iterations, outside the events
    barrier()
    clocks = sampler.stop(t_timed0, time.perf_counter()) if rank == 0 else None
    ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    agg = torch.tensor([sum(s.get("samples", 0) for s in stats), sum(s.get("rays", 0) for s in stats), sum(s.get("launches", 0) for s in stats),
                        sum(s.get("iterations", 0) for s in stats)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    ms_total = float(t.item())
    samples, rays, launches, iters = (float(x) for x in agg.tolist())

    # ---- e2e: the public API with HOST buffers.  Every step: scene description -> cr_scene_* (BVH build as
    # Scene::render_image does, scene/mod.rs:333) -> H2D -> render -> framebuffer exchange -> D2H of the bytes the
    # reference writes to its PPM (camera/mod.rs:306-311).
    h2d = sum(b[1].nbytes + b[2].nbytes + b[3].nbytes for b in desc.batches) + sum(im.nbytes for im in desc.images) + abi.C.sizeof(abi.CrCamera)
    d2h = H * W * 3 * (fps if movie else 1)
    pinned8 = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
    pinned8_np = pinned8.numpy()

    e2e_trace = os.environ.get("BENCH_E2E_TRACE") and rank == 0

    def e2e_step():
        ta = time.perf_counter()
        g = GpuScene(desc, local)
        tb = time.perf_counter()
        if movie:
            for f in range(rank, fps, world):
                cam.frame = f
                g.render(cam, seed=SEED, precision=precision, pool_paths=args.pool, want_rgb=False, out_rgb8=pinned8_np)
            cam.frame = 0
        elif world == 1:
            # the reference-facing call itself: cr_render with HOST buffers (H2D of the camera, D2H of the image inside)
            g.render(cam, seed=SEED, precision=precision, pool_paths=args.pool, want_rgb=False, out_rgb8=pinned8_np)
        else:
            _, full8, st = multigpu.render_sharded(g, cam, rank, world, seed=SEED, precision=precision, row_block=row_block,
                                                   pool_paths=args.pool, want_rgb=False, exchange=args.exchange)
            if rank == 0:
                pinned8.copy_(full8, non_blocking=True)
            torch.cuda.synchronize()
        tc = time.perf_counter()
        g.close()
        if e2e_trace:
            print(f"e2e step: scene {1e3 * (tb - ta):.1f} ms, render {1e3 * (tc - tb):.1f} ms, close {1e3 * (time.perf_counter() - tc):.1f} ms", file=sys.stderr, flush=True)

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
        total_samples_per_step = H * W * cam.samples * fps
        value = samples / (ms_total * 1e-3) / 1e6
        par = f"frames x{world}" if movie else f"rows x{world} (blocks of {row_block}), exchange {args.exchange if world > 1 else 'none'}"
        line = {"metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
                "config": {"workload": cfg["title"], "name": args.config, "parallelism": par, "l2": "flushed between steps (256 MiB write)",
                           "seed": SEED, "samples_per_step": total_samples_per_step, "spp": cam.samples, "output": "rgb8 bytes (device resident for `value`)"},
                "mrays_per_s": rays / (ms_total * 1e-3) / 1e6,
                "rays_per_sample": rays / samples,
                "clocks": clocks,
                "e2e": {"value": total_samples_per_step * args.steps / e2e_s / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h},
                "gpu_launches": int(launches)}
        # ---- roofline of the dominant kernel (k_trace): algorithmic work per ray segment from the oracle's
        # reference-order traversal counters, segments per launch from the run, duration from events
        ms_trace = sum(s.get("ms_trace", 0) for s in stats)
        ms_shade = sum(s.get("ms_shade", 0) for s in stats)
        ms_gen = sum(s.get("ms_raygen", 0) for s in stats)
        my_iters = sum(s.get("iterations", 0) for s in stats)
        my_rays = sum(s.get("rays", 0) for s in stats)
        my_ms = sum(s.get("ms_total", 0) for s in stats)
        f64p, f32p = abi.C.c_double(), abi.C.c_double()
        abi.check(abi.load().cr_measure_fma_peak(local, abi.C.byref(f64p), abi.C.byref(f32p)))
        threads = os.cpu_count() or 1
        if not args.no_cpu_baseline:
            cpu, row_step = sized_oracle_sample(desc, cam, 15.0, threads)  # ~15 s of CPU work
            line["cpu_baseline"] = {"value": cpu["samples"] / cpu["seconds"] / 1e6, "unit": "Msamples/s", "cores": threads, "kind": "port",
                                    "sample": f"rows j % {row_step} == {row_step // 2} of the same {W}x{H}x{cam.samples}spp job ({cpu['samples']} samples, {cpu['seconds']:.1f} s)",
                                    "mrays_per_s": cpu["rays"] / cpu["seconds"] / 1e6}
        else:
            cpu = oracle_sample(desc, cam, max(1, H // 3), threads, first_row=H // 3)
        # flops = 24 N_node + 40 N_sph + 50 N_tri + 60, bytes = 32 N_node + 16 N_sph + 36 N_tri + 16 per segment;
        # the Quad extension is counted as 45 flop / 64 B per test (plane + two cross-dot products; f32 record)
        per_seg_flops = (24.0 * cpu["node_tests"] + 40.0 * cpu["sphere_tests"] + 50.0 * cpu["tri_tests"] + 45.0 * cpu["quad_tests"]) / cpu["rays"] + 60.0
        per_seg_bytes = (32.0 * cpu["node_tests"] + 16.0 * cpu["sphere_tests"] + 36.0 * cpu["tri_tests"] + 64.0 * cpu["quad_tests"]) / cpu["rays"] + 16.0
        seg_per_launch = my_rays / max(my_iters, 1)
        dur_s = ms_trace * 1e-3 / max(my_iters, 1)
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", TRAFFIC_FILES.get(args.config, ""))
        if args.precision == "f64" and os.path.isfile(tpath):
            tj = json.load(open(tpath))  # DRAM bytes of one ncu --set full capture, scaled to this run's segments per launch
            traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) / tj["rays_in_launch"] * seg_per_launch
            traffic_src = tj["capture"]
        fma_peak = f64p.value if args.precision == "f64" else f32p.value
        flops_ach = per_seg_flops * seg_per_launch / dur_s / 1e12
        hbm_peak, hbm_src = measured_hbm_peak()
        bytes_ach = per_seg_bytes * seg_per_launch / dur_s / 1e9
        fp_view = {"bound": "fp64" if args.precision == "f64" else "fp32", "achieved": flops_ach, "peak": fma_peak, "unit": "TFLOP/s",
                   "frac": flops_ach / fma_peak if fma_peak else None,
                   "peak_source": "cr_measure_fma_peak (register-resident FMA micro-kernel, same run; MEASURED_PEAKS.json holds no FP32/FP64 vector peak)"}
        hbm_view = {"bound": "hbm", "achieved": bytes_ach, "peak": hbm_peak, "unit": "GB/s", "frac": bytes_ach / hbm_peak, "peak_source": hbm_src,
                    "dram_gbps_measured": (traffic / dur_s / 1e9) if traffic else None}
        # configs 0-2, 4: the scene is cache resident => the bound SURVEY 8d names is the FP issue rate; config 3 (11.6 M nodes,
        # 1.1 GB of node + triangle records) => HBM.  The other view is reported next to it.
        if cfg["bound"] == "hbm" and traffic and stats and stats[-1].get("trace_engine") in (2, 3):
            # The order-free engine does not walk the reference's tree, so SURVEY 8d's per-segment bytes (the reference-order
            # counters) are not what it reads: against them the figure exceeds 1.  The HBM bound is therefore stated on the
            # DRAM bytes the kernel really moved (ncu capture of the same launch shape); the 8d view is kept beside it.
            hbm_view = dict(hbm_view, basis="SURVEY 8d bytes of the reference-order traversal (not what this engine reads; > 1 = it needs fewer)")
            line_hbm = {"bound": "hbm", "achieved": traffic / dur_s / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": traffic / dur_s / 1e9 / hbm_peak,
                        "peak_source": hbm_src, "basis": "DRAM bytes per launch measured by ncu (traffic), the engine's real reads and writes",
                        "survey_8d_view": hbm_view}
            hbm_view = line_hbm
        main_view, other = (hbm_view, fp_view) if cfg["bound"] == "hbm" else (fp_view, hbm_view)
        kernel_name = {0: "k_trace (reference order, 64 registers)", 1: "k_trace (reference order, 48 registers)", 2: "k_trace_fast", 3: "k_trace_fast_smem"}.get(
            stats[-1].get("trace_engine") if stats else None, "k_trace")
        line["roofline"] = dict(main_view, kernel=kernel_name, traffic=traffic, traffic_source=traffic_src,
                                algorithmic_bytes_per_launch=per_seg_bytes * seg_per_launch, flops_per_segment=per_seg_flops,
                                bytes_per_segment=per_seg_bytes, segments_per_launch=seg_per_launch, avg_launch_ms=dur_s * 1e3,
                                kernel_share_of_step=ms_trace / max(my_ms, 1e-9), ms_trace=ms_trace / args.steps, ms_shade=ms_shade / args.steps,
                                ms_raygen=ms_gen / args.steps,
                                trace_engine={0: "reference order, 64 registers", 1: "reference order, 48 registers", 2: "order-free", 3: "order-free, search tree in shared memory"}.get(stats[-1].get("trace_engine"), None) if stats else None,
                                retried_rays=sum(s.get("retried_rays", 0) for s in stats), fp64_fma_peak_tflops=f64p.value, fp32_fma_peak_tflops=f32p.value)
        line["roofline"]["fp_view" if cfg["bound"] == "hbm" else "hbm_view"] = other
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
