#!/usr/bin/env python3
"""bench.py — Msamples/s (and Mrays/s) of the radiance loop on BASELINE.json's configurations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2] [--impl ptb|reference]

A step is one full render of the named configuration (default C2 = BASELINE.json configs[1]: 1,000,000-triangle
Phong mesh + HDR-style envmap, 1024x1024, 256 spp, depth 5) on synthetic, closed-form inputs.  Scene generation,
BVH build and upload happen once before the timed region and are reported separately.
  value   whole-job Msamples/s with the scene resident in HBM and the frame left in HBM (ptb_render_accum)
  e2e     the same through the reference-shaped call Raytracer::render_image_nopreviz() with HOST buffers:
          camera + parameters go host->device, imagedouble + sample_count + 8-bit image come back every step
  roofline  the dominant kernel (k_trace, closest-hit BVH8 traversal) against the measured HBM copy bandwidth
  cpu_baseline  the reference's own CPU code (oracle/_ref) or its C restatement (oracle/port) on this box's cores
For N > 1 (torchrun, one rank per GPU) the frame is tile-sharded and gathered once over NCCL; value = samples of
all ranks / max-over-ranks time: strong scaling on the fixed frame.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

WORKLOADS = {
    "C1": "default parametric scene (spheres + spherical light, Phong) 512x512, 64 spp, depth 5",
    "C2": "synthetic 1M-triangle diffuse/Phong mesh + envmap, 1024x1024, 256 spp, depth 5",
    "C3": "synthetic 2.5M-triangle fully transparent mesh (Fresnel) + normal/alpha maps, 1920x1080, 512 spp, depth 5",
    "C4": "synthetic MERL-format BRDF (90x90x180) on a 260k-triangle mesh with DoF, 1024x1024, 1024 spp, depth 5",
    "C5": "synthetic 24M-triangle scene, 3840x2160, 1024 spp, depth 5",
}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load": samples in the upper half of the power draw seen
        thr = 0.5 * max(pw)
        load = [s for s, p in zip(sm, pw) if p >= thr] or sm
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm), "reasons": sorted(reasons)}


def make_rt(lib, workload, device=0):
    from pathtracer_b200 import scenes
    return scenes.CONFIGS[workload](lib, device=device)


def cpu_reference(workload, steps, warmup, budget_s=15.0, emit=True):
    """The reference's own CPU implementation of the path on the host cores: oracle/_ref if it is here, else oracle/port."""
    from oracles import port_lib, ref_lib
    from pathtracer_b200 import _abi
    lib = ref_lib()
    kind = "reference"
    if lib is None:
        lib, kind = port_lib(), "port"
    cores = min(os.cpu_count() or 1, 64)       # the reference is hard-limited to 64 threads (Vector.h:29, Raytracer.h:114)
    rt = make_rt(lib, workload)
    full_spp = rt.nrays
    t0 = time.time()
    rt.commit()
    build_s = time.time() - t0
    rt.set_option(_abi.ORC_OPT_THREADS, cores)
    # bounded sample: full resolution, reduced spp (throughput is spp-independent), sized from a 1-spp probe
    rt.nrays = 1
    t0 = time.time(); rt.render_image_nopreviz(want_image=False); probe = time.time() - t0
    n_steps = max(1, steps) + max(0, warmup)
    spp = int(max(1, min(full_spp, budget_s / max(probe, 1e-3) / n_steps)))
    rt.nrays = spp
    times, rays = [], 0
    for i in range(n_steps):
        t0 = time.time(); rt.render_image_nopreviz(want_image=False); dt = time.time() - t0
        if i >= warmup:
            times.append(dt); rays = rt.stats["rays_closest"] + rt.stats["rays_shadow"]
    samples = rt.W * rt.H * spp
    ms = 1e3 * sum(times) / len(times)
    value = samples / ms / 1e3
    base = {"value": value, "unit": "Msamples/s", "cores": cores, "kind": kind,
            "sample": f"{workload} at full resolution {rt.W}x{rt.H}, {spp} spp of {full_spp} per step, depth {rt.nb_bounces}; BVH build {build_s:.1f}s excluded",
            "mrays_per_s": rays / ms / 1e3}
    rt.close()
    return base, ms, spp


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ptb", choices=["ptb", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    base_line = {"metric": "Msamples/s", "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                 "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                 "config": {"workload": f"{args.workload}: {WORKLOADS[args.workload]}", "spp_sharding": "image tiles (32x32 when shared, rows rotated), tile_id % n_gpus",
                            "pass_pipelines": int(os.environ.get("PTB_PIPES", "2")),
                            "l2": "working set per step (BVH + triangles + path pool, >2 GB) exceeds the 126 MB L2; no flush needed"}}

    if args.impl == "reference":
        if rank != 0:
            return 0
        cb, ms, spp = cpu_reference(args.workload, args.steps, args.warmup)
        line = dict(base_line)
        line.update({"impl": "reference", "value": cb["value"], "ms_per_step": ms, "n_gpus": args.gpus, "cpu_baseline": cb,
                     "e2e": {"value": cb["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})
        print(json.dumps(line), flush=True)
        return 0

    import torch
    import torch.distributed as dist
    import pathtracer_b200
    from pathtracer_b200 import _abi, multi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    lib = pathtracer_b200.load()
    N_PIPES = int(os.environ.get("PTB_PIPES", "2"))
    t0 = time.time(); rt = make_rt(lib, args.workload, device=local); gen_s = time.time() - t0
    t0 = time.time(); rt.commit(); commit_s = time.time() - t0
    rt.reuse_buffers = True     # like the reference, whose Raytracer owns its output vectors (Raytracer.h:90-105)
    info = rt.scene_info()
    samples_frame = rt.W * rt.H * rt.nrays

    def sync():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()

    rgbw = torch.zeros(rt.H * rt.W * 4, dtype=torch.float32, device=device)

    def step_resident():
        """inputs resident, frame left in HBM (rank 0 holds the gathered frame for N > 1)"""
        if world == 1:
            rgbw.zero_()
            return rt.render_accum(rgbw.data_ptr())
        return step_sharded(False)

    def step_sharded(to_host):
        L, ctx = rt.lib, rt._ctx
        rgbw.zero_()
        st = rt.render_accum(rgbw.data_ptr(), rank, world)
        sizes = []
        for r in range(world):
            n = C.c_int64(0); p = rt.params(r, world)
            L.check(L.shard_pack_size(C.byref(p), r, C.byref(n)), ctx); sizes.append(n.value)
        nmax = max(max(sizes), 4)
        packed = torch.zeros(nmax, dtype=torch.float32, device=device)
        p = rt.params(rank, world)
        if rank != 0 and sizes[rank]:
            L.check(L.shard_pack(ctx, C.byref(p), rank, C.c_void_p(rgbw.data_ptr()), C.c_void_p(packed.data_ptr())), ctx)
        if rank == 0:
            bufs = [torch.empty(nmax, dtype=torch.float32, device=device) for _ in range(world)]
            dist.gather(packed, gather_list=bufs, dst=0)
            torch.cuda.synchronize(device)
            for r in range(1, world):
                if sizes[r]:
                    pr = rt.params(r, world)
                    L.check(L.shard_unpack_add(ctx, C.byref(pr), r, C.c_void_p(bufs[r].data_ptr()), C.c_void_p(rgbw.data_ptr())), ctx)
            if to_host:
                rt.resolve(rgbw.data_ptr(), True)
        else:
            dist.gather(packed, gather_list=None, dst=0)
        return st

    def step_e2e():
        if world == 1:
            rt.render_image_nopreviz(want_image=True)
            return rt.stats
        return step_sharded(True)

    def timed(fn, k):
        sync()
        t0 = time.perf_counter()
        dev_ms, launches, rays = 0.0, 0, 0
        for _ in range(k):
            st = fn()
            dev_ms += st["ms_device"]; launches += st["kernel_launches"]; rays += st["rays_closest"] + st["rays_shadow"]
        sync()
        wall_ms = 1e3 * (time.perf_counter() - t0)
        t = torch.tensor([wall_ms, dev_ms], dtype=torch.float64, device=device)
        tot = torch.tensor([launches, rays], dtype=torch.int64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        return float(t[0]), float(t[1]), int(tot[0]), int(tot[1])

    # instrumented, untimed pass: traversal counters for the roofline's algorithmic bytes (same kernel, same config)
    rt.set_option(_abi.OPT_COUNT_TRAVERSAL, 1)
    full_spp = rt.nrays
    rt.nrays = min(full_spp, 8)
    rgbw.zero_(); rt.render_accum(rgbw.data_ptr(), rank, world)
    kt = rt.kernel_times()
    n_node = kt["extend"]["node_visits"] / max(1, kt["extend"]["items"])
    n_tri = kt["extend"]["tri_tests"] / max(1, kt["extend"]["items"])
    rt.nrays = full_spp
    rt.set_option(_abi.OPT_COUNT_TRAVERSAL, 0)

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    wall_ms, dev_ms, launches, rays = timed(step_resident, args.steps)
    # per-kernel CUDA-event timing (two event records per launch on the launching stream).  The timed region above runs the passes
    # on two pipelines (streams) whose kernels overlap, so a kernel's own duration is taken from one more step of the same
    # workload, right here, with the pipelines serialised: `roofline.launch_ms` is the kernel running alone on the GPU.
    rt.set_option(_abi.OPT_PIPES, 1)
    rt.set_option(_abi.OPT_TIME_KERNELS, 1)
    step_resident()
    kt = rt.kernel_times()
    serial_ms = rt.stats["ms_device"] if world == 1 else None
    rt.set_option(_abi.OPT_TIME_KERNELS, 0)
    rt.set_option(_abi.OPT_PIPES, N_PIPES)
    for _ in range(1):
        step_e2e()
    e2e_ms, _, _, _ = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    ms_per_step = wall_ms / args.steps
    value = samples_frame / ms_per_step / 1e3
    line = dict(base_line)
    line.update({"value": value, "ms_per_step": ms_per_step, "device_ms_per_step": dev_ms / args.steps,
                 "mrays_per_s": rays / args.steps / ms_per_step / 1e3, "rays_per_sample": rays / args.steps / samples_frame,
                 "gpu_launches": launches, "clocks": clocks,
                 "e2e": {"value": samples_frame / (e2e_ms / args.steps) / 1e3, "unit": "Msamples/s",
                         "h2d_bytes_per_step": C.sizeof(_abi.Camera) + C.sizeof(_abi.Params), "d2h_bytes_per_step": rt.W * rt.H * (12 + 4 + 3)},
                 "setup": {"scene_gen_s": gen_s, "commit_s": commit_s, "bvh_build_ms": info["ms_bvh_build"], "upload_ms": info["ms_upload"],
                           "triangles": info["n_triangles"], "bvh8_nodes": info["n_bvh_nodes"], "bvh8_depth": info["bvh_depth"],
                           "bytes_nodes": info["bytes_nodes"], "bytes_triangles": info["bytes_triangles"]}})
    # ---- roofline of the dominant kernel ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    ext = kt["extend"]
    bytes_per_ray = 80 * n_node + 48 * n_tri + 48
    ext_launch_ms = ext["ms"] / max(1, ext["launches"])
    rays_per_launch = ext["items"] / max(1, ext["launches"])
    achieved = rays_per_launch * bytes_per_ray / (ext_launch_ms * 1e-3) / 1e9 if ext_launch_ms > 0 else 0.0
    traffic, ncu = None, None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        rec = json.load(open(tpath)).get(args.workload, {})
        traffic = rec.get("k_trace_closest_dram_bytes_per_launch")
        ncu = {k: rec[k] for k in ("issue_active_pct", "active_lanes_per_instruction", "fma_pipe_active_pct", "dram_throughput_pct", "l2_throughput_pct",
                                   "l1_hit_pct", "l2_hit_pct", "source") if k in rec} or None
    step_kernel_ms = sum(v["ms"] for v in kt.values())
    line["roofline"] = {"bound": "hbm", "kernel": "k_trace<closest-hit> (BVH8 traversal)", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                        "peak_source": peak_src,
                        # counters of the same kernel from the committed ncu capture (profiles/): the traversal is bound by instruction issue
                        # under divergence (ALU / FMA pipes), its DRAM traffic is a few percent of the algorithmic bytes (L1 / L2 hits)
                        "ncu": ncu, "bytes_per_ray": bytes_per_ray, "n_node": n_node, "n_tri": n_tri,
                        "rays_per_launch": rays_per_launch, "launch_ms": ext_launch_ms, "launches_per_step": ext["launches"],
                        "share_of_step": ext["ms"] / step_kernel_ms if step_kernel_ms else None,
                        "timing": "CUDA events around every launch of one extra step with the pass pipelines serialised (PTB_OPT_PIPES=1), taken between the timed region and the e2e region",
                        "serialised_step_ms": serial_ms,
                        "kernel_ms_per_step": {k: v["ms"] for k, v in kt.items()}}
    if world > 1:
        dist.barrier()
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                cb, _, _ = cpu_reference(args.workload, 1, 0, budget_s=15.0)
                line["cpu_baseline"] = cb
            except Exception as e:  # the baseline is a reported figure; never lose the GPU line over it
                line["cpu_baseline"] = {"value": None, "unit": "Msamples/s", "cores": None, "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
