#!/usr/bin/env python3
"""Benchmark of the path-tracing hot path (BASELINE.json: Mpaths/s and Mrays/s on the named scenes).

    python bench.py --gpus N --steps K --warmup W            # our arm (torchrun launches N ranks)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU renderer (rank 0 only)

A "step" is one full render() of the workload.  Default workload: cornell_box exactly as
shipped (examples/cornell_box.cpp: 800x800, 64 bounces) at 128 spp PER GPU: with N GPUs the
render has 128*N samples per pixel and is sharded by interleaved 8-row strips, each rank
owning every N-th strip and all samples of its pixels (weak scaling: 81.92 M paths per GPU per
step); the partial films are combined by ONE NCCL sum-reduce.  The film is bit-identical to a
single-GPU render of the same configuration (ordered per-pixel accumulation, exact zeros
elsewhere).

`value`  whole-job Mpaths/s, scene and film resident in HBM, device-timed (CUDA events,
         max over ranks).
`e2e`    the same metric through the reference-facing call: at N = 1 the host library's
         render(camera, scene, spp, bounces) itself (harness entry qzh_render), which returns
         a RenderResult of pageable std::vectors -- per step the camera / sensor tables go
         host->device and the three film planes come back device->host, staged through the
         library's pinned area (`e2e.pinned_value`: the same through the C ABI with caller-pinned
         buffers).  At N > 1 each rank renders its strips into its device film, ONE NCCL reduce
         assembles them on rank 0's device and rank 0 copies the film to pinned host memory.
`secondary` (N = 1) the two BVH workloads -- obj_viewer on the synthetic 1M-triangle mesh and
         mandelbrot (2.87M triangles) -- a few steps each, with the traversal stage's roofline
         (algorithmic GB/s against the HBM peak, measured L2 and DRAM bytes, issue-slot share)
         and the scene-build times (OBJ parse, commit, GPU BVH build).
`roofline`     the dominant kernel stage against the measured HBM peak (MEASURED_PEAKS.json).
`cpu_baseline` the oracle (reference sources over the Embree shim -- NOT real Embree, which is
               absent) timed on this box's host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (scene, width, height, spp per GPU, max_bounces)
    "cornell_box": ("cornell_box", 800, 800, 128, 64),
    "cornell_4k": ("cornell_box", 3840, 2160, 128, 64),   # x8 GPUs = 1024 spp: BASELINE.json configs[4]
    "glass_spheres": ("glass_spheres", 800, 800, 512, 64),
    "textures": ("textures", 800, 800, 36, 32),
    "opposing_planes": ("opposing_planes", 1920, 1080, 256, 64),
    "obj_viewer": ("obj_viewer", 800, 600, 24, 32),         # synthetic ~1M-triangle mesh
    "mandelbrot": ("mandelbrot_full", 800, 800, 32, 12),    # examples/mandelbrot.cpp: 1200x1200 height grid (2.87M triangles), copper
}


def strip_rows_for(height: int, world: int) -> int:
    """Rows per interleaved strip: a height <= 8 that gives every rank the same number of rows (round 1's fixed 8-row
    strips gave 13 vs 12 strips per rank at 800 rows over 8 GPUs: 96 % efficiency from the imbalance alone), powers of
    two first: with strip * world dividing 128 a rank owns 1/world of the (y mod 128) pixel classes and its sample memo
    tabulates only those (csrc/sampler.cuh); 5-row strips over 8 ranks touch all 128 and the fill costs 8 x as much."""
    for rows in (8, 4, 2, 1, 7, 6, 5, 3):
        if height % rows == 0 and (height // rows) % world == 0:
            return rows
    return 1


def measured_peaks() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """SM clocks and throttle reasons of this rank's GPU during the timed region, read through NVML in-process (the library
    nvidia-smi is a front end of).  Spawning an `nvidia-smi` per rank every 200 ms -- the first version -- initialises
    every GPU of the box in a new process each time and holds driver locks while the renderer launches ~270 graphs per
    step: invisible at N = 1-2, -16 % at N = 4 and -31 % at N = 8 on the device-timed loop (the end-to-end loop, timed
    without the sampler, scaled 0.99).  `nvidia-smi` remains the fallback when NVML cannot be loaded."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()
        self._nvml = self._handle = None
        try:
            import pynvml

            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = index
            if visible and all(v.strip().isdigit() for v in visible.split(",")):
                phys = int(visible.split(",")[index])
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._nvml = pynvml
            self._max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        sm = float(n.nvmlDeviceGetClockInfo(self._handle, n.NVML_CLOCK_SM))
        try:
            mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self._handle))
        except Exception:
            mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self._handle))
        self.samples.append([str(sm), str(self._max)] + ["Active" if mask & bit else "Not Active" for bit in (0x8, 0x40, 0x20, 0x4)])

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.samples.append([f.strip() for f in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.15 if self._nvml is not None else 0.5)

    def stop(self) -> dict:
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) > 2 + i and s[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples), "source": "nvml" if self._nvml is not None else "nvidia-smi"}


class quiet_stdout:
    """render() prints its "Render time" line on stdout like the reference (render.cpp:392-394); the bench's stdout is
    ONE JSON line, so file descriptor 1 points at /dev/null while render() is being timed."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        self._null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self._null, 1)

    def __exit__(self, *exc):
        os.dup2(self._saved, 1)
        os.close(self._saved)
        os.close(self._null)


def ensure_mesh(n_tri: int = 1_000_000) -> str:
    sys.path.insert(0, str(ROOT / "tools"))
    import gen_mesh

    return gen_mesh.ensure_obj(os.environ.get("QZ_MESH_DIR", "/tmp"), n_tri)


def build_scene(harness, workload: str, width: int, height: int, mesh_tris: int):
    scene = WORKLOADS[workload][0]
    if scene == "obj_viewer":
        return harness.build_scene(scene, width, height, obj_path=ensure_mesh(mesh_tris), obj_material="alluminum", obj_light="point")
    return harness.build_scene(scene, width, height)


def oracle_library(workload: str) -> Path:
    """oracle/_ref: the reference's sources over the Embree shim.  The obj_viewer workload uses the variant whose ONLY
    difference is the OBJ text parser (oracle/Makefile, liboracle_ref_fastobj.so): everything that is timed is the reference's."""
    name = "liboracle_ref_fastobj.so" if WORKLOADS[workload][0] == "obj_viewer" else "liboracle_ref.so"
    return ROOT / "oracle" / "_ref" / name


def cpu_baseline(workload: str, width: int, height: int, max_bounces: int, mesh_tris: int, budget_s: float = 15.0) -> dict:
    """The oracle on this box's host cores, bounded sample: full resolution, reduced spp."""
    from quetzalcoatlus_b200.harness import Harness

    lib = oracle_library(workload)
    if not lib.exists():
        return {"value": None, "unit": "Mpaths/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
    orc = Harness(lib, "orc_")
    with build_scene(orc, workload, width, height, mesh_tris) as sc:
        probe = sc.render(1, max_bounces)
        spp = int(max(1, min(WORKLOADS[workload][3], budget_s / max(probe.seconds, 1e-3))))
        r = sc.render(spp, max_bounces) if spp > 1 else probe
    paths = width * height * spp
    return {"value": paths / r.seconds / 1e6, "unit": "Mpaths/s", "cores": r.n_threads, "kind": "reference",
            "mrays_per_s": r.rays / r.seconds / 1e6,
            "sample": f"{workload} {width}x{height}, {spp} spp of the workload's samples, {max_bounces} bounces, "
                      f"{r.seconds:.2f} s; reference integrator over the oracle's Embree shim (not real Embree)"
                      + (f"; the same {mesh_tris}-triangle OBJ text, parsed by the product's loader (oracle/Makefile: the reference's "
                         "regex loader needs ~25 min for it)" if WORKLOADS[workload][0] == "obj_viewer" else "")}


def run_reference(args) -> None:
    """The reference arm: the reference's own CPU renderer (oracle/_ref) on rank 0."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    _, w0, h0, spp0, mb = WORKLOADS[args.workload]
    width, height = args.width or w0, args.height or h0
    from quetzalcoatlus_b200.harness import Harness

    lib = oracle_library(args.workload)
    if not lib.exists():
        print(json.dumps({"impl": "reference", "unavailable": f"{lib.relative_to(ROOT)} is not built"}))
        return
    orc = Harness(lib, "orc_")
    with build_scene(orc, args.workload, width, height, args.mesh_triangles) as sc:
        probe = sc.render(1, mb)
        # each step = a bounded sample of the workload: as many of its spp as fit ~8 s
        spp = int(max(1, min(spp0 * args.gpus, 8.0 / max(probe.seconds, 1e-3))))
        for _ in range(min(args.warmup, 1)):
            sc.render(spp, mb)
        secs, rays, threads = [], 0, 0
        for _ in range(args.steps):
            r = sc.render(spp, mb)
            secs.append(r.seconds)
            rays, threads = r.rays, r.n_threads
    total = sum(secs)
    paths = width * height * spp
    value = paths * args.steps / total / 1e6
    sample = (f"{args.workload} {width}x{height}, {spp} of {spp0 * args.gpus} spp per step, {mb} bounces; reference "
              f"integrator over the oracle's Embree shim (not real Embree)")
    print(json.dumps({
        "impl": "reference", "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "width": width, "height": height, "spp_per_step": spp, "max_bounces": mb},
        "mrays_per_s": rays / secs[-1] / 1e6,
        "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def mesh_first_hit_fraction(sc, width: int, height: int) -> float:
    """obj_viewer: the share of camera rays (pixel centres, every 4th pixel of every 4th row) whose closest hit is the mesh
    (geomID 0: the OBJ is the scene's first geometry) -- through the closest-hit probe of the C ABI (qz_intersect)."""
    cf = sc.camera_fields()
    pos, bl, du, dv = cf[0], cf[4], cf[5], cf[6]
    xs, ys = np.meshgrid(np.arange(0, width, 4, dtype=np.float32) + 0.5, np.arange(0, height, 4, dtype=np.float32) + 0.5)
    d = bl[None, :] + du[None, :] * xs.reshape(-1, 1) + dv[None, :] * ys.reshape(-1, 1) - pos[None, :]
    rays = np.concatenate([np.broadcast_to(pos, d.shape), d], 1).astype(np.float32)
    hit = sc.intersect(rays)
    return float(((hit[:, 0] >= 0) & (hit[:, 6] == 0)).mean())


def load_counters() -> dict:
    """ncu counters per unit of work (profiles/r02_counters.json, written by tools/ncu_counters.py from the committed
    steady-state captures): DRAM bytes, L2 bytes and executed warp instructions per ray / per shaded bounce, per scene."""
    p = ROOT / "profiles" / "r02_counters.json"
    return json.loads(p.read_text()) if p.exists() else {}


def stage_report(inst: dict, scene_name: str, flags: int, sm_mhz: float | None) -> tuple[dict, dict]:
    """Roofline of the dominant stage from one instrumented step pair (traversal counters + per-stage events)."""
    n_rays = inst["rays_closest"] + inst["rays_shadow"]
    n_node = inst["node_visits"] / max(n_rays, 1)
    n_prim = inst["prim_tests"] / max(n_rays, 1)
    # B_ray = N_node*128 + N_prim*64 + 32 (ray read) + 32 (hit write; 4 for shadow rays)
    bytes_trav = inst["node_visits"] * 128 + inst["prim_tests"] * 64 + inst["rays_closest"] * 64 + inst["rays_shadow"] * 36
    # scenes of <= 96 primitives are intersected by the flat kernel (primitive records staged in shared memory
    # once per CTA, no BVH): per ray only the ray read and the result write are memory traffic
    n_prims_scene = (inst["bvh_bytes"] - inst["bvh_nodes"] * 128) // 64
    flat_scene = 0 < n_prims_scene <= 96 and not (flags & 8)
    if flat_scene:
        bytes_trav = inst["rays_closest"] * 64 + inst["rays_shadow"] * 36
    # per shaded bounce: the path record is read and written once (288 B); a bounce's draws are one 32-byte sector of
    # the path's memo row (late bounces: 32 B written by the sampler stage and read back)
    bytes_shade = inst["shade_calls"] * (288 + 32) + inst["rays_shadow"] * 48
    bytes_sample = inst["shade_calls"] * 64
    stages = {"traversal": (inst["ms_closest"] + inst["ms_shadow"], bytes_trav, n_rays),
              "shading": (inst["ms_shade"], bytes_shade, inst["shade_calls"]),
              "sampler": (inst["ms_sample"], bytes_sample, inst["shade_calls"])}
    dom = max(stages, key=lambda k: stages[k][0])
    d_ms, d_bytes, d_units = stages[dom]
    peak, how = measured_peaks()
    achieved = d_bytes / (d_ms * 1e-3) / 1e9 if d_ms > 0 else 0.0
    cnt = load_counters().get(scene_name, {}).get(dom)
    traffic = l2_gbs = issue = None
    if cnt:
        traffic = cnt["dram_bytes_per_unit"] * d_units
        l2_gbs = cnt["l2_bytes_per_unit"] * d_units / (d_ms * 1e-3) / 1e9 if d_ms > 0 else None
        clock = (sm_mhz or 1965.0) * 1e6
        issue_peak = 148 * 4 * clock                        # warp instructions per second: 148 SMs x 4 schedulers
        wips = cnt["warp_inst_per_unit"] * d_units / (d_ms * 1e-3) if d_ms > 0 else 0.0
        issue = {"warp_inst_per_s": wips, "peak": issue_peak, "frac": wips / issue_peak,
                 "active_lanes": cnt.get("threads_per_inst"), "source": "profiles/r02_counters.json x this step's counts"}
    kernel_names = {"traversal": "traversal (k_step_flat)" if flat_scene else "traversal (k_trace_lane<closest> + <any hit>)",
                    "shading": "shading (k_shade<family> + k_albedo_conductor)", "sampler": "sampler (k_sample)"}
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "algorithmic_bytes": d_bytes, "kernel": kernel_names[dom], "peak_source": how,
                "l2_gbs": l2_gbs, "dram_over_algorithmic": (traffic / d_bytes) if traffic else None, "issue": issue,
                "stage_ms": {**{k: v[0] for k, v in stages.items()}, "finish + regenerate + bin": inst["ms_other"], "step": inst["ms_total"]},
                "note": "algorithmic bytes per step (SURVEY.md 8.d) over the stage's device time in a serial (one-pipeline, "
                        "event-per-stage) step of this run; traffic / l2 / issue = ncu counters per unit of the committed "
                        "steady-state capture x this step's counts.  The analytic scenes' stages are bound by instruction "
                        "issue and latency, not HBM: `issue` is the roof that applies there (DESIGN.md section 6)"}
    extra = {"flat_intersection": flat_scene, "n_node_per_ray": n_node, "n_prim_per_ray": n_prim,
             "bounces_per_path": inst["shade_calls"] / max(inst["paths"], 1),
             "shadow_ray_fraction": inst["rays_shadow"] / max(n_rays, 1), "bvh_nodes": inst["bvh_nodes"],
             "whole_pipeline_algorithmic_gbs": (bytes_trav + bytes_shade + bytes_sample) / (inst["ms_total"] * 1e-3) / 1e9}
    return roofline, extra


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    os.environ["QZ_DEVICE"] = str(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from quetzalcoatlus_b200 import load_harness
    from quetzalcoatlus_b200.harness import (QZ_FLAG_COUNT_TRAVERSAL, QZ_FLAG_STAGE_TIMING, QzRegion, QzRenderOptions, QzStats)

    qz = load_harness()
    lib = qz.lib
    lib.qz_last_error.restype = ctypes.c_char_p
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(workload: str, spp_per_gpu: int, steps: int, warmup: int, with_e2e: bool) -> dict:
        """Device-timed steps (+ end-to-end steps) of one workload, and the instrumented step pair for its roofline."""
        _, w0, h0, spp0, mb = WORKLOADS[workload]
        spp = (spp_per_gpu or spp0) * world
        sc, build, width, height = scenes[workload]
        handle, cam = ctypes.c_void_p(sc.c_scene_handle()), sc.c_camera()
        strip_rows = strip_rows_for(height, world)
        region = QzRegion(strip_rows, world, rank)
        film = torch.zeros((3, height, width, 3), dtype=torch.float32, device="cuda")  # colour, normal, albedo planes

        def step(flags: int = 0) -> dict:
            """One render into the device-resident film + (N > 1) the single NCCL reduce."""
            if world > 1:
                film.zero_()
            st = QzStats()
            opts = QzRenderOptions(flags | args.flags, args.pool, 0, 0)
            rc = lib.qz_render_device(handle, ctypes.byref(cam), spp, mb, ctypes.byref(region), ctypes.byref(opts),
                                      ctypes.c_void_p(film[0].data_ptr()), ctypes.c_void_p(film[1].data_ptr()),
                                      ctypes.c_void_p(film[2].data_ptr()), ctypes.c_void_p(stream.cuda_stream), ctypes.byref(st))
            if rc != 0:
                raise RuntimeError(f"qz_render_device failed: {lib.qz_last_error().decode()}")
            if world > 1:
                dist.reduce(film, dst=0, op=dist.ReduceOp.SUM)
            return st.as_dict()

        for _ in range(warmup):
            step()
        barrier()
        # (rank 0 samples its own GPU: NVML queries take a driver lock that every rank's launches go through)
        clocks = ClockSampler(local_rank) if rank == 0 else None
        if clocks:
            clocks.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        stats = [step() for _ in range(steps)]
        e1.record(stream)
        barrier()
        clock_info = clocks.stop() if clocks else {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "rank 0 only"}
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms_total = float(ms.item())
        sums = torch.tensor([float(sum(s["rays_closest"] + s["rays_shadow"] for s in stats)),
                             float(sum(s["kernel_launches"] for s in stats))], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        mesh_frac = mesh_first_hit_fraction(sc, width, height) if WORKLOADS[workload][0] == "obj_viewer" and rank == 0 else None
        out = {"workload": workload, "scene": WORKLOADS[workload][0], "width": width, "height": height, "spp": spp, "max_bounces": mb,
               "mesh_first_hit_fraction": mesh_frac,
               "strip_rows": strip_rows, "ms_per_step": ms_total / steps, "value": width * height * spp * steps / (ms_total * 1e-3) / 1e6,
               "mrays_per_s": float(sums[0].item()) / (ms_total * 1e-3) / 1e6, "gpu_launches": int(sums[1].item()),
               "clocks": clock_info, "build": build}

        if with_e2e:
            film_bytes = 3 * height * width * 3 * 4
            host_film = torch.zeros((3, height, width, 3), dtype=torch.float32).pin_memory()
            host_ptrs = [ctypes.c_void_p(host_film[k].data_ptr()) for k in range(3)]

            def e2e_render():
                """N = 1: the host library's render() itself, returning a RenderResult of pageable std::vectors."""
                return sc.render_only(spp, mb)

            def e2e_pinned():
                st = QzStats()
                opts = QzRenderOptions(args.flags, args.pool, 0, 0)
                rc = lib.qz_render(handle, ctypes.byref(cam), spp, mb, None, ctypes.byref(opts), host_ptrs[0], host_ptrs[1], host_ptrs[2], ctypes.byref(st))
                assert rc == 0, lib.qz_last_error().decode()
                return float(host_film[0, 0, 0, 0])

            def e2e_sharded():
                """N > 1: strips into the device film, one NCCL reduce on the devices, rank 0 reads the film on the host."""
                step()
                if rank == 0:
                    host_film.copy_(film, non_blocking=True)
                    torch.cuda.synchronize()
                return float(host_film[0, 0, 0, 0])

            def time_fn(fn, n):
                for _ in range(2):
                    fn()
                barrier()
                t0 = time.perf_counter()
                for _ in range(n):
                    fn()
                barrier()
                dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
                return width * height * spp * n / float(dt.item()) / 1e6

            n_e2e = max(1, steps)
            if world == 1:
                # (flags / pool overrides only reach the C-ABI call: render() has no options argument)
                with quiet_stdout():
                    value_e2e = time_fn(e2e_render, n_e2e)
                pinned = time_fn(e2e_pinned, n_e2e)
                out["e2e"] = {"value": value_e2e, "unit": "Mpaths/s", "h2d_bytes_per_step": 3 * 471 * 4 + height * 4,
                              "d2h_bytes_per_step": film_bytes, "call": "render(camera, scene, spp, bounces) -> RenderResult (pageable std::vector planes)",
                              "pinned_value": pinned, "pinned_call": "qz_render() with caller-pinned buffers"}
            else:
                value_e2e = time_fn(e2e_sharded, n_e2e)
                out["e2e"] = {"value": value_e2e, "unit": "Mpaths/s", "h2d_bytes_per_step": (3 * 471 * 4 + height * 4) * world,
                              "d2h_bytes_per_step": film_bytes, "call": "qz_render_device() per rank + one NCCL reduce on the devices + rank 0 device->host"}

        # ---- roofline of the dominant stage: one instrumented step pair (traversal counters, stage events), untimed;
        # every rank runs them (they contain the collective); rank 0 reports its own stages
        inst = step(QZ_FLAG_COUNT_TRAVERSAL)
        timing = step(QZ_FLAG_STAGE_TIMING)
        for k in ("ms_closest", "ms_shadow", "ms_shade", "ms_other", "ms_total", "ms_sample"):
            inst[k] = timing[k]
        roofline, extra = stage_report(inst, WORKLOADS[workload][0], args.flags, clock_info.get("sm_mhz"))
        out["roofline"], out["extra"] = roofline, extra
        sc.close()
        del film
        torch.cuda.empty_cache()
        return out

    # every scene is built before anything is rendered: the build times are those of a fresh device (after a render has
    # released gigabytes of working memory, cudaMalloc alone takes hundreds of milliseconds)
    wanted = [args.workload] + ([w for w in ("obj_viewer", "mandelbrot") if w != args.workload] if world == 1 and not args.no_secondary else [])
    scenes = {}
    for wl in wanted:
        _, w0, h0, _, _ = WORKLOADS[wl]
        width, height = (args.width or w0, args.height or h0) if wl == args.workload else (w0, h0)
        sc = build_scene(qz, wl, width, height, args.mesh_triangles)
        scenes[wl] = (sc, qz.build_times(), width, height)

    head = measure(args.workload, args.spp, args.steps, args.warmup, with_e2e=True)

    secondary = None
    if world == 1 and not args.no_secondary:
        secondary = {}
        for wl in ("obj_viewer", "mandelbrot"):
            if wl == args.workload:
                continue
            m = measure(wl, 0, 3, 3, with_e2e=False)
            secondary[wl] = {"value": m["value"], "unit": "Mpaths/s", "mrays_per_s": m["mrays_per_s"], "ms_per_step": m["ms_per_step"], "steps": 3, "warmup": 3,
                             "config": {k: m[k] for k in ("scene", "width", "height", "spp", "max_bounces")},
                             "roofline": m["roofline"], **m["extra"],
                             "obj_parse_ms": m["build"].get("obj_parse_ms"), "scene_commit_ms": m["build"].get("commit_ms"),
                             "bvh_build_ms": m["build"].get("bvh_build_ms")}
            if wl == "obj_viewer":
                secondary[wl]["config"]["mesh_triangles"] = args.mesh_triangles
                secondary[wl]["mesh_first_hit_fraction"] = m["mesh_first_hit_fraction"]

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.workload, head["width"], head["height"], head["max_bounces"], args.mesh_triangles)

    if rank == 0:
        print(json.dumps({
            "metric": "Mpaths/s", "value": head["value"], "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "scene": head["scene"], "width": head["width"], "height": head["height"],
                       "spp": head["spp"], "spp_per_gpu": head["spp"] // world, "max_bounces": head["max_bounces"],
                       "parallelism": f"interleaved {head['strip_rows']}-row strips over {world} GPU(s), one NCCL sum-reduce of the film",
                       "l2": "per-step working set (path pool + result cells + sample memo, > 4 GB) exceeds L2"},
            "mrays_per_s": head["mrays_per_s"],
            "gpu_launches": head["gpu_launches"],
            "e2e": head["e2e"], "roofline": head["roofline"], "cpu_baseline": cpu, "clocks": head["clocks"],
            "bvh_build_ms": head["build"].get("bvh_build_ms"), "obj_parse_ms": head["build"].get("obj_parse_ms"),
            "secondary": secondary, **({"mesh_first_hit_fraction": head["mesh_first_hit_fraction"]} if head["mesh_first_hit_fraction"] is not None else {}),
            **head["extra"],
        }))
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="cornell_box")
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--spp", type=int, default=0, help="samples per pixel PER GPU (default: the workload's)")
    ap.add_argument("--pool", type=int, default=0, help="paths in flight (0 = library default)")
    ap.add_argument("--flags", type=int, default=0, help="extra QZ_FLAG_* bits for the timed steps (1 unsorted shading, 8 force BVH)")
    ap.add_argument("--mesh-triangles", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the obj_viewer / mandelbrot lines under `secondary`")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
