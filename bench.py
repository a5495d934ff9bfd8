#!/usr/bin/env python3
"""bench.py -- chunk-build throughput (Perlin density + marching-cubes mesh) on B200.

Contract (see the task statement):  python bench.py --gpus N --steps K --warmup W
prints ONE JSON line on rank 0.  A "step" is one pass of the whole hot path (K1 noise ->
K2 classify -> K3 scan -> K4 emit, one fused kernel) over one region of chunk positions.

Workload at EVERY N: BASELINE.json configs[2], the configuration its metric ("... at 1/2/4/8 B200") is quoted on:
the 128x128x32 region = 524 288 chunks of 12^3 cells, cut into N contiguous x-slabs (strong scaling; one process
per GPU; no collective on the compute path -- chunks are independent, chunk.rs:89-129).

  value     voxels/s (cells/s), device-resident: every rank's slab positions are already in its HBM, outputs stay in
            its own HBM; CUDA events on the launching stream around every step's launch, L2 flushed between steps,
            MAX over ranks.
  e2e       the same metric through the C ABI from HOST positions to meshes the renderer can draw: pinned H2D of
            every rank's slab, the fused kernels writing vertices / indices / descriptors straight into the
            RENDERING GPU's arenas (rank 0; NVLink peer stores from the other ranks -- uw_gather_*), the head flags,
            and the D2H of the draw list (descriptors of the meshed chunks) + heads on rank 0.  The reference's build_full ends the same
            way: in GPU buffers (wgpu, chunk.rs:291-305) plus host-side counts.  Host wall clock per step, barrier
            before every step, MAX over ranks.
  e2e_host  beside it: every rank streams its slab's full meshes into pinned HOST memory (uw_build_async /
            uw_batch_wait, slices of 8192 chunks, two in flight) -- PCIe-bound; `d2h_ceiling_gbs` is a bare pinned
            D2H copy of the same bytes by all ranks at once, measured in this run.
  configs   (N = 1 only) BASELINE configs[1] (2048 chunks: burst latency, pipelined host path, per-stage rooflines
            of the staged pipeline at 32 768 chunks), configs[3] (64^3) and configs[4] (flythrough).

--impl reference  times the CPU oracle (oracle/, faithful mode = the reference's algorithm, all host threads) on a
bounded sample of the same region.  The reference is Rust and cannot be compiled in this image -- see DESIGN.md.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

S = 12
L3 = (S + 1) ** 3
CELLS = S ** 3
SEED = 0
FLOP_PER_SAMPLE = 285          # algorithmic op count of the reference's density function
FP32_NOMINAL_TFLOPS = 74.4     # 148 SM x 128 lanes x 2 x 1.965 GHz
ISSUE_SLOTS_PER_CYCLE = 148 * 4

REGION = ((-64, 64), (-64, 64), (-16, 16))      # BASELINE configs[2]
WORKLOAD = ("large region 128x128x32 = 524288 chunks of 12^3 (BASELINE configs[2]) cut into N contiguous x-slabs, "
            "seed 0, octaves 3, iso -0.1")
HOST_SLICE = 8192


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def tracked_profile(name):
    """A tracked ncu summary under profiles/ (JSON written by tools/ncu_json.py); None if absent."""
    p = os.path.join(ROOT, "profiles", name)
    try:
        return json.load(open(p))
    except Exception:
        return None


class ClockSampler:
    """Samples SM clock / throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.004)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def region_positions() -> np.ndarray:
    from underwaterworld_b200 import region
    return region.box_region(*REGION)


# ------------------------------------------------------------------------------------------------
# CPU oracle timing (cpu_baseline leg and --impl reference)
# ------------------------------------------------------------------------------------------------
def cpu_sample(x_columns: int) -> np.ndarray:
    """Bounded sample of the region that keeps its composition: `x_columns` full x-planes (all 128 y, all 32 z
    layers) from the middle of the region -- 4096 chunks each, 5 of the 32 z layers can hold surface, as in the
    whole region."""
    from underwaterworld_b200 import region
    x0 = -(x_columns // 2)
    return region.box_region((x0, x0 + x_columns), REGION[1], REGION[2])


def time_cpu(pos: np.ndarray, threads: int, repeats: int = 1):
    from oracle import Oracle, MODE_FAITHFUL
    o = Oracle(S)
    perm = o.perm_table(SEED)
    best = None
    for _ in range(repeats):
        r = o.build_batch_timed(perm, pos, MODE_FAITHFUL, threads)
        if best is None or r["seconds"] < best["seconds"]:
            best = r
    return best


def base_line(args, world):
    return {
        "metric": "voxels/s (Perlin + MC mesh build)", "unit": "voxels/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "data": "synthetic",
        "config": {"workload": WORKLOAD, "region_chunks": 524288, "cells_per_chunk": CELLS, "samples_per_chunk": L3},
    }


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    # calibrate on one x-plane, then size the per-step sample so the whole run stays within ~2 minutes
    cal = time_cpu(cpu_sample(1), threads)
    rate = 4096 / max(cal["seconds"], 1e-6)
    budget_s = 100.0
    cols = int(max(1, min(128, rate * budget_s / max(1, args.steps + args.warmup) / 4096)))
    sample = cpu_sample(cols)
    for _ in range(args.warmup):
        time_cpu(sample, threads)
    t = 0.0
    for _ in range(args.steps):
        t += time_cpu(sample, threads)["seconds"]
    ms = 1e3 * t / args.steps
    value = len(sample) * CELLS / (ms / 1e3)
    line = base_line(args, args.gpus)
    line.update({
        "impl": "reference", "value": value, "chunks_per_s": len(sample) / (ms / 1e3), "ms_per_step": ms,
        "dtype": "f64 noise / f32 mesh",
        "cpu_baseline": {"value": value, "unit": "voxels/s", "cores": threads, "kind": "port",
                         "sample": f"{cols} of the region's 128 x-planes per step ({len(sample)} chunks: every y, every z layer), "
                                   "oracle faithful mode (linear-search dedup, per-index colour, per-cell Tri map), -O3"},
        "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is Rust; no cargo/rustc in this image -> CPU restatement (oracle/) timed instead",
    })
    line["config"]["sample_chunks_per_step"] = len(sample)
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU fallback")
    # One process per GPU: run (and first-touch pinned host memory) on the CPUs next to this rank's GPU, as a
    # production launcher would -- on a two-socket box a remote pinned arena halves the D2H rate of the host path.
    numa = "unpinned (single rank: the CPU baseline of this run uses every host core)"
    try:
        if world == 1 or os.environ.get("UW_BENCH_NO_PIN"):
            raise RuntimeError("single rank")
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
        numa = f"nvmlDeviceSetCpuAffinity(gpu {local}): {len(os.sched_getaffinity(0))} cpus"
    except Exception as e:                                   # not fatal: the numbers are simply measured unpinned
        if world > 1:
            numa = f"unpinned ({type(e).__name__})"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(x, op):
        if world == 1:
            return x
        t = torch.tensor(np.atleast_1d(np.asarray(x, dtype=np.float64)), device="cuda")
        dist.all_reduce(t, op=op)
        r = t.cpu().numpy()
        return float(r[0]) if np.ndim(x) == 0 else r

    def rmax(x):
        return reduce(x, dist.ReduceOp.MAX) if world > 1 else x

    def rsum(x):
        return reduce(x, dist.ReduceOp.SUM) if world > 1 else x

    import underwaterworld_b200 as uw
    from underwaterworld_b200 import _ffi, gather as G
    lib = uw.load_library()

    whole = region_positions()
    n_total = len(whole)
    first, n = G.slab_bounds(n_total, world, rank)
    # the request lives in PINNED host memory (the contract's "host->device copy from pinned host memory"): the library
    # then copies it to the device straight from this buffer
    whole_pin = torch.from_numpy(whole).pin_memory()           # every rank keeps the request; its slab is a slice of it
    del whole
    pos = whole_pin.numpy()[first:first + n]
    builder = uw.ChunkBuilder(uw.Perlin(SEED), internal_size=S, device=local)
    stream = torch.cuda.current_stream()
    builder.set_stream(stream.cuda_stream)
    d_pos = torch.from_numpy(pos).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")    # > 126 MB L2

    K, W = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(local)

    # ---- value: device-resident, every rank its slab --------------------------------------------------
    for i in range(W):
        flush.fill_(i & 0xFF)
        builder.build_device(d_pos.data_ptr(), n)
    builder.sync()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    barrier()
    with sampler:
        for i in range(K):
            flush.fill_(i & 0xFF)
            ev0[i].record(stream)
            builder.build_device(d_pos.data_ptr(), n)
            ev1[i].record(stream)
        builder.sync()
        barrier()
        step_ms = [a.elapsed_time(b) for a, b in zip(ev0, ev1)]
        launches_per_step = builder.stage_times()["launches"]
        dv = builder.device_view()
        my_verts, my_inds = int(dv.n_verts), int(dv.n_inds)

        # ---- e2e: host positions -> meshes in the rendering GPU's arenas (rank 0) + draw list on the host ----------
        # Gather-aware partition: the rendering rank's own output does not cross NVLink, so when the gather is bound by
        # its NVLink ingress (N = 8) it takes a larger slab.  The share is chosen during the warm-up steps by measuring
        # whole steps under a handful of candidate shares (RegionGather.tune: one small all_reduce between steps, never
        # inside a timed step) and is frozen for the timed ones.  Segments are sized for half the region each.
        half = (n_total + 1) // 2
        rg = G.RegionGather(builder, rank, world, n_total, dst=0, seg_vcap=half * 192 + 4096, seg_icap=half * 640 + 16384,
                            bcast_device="cuda" if world > 1 else None)
        res = None
        step_no = [0]

        def e2e_step():
            nonlocal res
            gf, gn = rg.plan(n_total)
            gpos = whole_pin.numpy()[gf:gf + gn]
            flush.fill_(step_no[0] & 0xFF)
            step_no[0] += 1
            barrier()
            t0 = time.perf_counter()
            rg.build(gpos, gf)
            if rank == 0:
                res = rg.wait(draw_to_host=True)
            builder.sync()
            return time.perf_counter() - t0

        if world > 1:
            # measured search over a handful of render shares (two cold steps, then two steps per candidate), warm-up only
            rg.tune(e2e_step, device="cuda")
        for i in range(W):
            e2e_step()
        gf, gn = rg.plan(n_total)
        gpos = whole_pin.numpy()[gf:gf + gn]
        e2e_t = []
        for i in range(K):
            flush.fill_(i & 0xFF)
            barrier()
            t0 = time.perf_counter()
            rg.build(gpos, gf)
            if rank == 0:
                res = rg.wait(draw_to_host=True)
            builder.sync()
            e2e_t.append(time.perf_counter() - t0)
        barrier()
        e2e_steps = rmax(np.array(e2e_t))                    # per step: the slowest rank (rank 0 waits for all)
        render_share = rg.render_share
        if rank == 0:
            g_verts, g_inds = int(res.n_verts), int(res.n_inds)
            g_mesh = sum(s["n_mesh"] for s in res.segments)
            g_blank = sum(s["n_blank"] for s in res.segments)
            g_guard = sum(s["guard"] for s in res.segments)
            g_counts = [int(s["n_chunks"]) for s in res.segments]
            assert res.n_chunks == n_total
        rg.close()

        # ---- e2e_host: every rank streams its slab's meshes into pinned host memory ----------------------------
        ctx = builder._ctx
        view = _ffi.UwBatchView()
        slices = [pos[a:a + HOST_SLICE] for a in range(0, n, HOST_SLICE)]

        def host_pass():
            """uw_build_async(k+1) before uw_batch_wait(k): batch k's D2H runs under batch k+1's kernel."""
            nv = ni = 0
            prev = None
            for sl in slices:
                h = C.c_void_p()
                if lib.uw_build_async(ctx, sl.ctypes.data, len(sl), C.byref(h)) != 0:
                    raise RuntimeError(lib.uw_last_error(ctx).decode())
                if prev is not None:
                    if lib.uw_batch_wait(prev) != 0:
                        raise RuntimeError(lib.uw_last_error(ctx).decode())
                    lib.uw_batch_view_get(prev, C.byref(view)); nv += view.n_verts; ni += view.n_inds
                    lib.uw_batch_free(prev)
                prev = h
            if lib.uw_batch_wait(prev) != 0:
                raise RuntimeError(lib.uw_last_error(ctx).decode())
            lib.uw_batch_view_get(prev, C.byref(view)); nv += view.n_verts; ni += view.n_inds
            lib.uw_batch_free(prev)
            return nv, ni

        KH = max(3, min(K, 10))
        for _ in range(2):
            host_pass()
        host_t = []
        hv = hi = 0
        for _ in range(KH):
            barrier()
            t0 = time.perf_counter()
            hv, hi = host_pass()
            host_t.append(time.perf_counter() - t0)
        barrier()
        host_steps = rmax(np.array(host_t))
        host_bytes = n * 32 + hv * 24 + hi * 2

        # bare pinned D2H of the same bytes by all ranks at once: the ceiling of the host path on this box
        pin = torch.empty(max(host_bytes, 1 << 20), dtype=torch.uint8, pin_memory=True)
        devbuf = torch.empty_like(pin, device="cuda")
        d2h_t = []
        for i in range(5):
            barrier()
            t0 = time.perf_counter()
            pin.copy_(devbuf, non_blocking=True)
            torch.cuda.synchronize()
            d2h_t.append(time.perf_counter() - t0)
        barrier()
        d2h_steps = rmax(np.array(d2h_t[1:]))
        del pin, devbuf

    ms_per_step = rmax(sum(step_ms) / K)
    tot_verts, tot_inds = int(rsum(my_verts)), int(rsum(my_inds))
    tot_host_bytes = rsum(float(host_bytes))
    e2e_ms = 1e3 * float(np.mean(e2e_steps))
    host_ms = 1e3 * float(np.median(host_steps))
    d2h_ms = 1e3 * float(np.median(d2h_steps))
    value = n_total * CELLS / (ms_per_step / 1e3)

    # ---- sustained leg (N = 1): >= 2 s of back-to-back launches, clocks sampled ------------------------------
    sustained = None
    if rank == 0 and world == 1 and not args.quick:
        s2 = ClockSampler(local)
        reps = int(max(50, 2200.0 / max(ms_per_step, 0.05)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with s2:
            e0.record(stream)
            for _ in range(reps):
                builder.build_device(d_pos.data_ptr(), n)
            e1.record(stream)
            builder.sync()
        sms = e0.elapsed_time(e1) / reps
        sustained = {"launches": reps, "seconds": e0.elapsed_time(e1) / 1e3, "ms_per_step": sms,
                     "voxels_per_s": n_total * CELLS / (sms / 1e3), "clocks": s2.summary(),
                     "note": "back-to-back launches of the same region, no L2 flush in between, CUDA events around the whole run"}

    # ---- roofline of the dominant kernel, from the SAME timing as ms_per_step ------------------------------
    peak, peak_src = measured_peaks()
    ffma_tflops = builder.ffma_peak_tflops() if rank == 0 else 0.0
    slab_flops = n * L3 * FLOP_PER_SAMPLE                        # this launch's algorithmic work (rank 0's slab)
    my_ms = sum(step_ms) / K
    out_bytes = n * 12 + n * 32 + 24 * my_verts + 2 * my_inds
    prof = tracked_profile("r02_ncu_fused_config3.json")        # ncu --set full of this kernel on the N = 1 launch
    roofline = None
    if rank == 0:
        ach = slab_flops / (my_ms / 1e3) / 1e12
        roofline = {
            "kernel": "k_build_fused<12,3,u16> (noise + classify + scan + emit in one persistent kernel; one launch per step)",
            "bound": "fp32", "achieved": ach, "peak": FP32_NOMINAL_TFLOPS, "unit": "TFLOP/s", "frac": ach / FP32_NOMINAL_TFLOPS,
            "algorithmic_flop_per_launch": slab_flops, "avg_launch_ms": my_ms,
            "timing": "CUDA events around each of the K timed launches on the launching stream (the same events ms_per_step is "
                      "computed from; this rank's mean)",
            "peak_source": "nominal FP32 FMA peak, 148 SMs x 128 lanes x 2 x 1.965 GHz (MEASURED_PEAKS.json holds HBM and bf16 only); "
                           "the FFMA-chain rate measured in this run is measured_ffma_peak_tflops",
            "measured_ffma_peak_tflops": ffma_tflops,
            "frac_of_measured_ffma": ach / ffma_tflops if ffma_tflops > 0 else None,
            "traffic": (prof or {}).get("dram_bytes") if world == 1 else None,
            "traffic_source": "profiles/r02_ncu_fused_config3.json: dram__bytes_read.sum + dram__bytes_write.sum of one launch "
                              "(ncu --set full, N = 1 region launch)" if (prof and world == 1) else None,
            "note": "achieved = SURVEY 8d's algorithmic count (285 FLOP per density sample x 2197 samples x chunks of this launch) over "
                    "the WHOLE fused kernel (extraction included).  The tensor-product factorisation executes several times fewer "
                    "instructions than that count, so this can exceed 1; `hardware_view` is the honest utilisation figure.",
            "hbm_view": {"achieved": out_bytes / (my_ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": out_bytes / (my_ms / 1e3) / 1e9 / peak, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": out_bytes,
                         "note": "densities never leave the SM: compulsory HBM traffic is positions in and mesh out only"},
        }
        if prof and world == 1 and prof.get("warp_inst"):
            clk = (sampler.summary().get("sm_mhz") or 1965.0) * 1e6
            roofline["hardware_view"] = {
                "warp_instructions_per_launch": prof["warp_inst"],
                "issue_slot_frac": prof["warp_inst"] / (ISSUE_SLOTS_PER_CYCLE * clk * my_ms / 1e3),
                "ncu_issue_active_pct": prof.get("issue_active_pct"), "ncu_fma_pipe_pct": prof.get("fma_pipe_pct"),
                "source": "instruction count from profiles/r02_ncu_fused_config3.json over this run's launch time and sampled SM clock; "
                          "148 SMs x 4 issue slots per cycle"}

    # ---- the other BASELINE configs, rank 0 at N = 1 -----------------------------------------------------------
    configs = None
    if rank == 0 and world == 1 and not args.quick:
        configs = other_configs(uw, torch, builder, stream, flush, local, peak, ffma_tflops, K, W)

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same region -----------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cols = 2
        sample = cpu_sample(cols)
        allc = time_cpu(sample, threads)
        if allc["seconds"] < 5.0:                   # fast host: take a bigger bite (10-30 s of CPU work in total)
            cols = int(min(16, max(2, 12.0 / max(allc["seconds"], 1e-3))))
            sample = cpu_sample(cols)
            allc = time_cpu(sample, threads)
        one = time_cpu(cpu_sample(1), 1)            # the reference is single-threaded (README.md:19)
        cpu_baseline = {
            "value": len(sample) * CELLS / allc["seconds"], "unit": "voxels/s", "cores": threads, "kind": "port",
            "sample": f"{cols} of the region's 128 x-planes ({len(sample)} chunks: every y, every z layer), once; oracle faithful "
                      "mode (linear-search dedup, per-index colour, per-cell Tri map), g++ -O3",
            "chunks_per_s": len(sample) / allc["seconds"],
            "single_thread": {"value": 4096 * CELLS / one["seconds"], "chunks_per_s": 4096 / one["seconds"], "cores": 1,
                              "sample": "1 x-plane (4096 chunks)"},
        }

    if rank == 0:
        h2d = n_total * 12
        d2h = g_mesh * 32 + 64 * world + 4
        line = base_line(args, world)
        line["steps"], line["warmup"] = K, W
        line.update({
            "value": value, "chunks_per_s": n_total / (ms_per_step / 1e3), "ms_per_step": ms_per_step,
            "dtype": "f32 noise + f64 guard band / f32 mesh / u16 indices",
            "nontrivial": {"chunks_with_mesh": g_mesh, "chunks_with_mesh_per_s": g_mesh / (ms_per_step / 1e3),
                           "voxels_per_s": g_mesh * CELLS / (ms_per_step / 1e3),
                           "note": "same time, counting only chunks that end with a mesh"},
            "e2e": {"value": n_total * CELLS / (e2e_ms / 1e3), "unit": "voxels/s", "chunks_per_s": n_total / (e2e_ms / 1e3),
                    "ms_per_step": e2e_ms, "ms_per_step_median": 1e3 * float(np.median(e2e_steps)),
                    "path": "host positions -> pinned H2D per rank -> fused kernels store meshes + descriptors straight into rank 0's "
                            "arenas (NVLink peer stores from ranks > 0; uw_gather_build) -> head flags -> rank 0: uw_gather_wait + D2H "
                            "of the draw list (descriptors of the chunks that ended with a mesh) and the heads into pinned host memory",
                    "timing": "host perf_counter per step on every rank (barrier, build, wait/sync), MAX over ranks per step, mean over K",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "render_share": render_share if render_share > 0 else 1.0 / world, "chunks_per_rank": g_counts,
                    "share_search_ms": [[round(a, 4), round(1e3 * b, 4)] for a, b in getattr(rg, "tune_trace", [])],
                    "partition": "gather-aware: the rendering rank takes a larger slab (its output does not cross NVLink); the share is "
                                 "the fastest of a handful of candidates measured over the warm-up steps (RegionGather.tune), "
                                 "frozen for the timed ones; value uses the even split",
                    "mesh_bytes_to_render_gpu": 24 * g_verts + 2 * g_inds,
                    "nvlink_bytes_per_step": (24 * (g_verts - int(res.segments[0]["n_verts"])) + 2 * (g_inds - int(res.segments[0]["n_inds"]))
                                              + 32 * (n_total - int(res.segments[0]["n_chunks"]))) if world > 1 else 0},
            "e2e_host": {"value": n_total * CELLS / (host_ms / 1e3), "unit": "voxels/s", "chunks_per_s": n_total / (host_ms / 1e3),
                         "ms_per_step": host_ms, "d2h_bytes_per_step": tot_host_bytes, "h2d_bytes_per_step": h2d,
                         "achieved_d2h_gbs": tot_host_bytes / (host_ms / 1e3) / 1e9,
                         "d2h_ceiling_gbs": tot_host_bytes / (d2h_ms / 1e3) / 1e9,
                         "frac_of_d2h_ceiling": d2h_ms / host_ms,
                         "path": f"every rank: uw_build_async / uw_batch_wait over slices of {HOST_SLICE} chunks, two in flight; full meshes "
                                 "(descriptors + vertices + indices) into pinned host memory",
                         "timing": f"host perf_counter around one pass over the rank's slab, barrier before, MAX over ranks, median of {KH} passes; "
                                   "d2h_ceiling = bare pinned cudaMemcpy of the same bytes by all ranks simultaneously, same run"},
            "gpu_launches": int(launches_per_step) * K,
            "kernels_per_step": ["k_build_fused"],
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "sustained": sustained,
            "configs": configs,
            "clocks": sampler.summary(),
            "mesh": {"n_verts": tot_verts, "n_inds": tot_inds, "chunks_with_mesh": g_mesh, "chunks_blank_early": g_blank,
                     "guard_band_reevals": int(g_guard)},
        })
        line["config"].update({"chunks_per_gpu": n, "parallelism": f"x-slabs x{world}", "host_affinity": numa,
                               "l2": "flushed between timed steps (256 MB write)"})
        emit(line)
    builder.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def other_configs(uw, torch, builder, stream, flush, local, peak, ffma_tflops, K, W):
    """BASELINE configs[1], [3], [4] and the north_star's per-stage targets (rank 0, N = 1)."""
    from underwaterworld_b200 import _ffi, region
    lib = uw.load_library()
    out = {}

    # ---- configs[1]: 2048 chunks, one batched launch (burst latency) ----------------------------------------------
    pos = region.box_region((-8, 8), (-8, 8), (-4, 4))
    n = len(pos)
    d_pos = torch.from_numpy(pos).cuda()
    for i in range(W):
        flush.fill_(i & 0xFF)
        builder.build_device(d_pos.data_ptr(), n)
    builder.sync()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    for i in range(K):
        flush.fill_(i & 0xFF)
        ev0[i].record(stream)
        builder.build_device(d_pos.data_ptr(), n)
        ev1[i].record(stream)
    builder.sync()
    ms = sum(a.elapsed_time(b) for a, b in zip(ev0, ev1)) / K
    dv = builder.device_view()
    nv, ni = int(dv.n_verts), int(dv.n_inds)
    ctx = builder._ctx
    view = _ffi.UwBatchView()

    def pipelined(steps):
        prev = C.c_void_p()
        if lib.uw_build_async(ctx, pos.ctypes.data, n, C.byref(prev)) != 0:
            raise RuntimeError(lib.uw_last_error(ctx).decode())
        for _ in range(1, steps):
            nxt = C.c_void_p()
            if lib.uw_build_async(ctx, pos.ctypes.data, n, C.byref(nxt)) != 0:
                raise RuntimeError(lib.uw_last_error(ctx).decode())
            if lib.uw_batch_wait(prev) != 0:
                raise RuntimeError(lib.uw_last_error(ctx).decode())
            lib.uw_batch_free(prev)
            prev = nxt
        if lib.uw_batch_wait(prev) != 0:
            raise RuntimeError(lib.uw_last_error(ctx).decode())
        lib.uw_batch_view_get(prev, C.byref(view))
        lib.uw_batch_free(prev)

    pipelined(max(W, 3))
    pt = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pipelined(K)
        pt.append(time.perf_counter() - t0)
    pipe_ms = 1e3 * float(np.median(pt)) / K
    single = []
    for i in range(K):
        flush.fill_(i & 0xFF)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h = C.c_void_p()
        lib.uw_build(ctx, pos.ctypes.data, n, C.byref(h))
        lib.uw_batch_free(h)
        single.append(time.perf_counter() - t0)
    flops = n * L3 * FLOP_PER_SAMPLE
    out["config2_spawn_2048"] = {
        "workload": "BASELINE configs[1]: 16x16x8 = 2048 chunks of 12^3 around the sub start, one batched launch",
        "ms_per_step": ms, "voxels_per_s": n * CELLS / (ms / 1e3), "chunks_per_s": n / (ms / 1e3),
        "timing": "CUDA events around each launch, L2 flushed between steps, mean of K",
        "roofline_fp32": {"achieved_tflops": flops / (ms / 1e3) / 1e12, "frac_of_nominal": flops / (ms / 1e3) / 1e12 / FP32_NOMINAL_TFLOPS,
                          "note": "latency / tail-bound at 3.5 chunks per resident CTA"},
        "e2e_host_pipelined": {"ms_per_step": pipe_ms, "chunks_per_s": n / (pipe_ms / 1e3), "voxels_per_s": n * CELLS / (pipe_ms / 1e3),
                               "d2h_bytes_per_step": n * 32 + nv * 24 + ni * 2, "h2d_bytes_per_step": n * 12},
        "e2e_host_single_call_ms": 1e3 * float(np.median(single)),
        "mesh": {"n_verts": nv, "n_inds": ni},
    }

    # ---- north_star per-stage targets: staged pipeline at 32 768 chunks -------------------------------------------
    big = region.box_region((-32, 32), (-32, 32), (-4, 4))
    d_big = torch.from_numpy(big).cuda()
    stg = uw.ChunkBuilder(uw.Perlin(SEED), internal_size=S, device=local, staged=True)
    stg.set_stream(stream.cuda_stream)
    stg.set_profiling(True)
    t_big = {"noise_ms": 0.0, "classify_ms": 0.0, "scan_ms": 0.0, "emit_ms": 0.0}
    for i in range(5):
        stg.build_device(d_big.data_ptr(), len(big)); stg.sync()
        if i >= 2:
            tt = stg.stage_times()
            for k in t_big:
                t_big[k] += tt[k] / 3
    vb = stg.device_view()
    stg.close()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(2):
        builder.build_device(d_big.data_ptr(), len(big)); builder.sync()
    e0.record(stream)
    for i in range(5):
        builder.build_device(d_big.data_ptr(), len(big))
    e1.record(stream); builder.sync()
    fused_big_ms = e0.elapsed_time(e1) / 5
    ext_bytes = len(big) * (4 * L3 + 44) + 24 * int(vb.n_verts) + 2 * int(vb.n_inds)
    ext_ms = t_big["classify_ms"] + t_big["scan_ms"] + t_big["emit_ms"]
    nprof = tracked_profile("r02_ncu_noise_32768.json")
    noise_tf = len(big) * L3 * FLOP_PER_SAMPLE / (t_big["noise_ms"] / 1e3) / 1e12
    out["north_star_stages_32768"] = {
        "batch_chunks": len(big),
        "noise_stage": {"kernel": "k_noise_spec<12,3> (staged pipeline)", "ms": t_big["noise_ms"],
                        "algorithmic_tflops": noise_tf, "frac_of_nominal_fp32_peak": noise_tf / FP32_NOMINAL_TFLOPS,
                        "frac_of_measured_ffma_peak": noise_tf / ffma_tflops if ffma_tflops else None,
                        "hardware_view": None if not nprof else {
                            "warp_instructions": nprof.get("warp_inst"),
                            "issue_slot_frac": nprof["warp_inst"] / (ISSUE_SLOTS_PER_CYCLE * 1.965e9 * t_big["noise_ms"] / 1e3),
                            "ncu_issue_active_pct": nprof.get("issue_active_pct"), "ncu_fma_pipe_pct": nprof.get("fma_pipe_pct"),
                            "ncu_lsu_pipe_pct": nprof.get("lsu_pipe_pct"),
                            "source": "profiles/r02_ncu_noise_32768.json (ncu --set full of this kernel at this batch size)"},
                        "density_write_gbs": len(big) * 4 * L3 / (t_big["noise_ms"] / 1e3) / 1e9},
        "extraction_stages": {"kernels": "k_classify_spec<12> + k_scan_chunks + k_emit_small<12,u16> (staged pipeline)", "ms": ext_ms,
                              "stages_ms": {"classify": t_big["classify_ms"], "scan": t_big["scan_ms"], "emit": t_big["emit_ms"]},
                              "algorithmic_bytes": ext_bytes, "algorithmic_gbs": ext_bytes / (ext_ms / 1e3) / 1e9,
                              "frac_of_measured_hbm": ext_bytes / (ext_ms / 1e3) / 1e9 / peak,
                              "classify_frac_of_measured_hbm": len(big) * (4 * L3 + 16) / (t_big["classify_ms"] / 1e3) / 1e9 / peak},
        "fused_kernel": {"ms": fused_big_ms, "chunks_per_s": len(big) / (fused_big_ms / 1e3),
                         "voxels_per_s": len(big) * CELLS / (fused_big_ms / 1e3)},
        "note": "targets: >= 60 % of peak FP32 in the noise stage (algorithmic FLOPs, SURVEY 8d), >= 50 % of peak HBM in the extraction stages",
    }
    del d_big

    # ---- configs[3]: the 2048-chunk region at 64^3 cells per chunk ------------------------------------------------
    try:
        b64 = uw.ChunkBuilder(uw.Perlin(SEED), internal_size=64, device=local)
        b64.set_stream(stream.cuda_stream)
        for i in range(2):
            b64.build_device(d_pos.data_ptr(), n); b64.sync()
        e0.record(stream)
        for i in range(3):
            b64.build_device(d_pos.data_ptr(), n)
        e1.record(stream); b64.sync()
        ms64 = e0.elapsed_time(e1) / 3
        b64.set_profiling(True)
        b64.build_device(d_pos.data_ptr(), n); b64.sync()
        st64 = b64.stage_times()
        v64 = b64.device_view()
        out["config4_highres_64"] = {
            "workload": "BASELINE configs[3]: the 16x16x8 region at 64^3 cells per chunk (u32 indices)",
            "ms_per_step": ms64, "voxels_per_s": n * 64 ** 3 / (ms64 / 1e3), "chunks_per_s": n / (ms64 / 1e3),
            "stages_ms": {k: st64[k] for k in ("noise_ms", "classify_ms", "scan_ms", "emit_ms")},
            "mesh": {"n_verts": int(v64.n_verts), "n_inds": int(v64.n_inds)},
            "hbm_frac_end_to_end": (n * 65 ** 3 * 4 * 3 + 24 * int(v64.n_verts) + 4 * int(v64.n_inds)) / (ms64 / 1e3) / 1e9 / peak,
        }
        b64.close()
    except Exception as e:                                  # pragma: no cover
        out["config4_highres_64"] = {"error": f"{type(e).__name__}: {e}"}

    # ---- configs[4]: scripted flythrough, per-frame batches -------------------------------------------------------
    try:
        from bench_flythrough import flythrough
        out["config5_flythrough"] = flythrough(builder)
    except Exception as e:                                  # pragma: no cover
        out["config5_flythrough"] = {"error": f"{type(e).__name__}: {e}"}
    return out


_JSON_FD = None


def emit(line):
    """The ONE JSON line of the contract, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    # Libraries print to stdout behind Python's back (NCCL's version banner under torchrun, for one): keep the
    # real stdout for the JSON line and point fd 1 at stderr for everything else.
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="skip the other configs and the sustained leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
