#!/usr/bin/env python3
"""BASELINE config 5: scripted flythrough.  The scheduler mirror (underwaterworld_b200/world.py) runs the
reference's window / frustum / priority rule every frame; whenever the recheck rule fires, ALL newly queued
chunks are handed to the GPU as one batch.  Reports p50 / p99 batch latency (host call -> host views valid,
i.e. uw_build incl. H2D, kernels, D2H) and batch sizes.  With --cpu also times the CPU oracle on the same
batches (1 thread, like the reference) for context.

    python tools/bench_flythrough.py [--frames-straight 600] [--frames-turn 600] [--cpu] [--out profiles/x.json]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def flythrough(builder, frames_straight=600, frames_turn=600, cpu=False):
    """Runs the scripted path against `builder` and returns the result record (bench.py calls this too)."""
    import underwaterworld_b200 as uw
    from underwaterworld_b200 import world as W, _ffi

    class _A:
        pass
    args = _A()
    args.frames_straight, args.frames_turn, args.cpu = frames_straight, frames_turn, cpu
    lib = uw.load_library()
    ctx = builder._ctx
    view = _ffi.UwBatchView()

    # pass 1: run the scheduler with real builds to get the exact batch sequence (world state depends on results)
    world = W.World()
    batches = []
    for frame, sub, cam in W.scripted_flythrough(args.frames_straight, args.frames_turn):
        world.remove_far_way(sub)
        if world.needs_recheck(sub):
            world.update_nearby(sub, cam)
            world.last_sub_pos, world.last_sub_bearing = sub.pos.copy(), sub.bearing().copy()
        if world.chunks_to_generate:
            pos = np.array(world.chunks_to_generate, dtype=np.int32)
            batches.append((frame, pos))
            world.build_batch(sub, builder)

    # pass 2: time the C-ABI call on every batch (3 repeats, keep the median per batch)
    lat_us, sizes, meshes = [], [], []
    for frame, pos in batches:
        ts = []
        for _ in range(3):
            h = C.c_void_p()
            t0 = time.perf_counter()
            st = lib.uw_build(ctx, pos.ctypes.data, len(pos), C.byref(h))
            lib.uw_batch_view_get(h, C.byref(view))
            dt = time.perf_counter() - t0
            assert st == 0
            nv, ni = view.n_verts, view.n_inds
            lib.uw_batch_free(h)
            ts.append(dt)
        lat_us.append(1e6 * float(np.median(ts)))
        sizes.append(len(pos))
        meshes.append((int(nv), int(ni)))
    lat = np.array(lat_us)
    res = {
        "workload": "BASELINE config 5: scripted flythrough, start (0,8,12), +x at 4 u/s, 60 Hz, "
                    f"{args.frames_straight} frames straight then yaw pi/6 rad/s for {args.frames_turn} frames",
        "frames": args.frames_straight + args.frames_turn, "batches": len(batches),
        "batch_chunks": {"first": sizes[0], "median_rest": float(np.median(sizes[1:])) if len(sizes) > 1 else None,
                         "min": int(min(sizes)), "max": int(max(sizes)), "total": int(sum(sizes))},
        "latency_us": {"p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99)),
                       "max": float(lat.max()), "first_batch": lat_us[0],
                       "p50_excluding_first": float(np.percentile(lat[1:], 50)) if len(lat) > 1 else None},
        "frame_budget_us_at_60hz": 16667,
        "note": "latency = uw_build(host positions) -> host-visible vertex/index views; 3 repeats per batch, median kept",
    }
    if args.cpu:
        from oracle import Oracle, MODE_FAITHFUL
        o = Oracle(12)
        perm = o.perm_table(0)
        cpu = [1e6 * o.build_batch_timed(perm, pos, MODE_FAITHFUL, 1)["seconds"] for _, pos in batches]
        res["cpu_oracle_1thread_latency_us"] = {"p50": float(np.percentile(cpu, 50)), "p99": float(np.percentile(cpu, 99)),
                                                "first_batch": cpu[0]}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames-straight", type=int, default=600)
    ap.add_argument("--frames-turn", type=int, default=600)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import underwaterworld_b200 as uw
    builder = uw.ChunkBuilder(uw.Perlin(0))
    res = flythrough(builder, args.frames_straight, args.frames_turn, args.cpu)
    line = json.dumps(res)
    print(line)
    if args.out:
        with open(args.out, "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
