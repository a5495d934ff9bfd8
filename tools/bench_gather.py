#!/usr/bin/env python3
"""Optional NVLink gather of finished meshes to the rendering GPU, timed separately from the
build.  Launch:  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_gather.py
Each rank builds its x-slab of a region device-resident; rank 0 then receives every rank's packed mesh over
NCCL point-to-point.  Prints one JSON line on rank 0 (build ms, gather ms, GB/s, and a correctness check of the
received bytes against a rebuild of that slab on rank 0)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import underwaterworld_b200 as uw
    from underwaterworld_b200.gather import device_batch_tensors, gather_meshes, as_numpy_batch

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    box = ((-32, 32), (-32, 32), (-4, 4))                                 # 32768 chunks over all ranks
    pos = uw.region.shard_region(*box, rank, world)
    b = uw.ChunkBuilder(uw.Perlin(0), device=local, ordered=True)          # request order: byte-comparable
    stream = torch.cuda.current_stream()
    b.set_stream(stream.cuda_stream)
    d_pos = torch.from_numpy(pos).cuda()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for it in range(3):
        dist.barrier(); torch.cuda.synchronize()
        ev[0].record(stream)
        b.build_device(d_pos.data_ptr(), len(pos))
        b.sync()
        ev[1].record(stream)
        descs, verts, inds = device_batch_tensors(b)
        ev[2].record(stream)
        got = gather_meshes(descs, verts, inds, dst=0)
        ev[3].record(stream)
        torch.cuda.synchronize()
    build_ms, gather_ms = ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])
    t = torch.tensor([build_ms, gather_ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        recv_bytes = sum(sum(x.numel() for x in g) for r, g in enumerate(got) if r != 0)
        ok = True
        if world > 1:                                                      # rebuild rank 1's slab here and compare bytes
            p1 = uw.region.shard_region(*box, 1, world)
            ref = b.build(p1)
            d, v, i = as_numpy_batch(*got[1])
            ok = bool(np.array_equal(d, ref.descs) and np.array_equal(v.view(np.uint8), ref.verts.view(np.uint8)) and np.array_equal(i, ref.inds))
        print(json.dumps({"world": world, "chunks_total": 32768, "build_ms_max": float(t[0]), "gather_ms_max": float(t[1]),
                          "gathered_MB": recv_bytes / 1e6, "gather_GBps": recv_bytes / (float(t[1]) / 1e3) / 1e9 if world > 1 else None,
                          "rank1_bytes_match_rebuild": ok}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
