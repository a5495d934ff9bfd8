import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
import underwaterworld_b200 as uw
# dirty the allocator's memory
xs = [torch.full((64 << 20,), 0xAB, dtype=torch.uint8, device="cuda") for _ in range(8)]
torch.cuda.synchronize(); del xs; torch.cuda.empty_cache()
pos = uw.region.box_region((-1, 1), (-1, 1), (-3, 2))
for kw in (dict(exact_f64=True, ordered=True), dict(), dict(staged=True)):
    with uw.ChunkBuilder(uw.Perlin(0), tris=True, **kw) as b:
        for rep in range(2):
            batch = b.build(pos)
            bad = []
            for i in range(len(pos)):
                t = batch.tri_cell_start[i].astype(np.int64)
                ok = (np.diff(t) >= 0).all() and t[-1] == batch.descs["index_count"][i] // 3 and t[0] == 0
                if not ok: bad.append((i, tuple(pos[i]), int(batch.descs["flags"][i]), int(batch.descs["index_count"][i]), t[:6].tolist()))
            print(kw, "rep", rep, "bad", bad[:5])
