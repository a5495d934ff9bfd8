import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
import underwaterworld_b200 as uw
pos = uw.region.box_region((-2, 2), (-2, 2), (-3, 2))      # 80 chunks
ref = None
for kw in (dict(ordered=True), dict(), dict(staged=True), dict(tris=True), dict(exact_f64=True, ordered=True)):
    with uw.ChunkBuilder(uw.Perlin(0), **kw) as b:
        batch = b.build(pos)
        per = [(batch.chunk(i).inds.tobytes(), batch.chunk(i).flags) for i in range(len(pos))]
        if ref is None: ref = per
        print(kw, "n_verts", batch.n_verts, "n_inds", batch.n_inds, "same topology:", per == ref)
with uw.ChunkBuilder(uw.Perlin(0), internal_size=10) as b:
    print("S=10", b.build(pos[:20]).n_inds)
with uw.ChunkBuilder(uw.Perlin(0), internal_size=7) as b:      # generic runtime-table kernels
    print("S=7", b.build(pos[:20]).n_inds)
with uw.ChunkBuilder(uw.Perlin(0), internal_size=64) as b:
    print("S=64", b.build(pos[30:33]).n_inds)
with uw.ChunkBuilder(uw.Perlin(0), internal_size=20) as b:
    print("S=20", b.build(pos[30:36]).n_inds)
print("done")
