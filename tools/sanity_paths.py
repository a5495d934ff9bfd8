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
# paths that need larger batches: cost-ordered hand-out (n > resident CTAs), multi-tile chunk scan (n > 1024),
# staged S=10 (k_classify_spec<10>), and two host batches in flight
big = uw.region.box_region((-6, 6), (-6, 6), (-4, 4))       # 1152 chunks
with uw.ChunkBuilder(uw.Perlin(0)) as b, uw.ChunkBuilder(uw.Perlin(0), staged=True) as s, uw.ChunkBuilder(uw.Perlin(0), ordered=True) as o:
    fb, sb, ob = b.build(big), s.build(big), o.build(big)
    same = all(fb.chunk(i).inds.tobytes() == sb.chunk(i).inds.tobytes() == ob.chunk(i).inds.tobytes()
               and fb.chunk(i).verts.tobytes() == sb.chunk(i).verts.tobytes() for i in range(len(big)))
    print("1152 chunks: fused(cost order) / staged(2 scan tiles) / ordered identical per chunk:", same, fb.n_inds)
    h0 = b.build_async(big[:600]); h1 = b.build_async(big[600:])
    a0 = b.wait(h0); a1 = b.wait(h1)
    print("two batches in flight:", a0.n_inds + a1.n_inds == fb.n_inds)
with uw.ChunkBuilder(uw.Perlin(0), internal_size=10, staged=True) as s10, uw.ChunkBuilder(uw.Perlin(0), internal_size=10) as f10:
    print("S=10 staged == fused:", s10.build(pos).n_inds == f10.build(pos).n_inds)
with uw.ChunkBuilder(uw.Perlin(0), analytic_skip=True) as b:
    print("analytic skip, cost order:", b.build(big).n_inds == fb.n_inds)
# a tall region in request order (n > 9472): ~25 chunks per CTA, most of them without a mesh -- runs of chunks that skip
# the end-of-chunk barrier, with the spare warp hashing ahead and the vote flags / ticket slots alternating
tall = uw.region.box_region((-12, 12), (-12, 12), (-13, 13))  # 14 976 chunks
with uw.ChunkBuilder(uw.Perlin(0)) as b, uw.ChunkBuilder(uw.Perlin(0), staged=True) as s:
    fb2, sb2 = b.build(tall), s.build(tall)
    print("tall region (request order, barrier-free mesh-less chunks): fused == staged:", fb2.n_inds == sb2.n_inds and fb2.n_verts == sb2.n_verts, fb2.n_inds)
print("done 2")
# round 2 paths: gather segments (local + forced staged stores), uw_multi_build, draw list, collision ray casts
import os
with uw.ChunkBuilder(uw.Perlin(0)) as b:
    ref = b.build(big)
    info = b.gather_create(1, len(big), seg_vcap=ref.n_verts, seg_icap=ref.n_inds)
    b.gather_attach(info, 0)
    b.gather_build(big, 0)
    res = b.gather_wait(descs_to_host=True, draw_to_host=True)
    print("gather (one local segment, register stores):", res.n_inds == ref.n_inds, res.n_draw)
    b.gather_detach(); b.gather_destroy()
os.environ["UW_STAGED_STORES"] = "1"
with uw.ChunkBuilder(uw.Perlin(0), ordered=True) as b, uw.ChunkBuilder(uw.Perlin(0), index32=True) as b32:
    st = b.build(big)
    print("staged 16-byte stores (u16):", st.n_inds == ob.n_inds if 'ob' in dir() else st.n_inds, " u32:", b32.build(pos).n_inds)
del os.environ["UW_STAGED_STORES"]
with uw.MultiBuilder(uw.Perlin(0), devices=[0]) as mb:
    print("uw_multi_build:", mb.build(big, draw_to_host=True).n_inds == ref.n_inds)
with uw.ChunkBuilder(uw.Perlin(0), tris=True) as b:
    b.build(pos)
    rng = np.random.default_rng(0)
    o = rng.uniform(-30, 30, size=(512, 3)).astype(np.float32); d = rng.normal(size=(512, 3)).astype(np.float32)
    print("raycast hits:", int((b.raycast_tris(o, d, 3) >= 0).sum()))
print("done 3")
