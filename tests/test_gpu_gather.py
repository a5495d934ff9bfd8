"""GPU tests of the multi-GPU path (include/uwcuda.h uw_gather_* / uw_multi_*; SURVEY §8e, BASELINE configs[2]).

Chunks are independent (chunk.rs:89-129), so a region built as slabs -- by several producers, into segments of the
rendering GPU's arenas, through local, peer or CUDA-IPC addresses -- must give every chunk exactly the buffers a
plain single-context build gives it.  Everything below runs on ONE GPU too (several producers may share a device);
with two or more GPUs visible the same tests also cross NVLink.
"""
import multiprocessing as mp
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def uw():
    import underwaterworld_b200 as m
    m.load_library()
    return m


def _device_count():
    import torch
    return torch.cuda.device_count()


def _assert_same_chunks(descs, verts, inds, ref_batch, first=0):
    """Arena contents (descs/verts/inds addressed through the descriptors) == the reference batch, chunk by chunk,
    bit for bit (same kernel, same arithmetic: only the placement differs)."""
    for i in range(len(ref_batch)):
        d, r = descs[first + i], ref_batch.chunk(i)
        assert tuple(int(v) for v in d["pos"]) == r.pos
        assert int(d["flags"]) == r.flags
        assert int(d["vert_count"]) == len(r.verts) and int(d["index_count"]) == len(r.inds), f"counts of chunk {i}"
        vo, io = int(d["vert_offset"]), int(d["index_offset"])
        assert np.array_equal(verts[vo:vo + len(r.verts)].view(np.uint8), r.verts.view(np.uint8)), f"vertices of chunk {i}"
        assert np.array_equal(inds[io:io + len(r.inds)], r.inds), f"indices of chunk {i}"


def test_slab_bounds_match_the_library(uw):
    import ctypes as C
    lib = uw.load_library()
    for n, parts in [(0, 3), (1, 4), (7, 3), (2048, 8), (524288, 8), (17 * 6 * 8, 5)]:
        cover = 0
        for p in range(parts):
            f, c = C.c_uint32(), C.c_uint32()
            lib.uw_slab_bounds(n, parts, p, C.byref(f), C.byref(c))
            assert (f.value, c.value) == uw.gather.slab_bounds(n, parts, p)
            assert f.value == cover
            cover += c.value
        assert cover == n


def test_single_segment_gather_equals_plain_build(uw):
    pos = uw.region.box_region((-3, 3), (-3, 3), (-3, 2))            # 180 chunks through the surface layers
    with uw.ChunkBuilder(uw.Perlin(0)) as ref_b, uw.ChunkBuilder(uw.Perlin(0)) as b:
        ref = ref_b.build(pos)
        info = b.gather_create(1, len(pos), seg_vcap=ref.n_verts + 2, seg_icap=ref.n_inds)     # exactly enough
        b.gather_attach(info, 0)
        for rep in range(2):                                          # epochs advance; the arena is reused
            b.gather_build(pos, 0)
            res = b.gather_wait(descs_to_host=True)
            assert res.epoch == rep + 1 and res.n_segments == 1
            assert res.n_chunks == len(pos) and res.n_inds == ref.n_inds
            assert res.segments[0]["n_mesh"] == int((ref.descs["index_count"] > 0).sum())
            assert res.segments[0]["n_blank"] == int((ref.descs["flags"] & 1).sum())
            descs, verts, inds = res.download()
            _assert_same_chunks(descs, verts, inds, ref)
            assert np.array_equal(res.host_descs().view(np.uint8), descs.view(np.uint8))
        b.gather_detach()
        b.gather_destroy()
        assert np.array_equal(b.build(pos).descs["index_count"], ref.descs["index_count"])     # plain builds still work


def test_three_producers_fill_three_segments(uw):
    """Three contexts -- spread over the visible GPUs, so with >= 2 GPUs two of them write over NVLink -- each build
    one slab of the request into its own segment of context 0's arenas."""
    ndev = _device_count()
    pos = uw.region.box_region((-4, 5), (-2, 2), (-3, 2))            # 9 x-columns: an uneven 3-way split
    with uw.ChunkBuilder(uw.Perlin(3), device=0) as ref_b:
        ref = ref_b.build(pos)
    builders = [uw.ChunkBuilder(uw.Perlin(3), device=g % ndev) for g in range(3)]
    try:
        info = builders[0].gather_create(3, len(pos), seg_vcap=ref.n_verts, seg_icap=ref.n_inds)
        for g, b in enumerate(builders):
            b.gather_attach(info, g)
        for g in (1, 2, 0):
            f, c = uw.gather.slab_bounds(len(pos), 3, g)
            builders[g].gather_build(pos[f:f + c], f)
        res = builders[0].gather_wait()
        for b in builders:
            b.sync()
        assert res.n_chunks == len(pos) and res.n_inds == ref.n_inds
        assert [s["first_chunk"] for s in res.segments] == [uw.gather.slab_bounds(len(pos), 3, g)[0] for g in range(3)]
        descs, verts, inds = res.download()
        _assert_same_chunks(descs, verts, inds, ref)
        for g in range(3):                                            # every chunk's buffers lie inside its producer's segment
            f, c = uw.gather.slab_bounds(len(pos), 3, g)
            d = descs[f:f + c]
            d = d[d["index_count"] > 0]
            assert (d["vert_offset"] >= g * res.seg_vcap).all() and (d["vert_offset"] + d["vert_count"] <= (g + 1) * res.seg_vcap).all()
            assert (d["index_offset"] >= g * res.seg_icap).all() and (d["index_offset"] + d["index_count"] <= (g + 1) * res.seg_icap).all()
    finally:
        for b in builders[::-1]:
            b.gather_detach()
        for b in builders:
            b.close()


def test_multi_builder_region_equals_plain_build(uw):
    """uw_multi_build over every visible GPU (one process): slabs, fused launches, gather, wait -- and a second,
    larger request that makes it recreate the arena."""
    ndev = min(_device_count(), 8)
    small = uw.region.box_region((-2, 2), (-2, 2), (-3, 2))
    large = uw.region.box_region((-8, 8), (-8, 8), (-4, 4))          # config 2
    with uw.ChunkBuilder(uw.Perlin(0), device=0) as ref_b, uw.MultiBuilder(uw.Perlin(0), devices=list(range(ndev))) as mb:
        for pos in (small, large, small):
            ref = ref_b.build(pos)
            res = mb.build(pos, descs_to_host=True, draw_to_host=True)
            assert res.n_segments == ndev and res.n_chunks == len(pos) and res.n_inds == ref.n_inds
            descs, verts, inds = res.download()
            _assert_same_chunks(descs, verts, inds, ref)
            assert np.array_equal(res.host_descs().view(np.uint8), descs.view(np.uint8))
            # the draw list = exactly the descriptors of the chunks that ended with a mesh (any order inside a segment)
            draw = res.host_draw()
            meshed = descs[descs["index_count"] > 0]
            assert res.n_draw == len(meshed) == len(draw)
            key = lambda a: sorted(bytes(x) for x in a.view(np.uint8).reshape(len(a), 32))
            assert key(draw) == key(meshed)
            at = 0
            for g, seg in enumerate(res.segments):                    # segment g's entries point into segment g
                part = draw[at:at + seg["n_mesh"]]
                assert (part["vert_offset"] // res.seg_vcap == g).all()
                at += seg["n_mesh"]


def test_multi_builder_adapts_the_render_share_without_changing_results(uw):
    """uw_multi_build balances the rendering GPU's slab against the slowest producer over successive requests
    (gather-aware partition); whatever the split, every chunk's buffers stay those of the plain build."""
    import ctypes as C
    ndev = min(_device_count(), 8)
    pos = uw.region.box_region((-24, 24), (-24, 24), (-6, 4))        # 23 040 chunks
    with uw.ChunkBuilder(uw.Perlin(0), device=0) as ref_b, uw.MultiBuilder(uw.Perlin(0), devices=list(range(ndev))) as mb:
        ref = ref_b.build(pos)
        shares = []
        for _ in range(5):
            res = mb.build(pos)
            shares.append(mb._lib.uw_multi_render_share(mb._m))
            assert res.n_chunks == len(pos) and res.n_inds == ref.n_inds
        descs, verts, inds = res.download()
        _assert_same_chunks(descs, verts, inds, ref)
        assert [s["first_chunk"] for s in res.segments] == [uw.gather.slab_bounds_weighted(len(pos), ndev, g, 0, shares[-2])[0] for g in range(ndev)]
        if ndev == 1:
            assert shares == [0] * 5
        else:
            assert all(0 <= s <= 500 for s in shares)


def test_segment_overflow_is_reported_and_multi_build_regrows(uw):
    pos = uw.region.box_region((-4, 4), (-4, 4), (-1, 0))            # one surface layer: ~ 500 vertices per chunk
    with uw.ChunkBuilder(uw.Perlin(0)) as b:
        info = b.gather_create(1, len(pos), seg_vcap=1024, seg_icap=4096)
        b.gather_attach(info, 0)
        b.gather_build(pos, 0)
        with pytest.raises(uw.UwError) as e:
            b.sync()
        assert e.value.status == 4 and "overflow" in str(e.value)     # UW_ERR_OOM
        b.gather_detach()
        b.gather_destroy()
        ref = b.build(pos)
    # uw_multi_build starts from the default estimate (192 vertices per chunk) and must grow for this request
    dense = uw.region.box_region((-8, 8), (-8, 8), (-1, 0))
    with uw.ChunkBuilder(uw.Perlin(0)) as ref_b, uw.MultiBuilder(uw.Perlin(0), devices=[0]) as mb:
        ref = ref_b.build(dense)
        assert ref.n_verts > len(dense) * 192 + 4096, "the request must exceed the default capacity for this test to bite"
        res = mb.build(dense)
        descs, verts, inds = res.download()
        _assert_same_chunks(descs, verts, inds, ref)


def _ipc_child(info_bytes, first, count, seed, q):
    try:
        import underwaterworld_b200 as uw
        pos = uw.region.box_region((-3, 3), (-3, 3), (-3, 2))[first:first + count]
        ndev = _device_count()
        with uw.ChunkBuilder(uw.Perlin(seed), device=1 % ndev) as b:
            b.gather_attach(uw.gather.info_from_bytes(info_bytes), 1)
            b.gather_build(pos, first)
            b.sync()
            b.gather_detach()
        q.put("ok")
    except Exception as e:                                            # pragma: no cover
        q.put(f"child failed: {type(e).__name__}: {e}")


@pytest.mark.timeout(300)
def test_producer_in_another_process_attaches_through_cuda_ipc(uw):
    """One process per GPU (the torchrun layout of bench.py): the producer maps the rendering process's arena with
    the cudaIpcMemHandle inside uw_gather_info and writes its segment from its own process (and, when a second GPU
    is visible, from that GPU over NVLink)."""
    pos = uw.region.box_region((-3, 3), (-3, 3), (-3, 2))
    f1, c1 = uw.gather.slab_bounds(len(pos), 2, 1)
    with uw.ChunkBuilder(uw.Perlin(5), device=0) as b:
        ref = b.build(pos)
        info = b.gather_create(2, len(pos), seg_vcap=ref.n_verts, seg_icap=ref.n_inds)
        b.gather_attach(info, 0)
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        child = ctx.Process(target=_ipc_child, args=(uw.gather.info_to_bytes(info), f1, c1, 5, q))
        child.start()
        b.gather_build(pos[:f1], 0)
        msg = q.get(timeout=240)
        child.join(60)
        assert msg == "ok", msg
        res = b.gather_wait()
        assert res.n_chunks == len(pos) and res.n_inds == ref.n_inds
        descs, verts, inds = res.download()
        _assert_same_chunks(descs, verts, inds, ref)
        b.gather_detach()
        b.gather_destroy()


def test_staged_stores_give_the_same_buffers_as_register_stores(uw):
    """Producers that write into ANOTHER GPU's memory leave vertices and indices through shared memory as whole
    16-byte vectors (FusedOut::staged_stores); the GPU's own HBM is written straight from registers.  Same chunks,
    same bytes -- UW_STAGED_STORES=1 forces the staged path for a local arena so one GPU can compare the two,
    including chunks with more indices than one staging tile holds."""
    pos = uw.region.config_positions("spawn")
    with uw.ChunkBuilder(uw.Perlin(0), ordered=True) as b:
        want = b.build(pos)
    os.environ["UW_STAGED_STORES"] = "1"
    try:
        with uw.ChunkBuilder(uw.Perlin(0), ordered=True) as b:
            got = b.build(pos)
        with uw.ChunkBuilder(uw.Perlin(0), ordered=True, index32=True) as b:
            got32 = b.build(pos)
    finally:
        del os.environ["UW_STAGED_STORES"]
    assert want.descs["index_count"].max() > 3072                      # more than one u16 staging tile
    assert np.array_equal(got.descs.view(np.uint8), want.descs.view(np.uint8))
    assert np.array_equal(got.inds, want.inds) and np.array_equal(got.verts.view(np.uint8), want.verts.view(np.uint8))
    for i in range(0, len(pos), 7):
        a, w = got32.chunk(i), want.chunk(i)
        assert np.array_equal(a.inds, w.inds.astype(np.uint32)) and np.array_equal(a.verts.view(np.uint8), w.verts.view(np.uint8))


def test_raycast_against_collision_triangles_matches_the_reference_loops(uw):
    """uw_raycast_tris (SURVEY 8f-1's consumer) against the oracle's restatement of boid.rs:175-240 +
    Chunk::tris_around (chunk.rs:315-342) + Tri::intersects (util.rs:22-59): the smallest hit distance of every
    ray, bit for bit (same candidate triangles, same f32 operation order), -1 where the reference finds None."""
    from oracle import Oracle
    o = Oracle(12)
    perm = o.perm_table(0)
    pos = uw.region.box_region((-2, 2), (-2, 2), (-2, 1))              # 48 chunks around the surface
    rng = np.random.default_rng(7)
    n = 6000
    org = rng.uniform([-34, -34, -34], [34, 34, 18], size=(n, 3)).astype(np.float32)     # some origins outside every built chunk
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    # boids sit near the surface: pull most origins close to a vertex of the mesh.  The exact-f64 builder's triangles
    # are bit-identical to the oracle's (test_collision_tris_bit_exact), so its hit distances must be too.
    with uw.ChunkBuilder(uw.Perlin(0), tris=True, exact_f64=True) as b:
        batch = b.build(pos)
        verts, _ = batch.compact()
        pick = rng.integers(0, len(verts), size=n - 500)
        org[500:] = verts["pos"][pick] + rng.normal(scale=1.5, size=(n - 500, 3)).astype(np.float32)
        got = b.raycast_tris(org, d, 3)
        again = b.raycast_tris(org[:100], d[:100], 3)                   # the chunk table is reused
        wide = b.raycast_tris(org[:800], d[:800], 5)
    want = o.raycast(perm, pos, org, d, 3)
    assert (want >= 0).sum() > 1000 and (want < 0).sum() > 1000
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(again, got[:100])
    assert np.array_equal(wide.view(np.uint32), o.raycast(perm, pos, org[:800], d[:800], 5).view(np.uint32))
    # the shipped FP32 path: same triangles up to the stated vertex tolerance -> same hits except for rays that graze
    # an edge, same distances within that tolerance
    with uw.ChunkBuilder(uw.Perlin(0), tris=True) as b:
        b.build(pos)
        fast = b.raycast_tris(org, d, 3)
    both = (fast >= 0) & (want >= 0)
    assert ((fast >= 0) != (want >= 0)).sum() <= 0.005 * n
    assert np.abs(fast[both] - want[both]).max() < 1e-3
    with uw.ChunkBuilder(uw.Perlin(0)) as plain:
        plain.build(pos[:2])
        with pytest.raises(uw.UwError):
            plain.raycast_tris(org[:1], d[:1])                          # needs UW_FLAG_TRIS


def test_pinned_requests_skip_the_host_scan_and_are_validated_on_the_device(uw):
    """A request of >= 4096 chunks in page-locked memory goes to the device straight from the caller's buffer and is
    not scanned on the host; the fused kernel checks every position it fetches (|pos| <= 2^24) and
    the build fails with UW_ERR_INVALID at its wait -- through uw_build and through uw_gather_build alike."""
    import torch
    pos = uw.region.box_region((-16, 16), (-16, 16), (-6, 6))          # 12 288 chunks: beyond the cost-ordered range
    pin = torch.from_numpy(pos.copy()).pin_memory()
    good = pin.numpy()
    with uw.ChunkBuilder(uw.Perlin(0)) as b, uw.ChunkBuilder(uw.Perlin(0)) as ref_b:
        ref = ref_b.build(pos)                                          # pageable: staged + scanned on the host
        got = b.build(good)                                             # pinned: direct
        assert np.array_equal(got.descs["index_count"], ref.descs["index_count"])
        bad = torch.from_numpy(pos.copy()).pin_memory()
        bad.numpy()[5000, 1] = (1 << 24) + 1
        with pytest.raises(uw.UwError) as e:
            b.build(bad.numpy())
        assert e.value.status == 1 and "out of supported range" in str(e.value)      # UW_ERR_INVALID
        assert np.array_equal(b.build(good).descs["index_count"], ref.descs["index_count"])   # the context recovers
        info = b.gather_create(1, len(pos), seg_vcap=ref.n_verts, seg_icap=ref.n_inds)
        b.gather_attach(info, 0)
        b.gather_build(bad.numpy(), 0)
        with pytest.raises(uw.UwError) as e:
            b.sync()
        assert e.value.status == 1
        b.gather_build(good, 0)
        res = b.gather_wait()
        assert res.n_inds == ref.n_inds
        b.gather_detach()
        b.gather_destroy()
