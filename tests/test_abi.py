"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/uwcuda.h declares; the host logic (regions, slabs, Chunk mirror) behaves.  No compute
calls -- those need a GPU and live in test_gpu_parity.py."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import underwaterworld_b200 as uw
from underwaterworld_b200 import _ffi, region
from underwaterworld_b200.build import build_library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build_library()
    return _ffi.load_library()


def test_header_symbols_are_all_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "uwcuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(uw_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_ffi.EXPORTS)
    for sym in declared:
        assert hasattr(lib, sym), f"libuwcuda.so does not export {sym}"
    assert lib.uw_abi_version() == 2


def test_config_default_matches_reference_constants(lib):
    cfg = _ffi.UwConfig()
    lib.uw_config_default(C.byref(cfg))
    # chunk.rs:5-17, world.rs:11-12
    assert (cfg.internal_size, cfg.chunk_size, cfg.octaves) == (12, 16, 3)
    assert cfg.iso_level == np.float32(-0.1) and cfg.max_height == 32.0 and cfg.adj_z_mod == 0.25
    assert (cfg.min_hue, cfg.max_hue) == (-150.0, 60.0)
    assert cfg.saturation == np.float32(0.6) and cfg.base_value == np.float32(0.4)
    assert (cfg.min_z, cfg.max_z) == (-2.0, 2.0)
    assert C.sizeof(_ffi.UwConfig) == 80


def test_struct_layouts():
    assert _ffi.VERT_DTYPE.itemsize == 24          # draw.rs:4-9
    assert _ffi.VERT_DTYPE.fields["color"][1] == 12  # attribute offsets 0 and 12, draw.rs:19-29
    assert _ffi.DESC_DTYPE.itemsize == 32
    assert _ffi.TRI_DTYPE.itemsize == 48           # util.rs:7-10


def test_create_rejects_bad_config_and_reports_errors(lib):
    cfg = _ffi.UwConfig()
    lib.uw_config_default(C.byref(cfg))
    ctx = C.c_void_p()
    assert lib.uw_create(None, C.byref(ctx)) == _ffi.UW_ERR_INVALID
    cfg.octaves = 9
    assert lib.uw_create(C.byref(cfg), C.byref(ctx)) == _ffi.UW_ERR_INVALID
    assert b"octaves" in lib.uw_last_error(None)
    cfg.octaves = 3
    cfg.internal_size = 0
    assert lib.uw_create(C.byref(cfg), C.byref(ctx)) in (_ffi.UW_ERR_INVALID, _ffi.UW_ERR_UNSUPPORTED)
    assert not ctx.value


def test_no_cpu_fallback_without_a_device(lib):
    """Without a CUDA device uw_create must fail loudly (UW_ERR_NO_DEVICE), never compute on the CPU."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(uw.UwError) as ei:
        uw.ChunkBuilder()
    assert ei.value.status == _ffi.UW_ERR_NO_DEVICE
    assert "no CPU fallback" in str(ei.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "underwaterworld_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
                assert "uw_oracle" not in src and "libuw_oracle" not in src, fn


def test_regions_and_slabs():
    p = region.config_positions("spawn")
    assert p.shape == (2048, 3) and p.dtype == np.int32
    assert tuple(p[0]) == (-8, -8, -4) and tuple(p[1]) == (-8, -8, -3) and tuple(p[8]) == (-8, -7, -4)   # z fastest
    assert region.config_positions("large").shape == (524288, 3)
    for world in (1, 2, 3, 4, 8):
        parts = [region.shard_region((-64, 64), (-2, 2), (-1, 1), r, world) for r in range(world)]
        whole = np.concatenate(parts)
        assert np.array_equal(whole, region.box_region((-64, 64), (-2, 2), (-1, 1)))   # contiguous x-slabs, in order
        sizes = [len(q) for q in parts]
        assert max(sizes) - min(sizes) <= 2 * 4 * 2     # at most one x-column of imbalance
    assert region.slab_bounds(10, 0, 3) == (0, 4) and region.slab_bounds(10, 2, 3) == (7, 10)


def test_chunk_mirror_host_logic():
    c = uw.Chunk.new((1, -2, 3))
    assert c.chunk_offset == (16, -32, 48)          # chunk.rs:90-94
    assert not c.not_blank() and c.num_inds() == 0
    with pytest.raises(RuntimeError):               # the reference unwraps None -> panic (chunk.rs:346)
        c.verts_buffer_slice()
    mesh = uw.ChunkMesh((0, 0, 0), _ffi.CHUNK_HAS_MESH, np.zeros(3, _ffi.VERT_DTYPE), np.arange(3, dtype=np.uint16))
    c._adopt(mesh)
    assert c.not_blank() and c.num_inds() == 3 and len(c.verts_buffer_slice()) == 3
    assert uw.Perlin(2 ** 32 + 5).seed() == 5


def test_weighted_slab_bounds_mirror_the_library(lib):
    """Gather-aware partition (uw_slab_bounds_weighted, pure host arithmetic): the Python mirror RegionGather.plan uses
    equals the C function uw_multi_build uses; slabs tile the request in part order; the rendering slab never gets less
    than an even share."""
    from underwaterworld_b200 import gather
    for n, parts, rp, pm in [(524288, 8, 0, 232), (524288, 8, 0, 0), (100, 4, 2, 500), (7, 3, 0, 400), (2048, 8, 0, 100),
                             (10, 2, 1, 900), (0, 4, 0, 300), (5, 1, 0, 700), (1000, 5, 4, 1000)]:
        cover, counts = 0, []
        for p in range(parts):
            f, c = C.c_uint32(), C.c_uint32()
            lib.uw_slab_bounds_weighted(n, parts, p, rp, pm, C.byref(f), C.byref(c))
            assert (f.value, c.value) == gather.slab_bounds_weighted(n, parts, p, rp, pm)
            assert f.value == cover
            cover += c.value
            counts.append(c.value)
        assert cover == n
        if parts > 1 and pm * parts > 1000:
            assert counts[rp] == (n * min(pm, 1000) + 500) // 1000 and counts[rp] >= n // parts
    # the balance controller: moves towards equal finish times, stays within [even, 1/2], holds when balanced
    s = 0.0
    for _ in range(12):
        t_r = max(s, 1 / 8) * 3.8
        t_o = max((1 - max(s, 1 / 8)) / 7 * 3.8, (1 - max(s, 1 / 8)) * 1.14)
        s = gather.balance_share(s, 8, t_r, t_o)
    assert 0.20 < s < 0.26
    assert gather.balance_share(0.3, 4, 1.0, 1.02) == 0.3 and gather.balance_share(0.0, 4, 2.0, 1.0) == 0.25
    # the search uw_multi_build runs between requests (uw_share_search_next, a pure function of its state): the C
    # function and its Python mirror take the same steps; on a model of the 8-GPU gather (the rendering GPU's kernel is
    # slowed by the traffic arriving over NVLink, the producers are ingress-bound) it settles at the cost minimum, where
    # equalising kernel times (above) would stop early
    for parts, model in ((8, lambda f: max(3.25 * f * (1.0 + 1.2 * (1.0 - f)), 1.03 * (1.0 - f))),
                         (4, lambda f: max(3.25 * f, 0.82 * (1.0 - f) / 3 * 4, 1.03 * (1.0 - f))),
                         (2, lambda f: 1.0 + f)):
        st, mirror = _ffi.UwShareSearch(), {}
        share = lib.uw_share_search_next(C.byref(st), parts, 0.0)          # not started: the even split
        assert share == gather.share_search_next(mirror, parts, 0.0) == 1.0 / parts
        seen = []
        for k in range(40):
            cost = model(share) * (1.0 + 0.004 * ((k * 7919) % 5 - 2))       # +-0.8 % measurement noise
            seen.append((cost, share))
            share = lib.uw_share_search_next(C.byref(st), parts, cost)
            assert share == gather.share_search_next(mirror, parts, cost)
            assert 1.0 / parts <= share <= 0.5
        grid = [1.0 / parts + k * (0.5 - 1.0 / parts) / 400 for k in range(401)]
        best = min(grid, key=model)
        assert abs(share - best) <= 0.02 and model(share) <= model(best) * 1.03, (parts, share, best)
        assert st.settled == 1 and share == st.best_share                   # ... and stays there
        # another workload (everything 30 % slower): the search starts over from where it is
        assert lib.uw_share_search_next(C.byref(st), parts, 1.3 * model(share)) == gather.share_search_next(mirror, parts, 1.3 * model(share))
        assert st.settled == 0 or parts == 2
    assert lib.uw_share_search_next(C.byref(_ffi.UwShareSearch()), 1, 1.0) == 0.0 and C.sizeof(_ffi.UwShareSearch) == 56
