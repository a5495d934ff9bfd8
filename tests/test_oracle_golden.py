"""CPU tests: the oracle against outputs of the reference's own shipped binary (tests/golden/
ref_wasm_*.npz, produced by oracle/gen_golden.py), its frozen KATs, and analytic invariants."""
import os

import numpy as np
import pytest

from oracle import MODE_FAITHFUL, MODE_FAST, MODE_RULE, FLAG_BLANK_EARLY, FLAG_HAS_MESH


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32 if a.dtype == np.float32 else np.uint64)


def test_perm_table_matches_reference_binary(oracle12, golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_wasm_perm.npz"))
    for seed, perm in zip(g["seeds"], g["perms"]):
        got = oracle12.perm_table(int(seed))
        assert np.array_equal(got, perm), f"seed {seed}"
        assert sorted(got.tolist()) == list(range(256))


def test_perlin3_bit_exact_vs_reference_binary(oracle12, golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_wasm_perlin3.npz"))
    perms = {}
    bad = 0
    for seed, pt, want in zip(g["seeds"], g["points"], g["values"]):
        seed = int(seed)
        if seed not in perms:
            perms[seed] = oracle12.perm_table(seed)
        got = oracle12.perlin3(perms[seed], *pt.tolist())
        bad += np.float64(got).view(np.uint64) != np.float64(want).view(np.uint64)
    assert bad == 0
    assert len(g["values"]) == 4000


def test_perlin3_is_zero_on_the_integer_lattice(oracle12):
    perm = oracle12.perm_table(42)
    rng = np.random.default_rng(0)
    for p in rng.integers(-300, 300, size=(200, 3)):
        assert oracle12.perlin3(perm, float(p[0]), float(p[1]), float(p[2])) == 0.0


def test_chunk_build_matches_reference_binary_s10(oracle10, golden_dir):
    """Whole Chunk::build_full (S=10 as compiled into the shipped wasm): densities, positions and
    indices bit-exact; colours within 1 ulp (the wasm links Rust's libm port for powf, the oracle
    glibc's -- the reference itself differs between its web and native builds there)."""
    g = np.load(os.path.join(golden_dir, "ref_wasm_chunks_s10.npz"))
    meta = g["meta"]
    assert len(meta) >= 9
    n_mesh = 0
    for ci, (seed, x, y, z, finished, at_buffer, num_inds, calls) in enumerate(meta):
        perm = oracle10.perm_table(int(seed))
        r = oracle10.build_chunk(perm, (int(x), int(y), int(z)), MODE_FAITHFUL)
        if f"isos_{ci}" in g:
            assert np.array_equal(_bits(r["isos"]), _bits(g[f"isos_{ci}"])), f"densities chunk {ci}"
        if at_buffer:
            n_mesh += 1
            assert r["flags"] & FLAG_HAS_MESH
            want_v, want_i = g[f"verts_{ci}"], g[f"inds_{ci}"]
            assert len(r["inds"]) == num_inds == len(want_i)
            assert np.array_equal(r["inds"].astype(np.uint16), want_i), f"indices chunk {ci}"
            assert len(r["verts"]) == len(want_v)
            assert np.array_equal(_bits(r["verts"]["pos"]), _bits(want_v[:, :3])), f"positions chunk {ci}"
            np.testing.assert_allclose(r["verts"]["color"], want_v[:, 3:], rtol=0, atol=6e-8)
        else:
            assert finished and num_inds == 0
            assert len(r["inds"]) == 0 and not (r["flags"] & FLAG_HAS_MESH)
            if calls == 1:      # finished inside the Iso step -> early blank
                assert r["flags"] & FLAG_BLANK_EARLY
            else:               # went through the mesh stage and emitted nothing (all solid)
                assert not (r["flags"] & FLAG_BLANK_EARLY)
    assert n_mesh >= 6


def test_oracle_kats_s12(oracle12, golden_dir):
    g = np.load(os.path.join(golden_dir, "oracle_kat_s12.npz"))
    for ci, (seed, x, y, z, flags, nv, ni) in enumerate(g["meta"]):
        perm = oracle12.perm_table(int(seed))
        r = oracle12.build_chunk(perm, (int(x), int(y), int(z)), MODE_FAITHFUL)
        assert r["flags"] == flags and len(r["verts"]) == nv and len(r["inds"]) == ni
        assert np.array_equal(_bits(r["isos"]), _bits(g[f"isos_{ci}"]))
        assert np.array_equal(r["cases"], g[f"cases_{ci}"])
        assert np.array_equal(r["inds"].astype(np.uint16), g[f"inds_{ci}"])
        got = np.concatenate([r["verts"]["pos"], r["verts"]["color"]], axis=1)
        assert np.array_equal(_bits(got), _bits(g[f"verts_{ci}"]))


@pytest.mark.parametrize("S", [3, 5, 10, 12])
def test_modes_agree_on_random_fields(S):
    """faithful (linear-search dedup) == fast (edge map) == rule (static ownership + prefix sums,
    SURVEY App. B.4 -- the numbering the CUDA kernels implement), on terrain and on pure noise."""
    from oracle import Oracle
    o = Oracle(S)
    perm = o.perm_table(3)
    rng = np.random.default_rng(S)
    L = S + 1
    fields = [rng.uniform(-1, 1, L ** 3).astype(np.float32) for _ in range(4)]
    fields.append(np.where(rng.uniform(size=L ** 3) < 0.05, -1.0, 1.0).astype(np.float32))
    fields.append(np.where(rng.uniform(size=L ** 3) < 0.95, -1.0, 1.0).astype(np.float32))
    f = rng.uniform(-1, 1, L ** 3).astype(np.float32)
    f[::7] = np.float32(-0.1)           # corners exactly at the isovalue: "outside" for classify
    fields.append(f)
    for pos in [(0, 0, -1), (2, -3, 0)]:
        fields.append(o.densities(perm, pos))
    for f in fields:
        a = o.build_chunk(perm, (1, -2, 0), MODE_FAITHFUL, isos=f)
        for mode in (MODE_FAST, MODE_RULE):
            b = o.build_chunk(perm, (1, -2, 0), mode, isos=f)
            assert a["flags"] == b["flags"]
            assert np.array_equal(a["cases"], b["cases"])
            assert np.array_equal(a["inds"], b["inds"])
            assert np.array_equal(a["verts"].view(np.uint8), b["verts"].view(np.uint8))


def test_blank_uses_gt_and_classify_uses_lt(oracle12):
    """chunk.rs:132 vs :159 -- a corner exactly at ISO_LEVEL defeats the early-out but is 'outside'."""
    L = 13
    perm = oracle12.perm_table(0)
    iso = np.float32(-0.1)
    f = np.full(L ** 3, 1.0, dtype=np.float32)
    r = oracle12.build_chunk(perm, (0, 0, 0), MODE_FAITHFUL, isos=f)
    assert r["flags"] == FLAG_BLANK_EARLY
    f[100] = iso
    r = oracle12.build_chunk(perm, (0, 0, 0), MODE_FAITHFUL, isos=f)
    assert r["flags"] == 0 and len(r["inds"]) == 0 and not r["cases"].any()
    f[:] = -1.0                                    # all solid: mesh stage runs, emits nothing
    r = oracle12.build_chunk(perm, (0, 0, 0), MODE_FAITHFUL, isos=f)
    assert r["flags"] == 0 and len(r["inds"]) == 0 and (r["cases"] == 255).all()


def test_provably_trivial_layers(oracle12):
    """chunk z >= 2 is always blank, z <= -4 always solid (|noise| <= 1)."""
    perm = oracle12.perm_table(1)
    for z in (2, 3, 9):
        assert oracle12.build_chunk(perm, (3, -1, z), MODE_FAST)["flags"] == FLAG_BLANK_EARLY
    for z in (-4, -5, -11):
        r = oracle12.build_chunk(perm, (3, -1, z), MODE_FAST)
        assert r["flags"] == 0 and (r["cases"] == 255).all()


def test_colour_helpers(oracle12):
    # util.rs:122-153: pure hues; util.rs:106-112: srgb of 255 -> 1.0
    np.testing.assert_allclose(oracle12.hsv_to_rgb(0.0, 1.0, 1.0), [255, 0, 0], atol=1e-4)
    np.testing.assert_allclose(oracle12.hsv_to_rgb(120.0, 1.0, 1.0), [0, 255, 0], atol=1e-4)
    np.testing.assert_allclose(oracle12.hsv_to_rgb(-120.0, 1.0, 1.0), [0, 0, 255], atol=1e-4)   # rem_euclid
    np.testing.assert_allclose(oracle12.to_srgb([255.0, 255.0, 255.0]), [1, 1, 1], atol=1e-6)
    c = oracle12.vertex_color(-8.0, 5)
    assert c.shape == (3,) and np.all(c > 0) and np.all(c < 1)


def test_tri_lists_cover_every_triangle(oracle12):
    perm = oracle12.perm_table(0)
    r = oracle12.build_chunk(perm, (0, 0, -1), MODE_FAITHFUL, want_tris=True)
    assert len(r["tris"]) == len(r["inds"]) // 3
    assert r["tri_cell_start"][-1] == len(r["tris"])
    v = r["verts"]["pos"][r["inds"].reshape(-1, 3)]
    assert np.array_equal(v.view(np.uint32), r["tris"]["verts"].view(np.uint32))
    n = np.linalg.norm(r["tris"]["normal"], axis=1)
    assert np.all((np.abs(n - 1) < 1e-5) | (n == 0))


@pytest.mark.parametrize("S", [10, 12, 64])
def test_far_lattice_plane_carries_no_weight(S):
    """The argument behind SpecDims::prune() (uw_kernels.cuh): the chunk's last sample (k = S) is the only one that
    reaches the (2^o + 1)-th noise-lattice plane of an axis, and its fade weight on that plane is far below half an
    ulp of anything it is added to (exactly 0 when SIZE_SCALE is exact), so the FP32 kernels never touch the plane.
    f64 arithmetic of chunk.rs:107-116 / the noise crate's quintic fade, as the host builds its axis tables."""
    scale = float(np.float32(16.0) / np.float32(S))                  # SIZE_SCALE (f32), chunk.rs:7
    for o in range(3):
        cells = [int(np.floor(((k * scale) / 16.0) * (1 << o))) for k in range(S + 1)]
        assert max(cells[:-1]) == (1 << o) - 1 and cells[-1] == (1 << o)
        p = ((S * scale) / 16.0) * (1 << o)
        d = p - np.floor(p)
        w = (d * d * d) * (d * (d * 6.0 - 15.0) + 10.0)
        assert 0.0 <= w < 1e-15
        # dropped term <= w * |a1 - a0| <= 1e-15 * 4: nothing next to the FP32 path's 2e-6 density tolerance, and
        # samples inside the guard band are re-evaluated in f64 with the term in place
        assert w * 4.0 < 0.5 * float(np.spacing(np.float32(1e-6)))


def test_table_driven_pow24_scheme_is_within_two_ulp():
    """The arithmetic of pow24_tab (uw_kernels.cuh: 16-interval table of 1/c_i and a two-float log2 c_i, degree-4 and
    degree-6 polynomials, the product with 2.4f carried as a two-float sum) restated in numpy float32 (FMA emulated
    through float64, whose product of two floats is exact): <= 2.5 ulp against the unrounded f64 pow over the colour's
    input range, i.e. at most 2 ulp from the correctly rounded f32 result.
    The device code itself is checked on the GPU (test_vertex_colour_function_over_the_whole_hue_range)."""
    f32 = np.float32

    def fma(a, b, c):
        return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)

    i = np.arange(16)
    inv_c = (1.0 / (1.0 + (i + 0.5) / 16.0)).astype(f32)
    l2 = -np.log2(inv_c.astype(np.float64))
    hi = np.rint(l2 * 65536.0) / 65536.0
    lh, ll = hi.astype(f32), (l2 - hi).astype(f32)
    assert (lh.astype(np.float64) == hi).all()                       # 16 fractional bits: exact in f32
    P = [f32(v) for v in (1.4426950216293335, -0.721347451210022, 0.48089829087257385, -0.3609875738620758, 0.2888079881668091)]
    Q = [f32(v) for v in (1.0, 0.6931471824645996, 0.24022650718688965, 0.05550327152013779, 0.009618056938052177,
                          0.0013400427997112274, 0.00015461444854736328)]
    rng = np.random.default_rng(3)
    b = np.concatenate([rng.uniform(0.04, 1.1, 300000), np.linspace(0.05, 1.0, 100001)]).astype(f32)
    bits = b.view(np.uint32)
    e = (bits >> 23).astype(np.int32) - 127
    m = ((bits & 0x7FFFFF) | 0x3F800000).view(f32)
    k = ((bits >> 19) & 15).astype(np.int64)
    r = fma(m, inv_c[k], f32(-1.0))
    p = fma(P[4], r, P[3])
    for c in (P[2], P[1], P[0]):
        p = fma(p, r, c)
    Lh = e.astype(f32) + lh[k]
    Ll = fma(p, r, ll[k])
    pw = f32(2.4)
    yh = (pw * Lh).astype(f32)
    yl = fma(pw, Ll, fma(pw, Lh, -yh))
    n = np.rint(yh)
    f = ((yh - n).astype(f32) + yl).astype(f32)
    q = fma(Q[6], f, Q[5])
    for c in (Q[4], Q[3], Q[2], Q[1], Q[0]):
        q = fma(q, f, c)
    got = np.ldexp(q, n.astype(np.int32)).astype(f32)
    want = np.power(b.astype(np.float64), float(pw))
    ulp = np.abs(got.astype(np.float64) - want) / np.spacing(want.astype(f32)).astype(np.float64)
    assert ulp.max() <= 2.5, ulp.max()
    assert (ulp > 1.0).mean() < 0.01


def test_factorised_lattice_hash_of_the_spare_warp(oracle12):
    """noise_stage_h_warp (uw_kernels.cuh): the lattice hash perm[perm[perm[X] ^ Y] ^ Z] of every touched lattice point,
    factorised -- one lane per (octave, cx, cy) computes perm[perm[X] ^ Y] (25 lanes for the top octave, 4 + 9 for the
    others), the G^3 last-level lookups take it by shuffle from lane (cx G + cy) [+ the octave's offset], six rounds of
    32 lanes, lanes past the end repeating the last point.  Restated lane by lane in numpy against the direct triple
    lookup the other kernels (and the reference, through the noise crate's PermutationTable::hash) use."""
    perm = oracle12.perm_table(0).astype(np.int64)
    NOCT, OT = 3, 2
    G = [(1 << o) + 1 for o in range(NOCT)]                          # PRUNE: 2, 3, 5 lattice planes per axis
    g2_base = [0, 4, 13]
    lat_base = [0, 8, 35]
    lane = np.arange(32)
    for px, py, pz in ((0, 0, 0), (3, -2, 1), (-70000, 123456, -5), (2 ** 24, -(2 ** 24), 255)):
        # second level
        l = np.minimum(lane, G[OT] ** 2 - 1)
        cx, cy = l // G[OT], l % G[OT]
        hb_top = perm[perm[((px << OT) + cx) & 255] ^ (((py << OT) + cy) & 255)]
        o_of = np.zeros(32, dtype=np.int64); cxl = np.zeros(32, dtype=np.int64); cyl = np.zeros(32, dtype=np.int64)
        for p in range(OT):
            r = lane - g2_base[p]
            sel = (r >= 0) & (r < G[p] ** 2)
            o_of[sel], cxl[sel], cyl[sel] = p, r[sel] // G[p], r[sel] % G[p]
        hb_low = perm[perm[((px << o_of) + cxl) & 255] ^ (((py << o_of) + cyl) & 255)]
        # last level, round by round
        got = np.full(160, -1, dtype=np.int64)
        rounds = 0
        for o in range(NOCT):
            n = G[o] ** 3
            for t0 in range(0, n, 32):
                tt = np.minimum(t0 + lane, n - 1)
                cxy, cz = tt // G[o], tt % G[o]
                src = (0 if o == OT else g2_base[o]) + cxy
                assert src.max() < 32
                hb = (hb_top if o == OT else hb_low)[src]
                got[lat_base[o] + tt] = perm[hb ^ (((pz << o) + cz) & 255)]
                rounds += 1
        assert rounds == 6
        want = np.empty(160, dtype=np.int64)
        for o in range(NOCT):
            F = 1 << o
            for t in range(G[o] ** 3):
                cx, r = divmod(t, G[o] ** 2)
                cy, cz = divmod(r, G[o])
                want[lat_base[o] + t] = perm[perm[perm[(F * px + cx) & 255] ^ ((F * py + cy) & 255)] ^ ((F * pz + cz) & 255)]
        assert np.array_equal(got, want)


def test_stage_x_by_cell_groups_covers_every_x_sample_once():
    """Stage X of the S = 12 kernels (noise_chunk_spec, XG): thread (q, octave, r) handles the three samples
    i = 3q .. 3q + 2 from ONE pair of lattice planes, c = (q << o) >> 2 and c + 1; 38 more threads write the last sample
    plane i = 12 from plane G - 1 alone.  Restated in numpy: the (octave, i, r) items are covered exactly once, with the
    lattice planes the per-item loop uses ((i << o) / 12 and min(c + 1, G - 1))."""
    S, NOCT, NG, GS = 12, 3, 4, 3
    G = [(1 << o) + 1 for o in range(NOCT)]
    g2 = [g * g for g in G]
    g2_base = [0, 4, 13]
    SG2 = sum(g2)                                                    # 38
    seen = {}
    for tid in range(NG * SG2):                                      # the triple items, q-major
        q, u = divmod(tid, SG2)
        o = max(p for p in range(NOCT) if u >= g2_base[p])
        r = u - g2_base[o]
        assert 0 <= r < g2[o]
        c = (q << o) >> (NOCT - 1)
        for m in range(GS):
            i = q * GS + m
            assert (o, i, r) not in seen
            seen[(o, i, r)] = (c, c + 1)
    for t in range(SG2):                                             # the last sample plane
        o = max(p for p in range(NOCT) if t >= g2_base[p])
        r = t - g2_base[o]
        assert (o, S, r) not in seen
        seen[(o, S, r)] = (1 << o, 1 << o)
    want = {}
    for o in range(NOCT):
        for i in range(S + 1):
            c = (i << o) // S
            for r in range(g2[o]):
                want[(o, i, r)] = (c, min(c + 1, G[o] - 1))
    assert seen == want and len(seen) == 13 * SG2 == 494
    assert NG * SG2 == 152 and 160 + SG2 <= 224 and 152 + SG2 <= 192    # one round: fused (224 threads) and staged (192)
