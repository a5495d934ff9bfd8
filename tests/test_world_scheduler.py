"""CPU tests of the host-side scheduler mirror (underwaterworld_b200/world.py) against a scalar restatement
of the reference loops (src/world.rs:148-234, src/util.rs:77-85) written out chunk by chunk."""
import math

import numpy as np

from underwaterworld_b200 import world as W
from underwaterworld_b200.chunk import Batch, Chunk, ChunkMesh
from underwaterworld_b200._ffi import DESC_DTYPE, VERT_DTYPE, CHUNK_HAS_MESH


def _scalar_update_nearby(sub_pos, sub_chunk, cam):
    """world.rs:152-233 as nested loops (no chunk map: everything is 'to generate')."""
    view_vp = cam.chunk_generation_frustum_matrix(W.VIEW_FRUST_FOVY)
    gen_vp = cam.chunk_generation_frustum_matrix(W.GENERATE_FRUST_FOVY)

    def inside(pt, m):
        c = m @ np.array([pt[0], pt[1], pt[2], 1.0])
        c = c / c[3]
        return abs(c[0]) <= 1.0 and abs(c[1]) <= 1.0 and 0.0 <= c[2] <= 1.0

    out = []
    start_z, end_z = max(sub_chunk[2] - W.GENERATION_DIST, W.MIN_Z), min(sub_chunk[2] + W.GENERATION_DIST, W.MAX_Z)
    for x in range(-W.GENERATION_DIST, W.GENERATION_DIST):
        for y in range(-W.GENERATION_DIST, W.GENERATION_DIST):
            for cz in range(start_z, end_z + 1):
                cx, cy = sub_chunk[0] + x, sub_chunk[1] + y
                center = np.array([(cx + 0.5) * 16, (cy + 0.5) * 16, (cz + 0.5) * 16])
                dist = np.linalg.norm(np.asarray(sub_pos) - center)
                if dist > W.GENERATION_DIST * 16:
                    continue
                in_view = in_gen = False
                for c in [(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0), (0, 0, 1), (1, 0, 1), (0, 1, 1), (1, 1, 1)]:
                    pt = ((cx + c[0]) * 16, (cy + c[1]) * 16, (cz + c[2]) * 16)
                    if inside(pt, gen_vp):
                        in_gen = True
                        if inside(pt, view_vp):
                            in_view = True
                            break
                out.append(((cx, cy, cz), dist, in_view, in_gen))
    return out


def test_candidates_match_the_scalar_reference_loops():
    n_checked = 0
    for frame, sub, cam in W.scripted_flythrough(40, 200):
        if frame % 37:
            continue
        pos, dist, in_view, in_gen = W.nearby_candidates(sub.pos, sub.chunk(), cam)
        ref = _scalar_update_nearby(sub.pos, sub.chunk(), cam)
        assert [tuple(p) for p in pos] == [r[0] for r in ref]                  # same set, same enumeration order
        np.testing.assert_allclose(dist, [r[1] for r in ref], rtol=1e-12)
        assert in_view.tolist() == [r[2] for r in ref] and in_gen.tolist() == [r[3] for r in ref]
        assert in_view.sum() > 0 and (in_gen & ~in_view).sum() > 0 and (~in_gen).sum() > 0
        n_checked += 1
    assert n_checked >= 5


def test_build_order_is_the_reference_pop_order():
    """The reference sorts descending by GenPrio::compare and pops from the back (world.rs:114,231-233)."""
    import functools
    rng = np.random.default_rng(3)
    n = 300
    dist, z = rng.uniform(0, 80, n), rng.integers(-2, 3, n) * 16.0
    in_gen = rng.uniform(size=n) < 0.6
    in_view = in_gen & (rng.uniform(size=n) < 0.5)

    def compare(a, b):                                   # GenPrio::compare(self=a, other=b), world.rs:40-59
        if in_view[a] and not in_view[b]: return -1
        if not in_view[a] and in_view[b]: return 1
        if in_gen[a] and not in_gen[b]: return -1
        if not in_gen[a] and in_gen[b]: return 1
        sa, sb = dist[a] ** 2 + z[a], dist[b] ** 2 + z[b]
        return -1 if sa < sb else (1 if sa > sb else 0)

    queue = sorted(range(n), key=functools.cmp_to_key(lambda a, b: compare(b, a)))   # sort_unstable_by(|1,2| 2.compare(1))
    pop_order = queue[::-1]
    assert W.gen_prio_order(dist, z, in_view, in_gen).tolist() == pop_order


class _StubBuilder:
    """No GPU: every chunk with z <= 0 gets a 3-index mesh."""
    def __init__(self):
        self.calls = []

    def build(self, positions):
        p = np.asarray(positions, dtype=np.int32).reshape(-1, 3)
        self.calls.append(len(p))
        d = np.zeros(len(p), dtype=DESC_DTYPE)
        d["pos"] = p
        has = p[:, 2] <= 0
        d["flags"] = np.where(has, CHUNK_HAS_MESH, 1)
        d["vert_count"] = np.where(has, 3, 0)
        d["index_count"] = np.where(has, 3, 0)
        d["vert_offset"] = np.concatenate([[0], np.cumsum(d["vert_count"])[:-1]])
        d["index_offset"] = d["vert_offset"]
        nv = int(d["vert_count"].sum())
        return Batch(d, np.zeros(nv, VERT_DTYPE), np.tile(np.arange(3, dtype=np.uint16), nv // 3))


def test_world_update_batches_and_recheck_rule():
    world, builder = W.World(), _StubBuilder()
    built = []
    for frame, sub, cam in W.scripted_flythrough(120, 120):
        built.append(world.update(sub, cam, builder))
    assert built[0] > 100 and world.total_count() > 0
    # recheck fires only after > 4 units of travel (1 s at MIDDLE_SPEED) or > 0.33 rad of turn (world.rs:5-6)
    fires = [i for i, b in enumerate(built) if b > 0]
    assert all(b - a >= 30 for a, b in zip(fires[:4], fires[1:5]))
    assert sum(builder.calls) == sum(built) and max(builder.calls) == built[0]
    # nothing is built twice, everything kept is within KEEP_DIST of the sub, z stays inside MIN_Z..MAX_Z
    sc = sub.chunk()
    assert all(sum((a - b) ** 2 for a, b in zip(p, sc)) < W.KEEP_DIST ** 2 for p in world.chunks)
    assert all(W.MIN_Z <= p[2] <= W.MAX_Z for p in world.chunks)
    assert all(world.chunks[p].not_blank() for p in world.chunks_to_render)
    # max_batch caps the hand-over like the reference's one-per-frame policy would with max_batch=1
    w2, b2 = W.World(), _StubBuilder()
    _, sub0, cam0 = next(iter(W.scripted_flythrough(1, 0)))
    assert w2.update(sub0, cam0, b2, max_batch=1) == 1 and w2.generate_count() > 100


def test_start_up_phase_renders_without_the_distance_filter():
    """world.rs:105-110: while should_full_build is set, build_full_step pushes EVERY not-blank chunk to
    chunks_to_render (world.rs:117-119); build_step afterwards also demands dist_sq <= (VIEW_DIST + 1)^2
    (world.rs:132-134).  The flag clears when the queue is empty or STOP_FULL_BUILD chunks exist."""
    _, sub, cam = next(iter(W.scripted_flythrough(1, 0)))
    sc = sub.chunk()
    far = lambda p: sum((a - b) ** 2 for a, b in zip(p, sc)) > (W.VIEW_DIST + 1) ** 2
    # start-up: one batch takes the whole queue -> everything with a mesh is rendered, far chunks included
    w, b = W.World(), _StubBuilder()
    assert w.should_full_build
    w.update(sub, cam, b)
    meshed = [p for p, c in w.chunks.items() if c.not_blank()]
    assert sorted(w.chunks_to_render) == sorted(meshed)
    assert any(far(p) for p in meshed), "the window must reach beyond VIEW_DIST + 1 for this test to bite"
    assert not w.should_full_build                                   # queue empty -> phase over
    # steady state: the same queue built with the flag cleared applies the distance filter
    w2, b2 = W.World(), _StubBuilder()
    w2.should_full_build = False
    w2.update(sub, cam, b2)
    assert sorted(w2.chunks_to_render) == sorted(p for p in meshed if not far(p))
    # the phase also ends once STOP_FULL_BUILD chunks exist, even with a non-empty queue (world.rs:107)
    w3, b3 = W.World(), _StubBuilder()
    w3.update(sub, cam, b3, max_batch=W.STOP_FULL_BUILD - 1)
    assert w3.should_full_build and w3.generate_count() > 0
    w3.update(sub, cam, b3, max_batch=1)
    assert not w3.should_full_build and w3.generate_count() > 0


def test_camera_matrices():
    cam = W.Camera()
    m = cam.chunk_generation_frustum_matrix(90.0)
    # a point straight ahead inside near..far is inside; one behind the eye is not (util.rs:77-85)
    assert W.in_frustum(np.array([[10.0, 0.0, 0.0]]), m)[0]
    assert not W.in_frustum(np.array([[-10.0, 0.0, 0.0]]), m)[0]
    assert not W.in_frustum(np.array([[W.Z_FAR + 5.0, 0.0, 0.0]]), m)[0]
    assert math.isclose(W.Z_FAR, 80.0)
