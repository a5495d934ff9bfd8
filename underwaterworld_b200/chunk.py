"""Host-side mirror of the reference's chunk-build interface, on top of the C ABI.

Mirrors (names, argument meaning, error behaviour) the Rust API of
``underwater_world/src/chunk.rs``:

    Chunk::new(pos)                      chunk.rs:89      -> Chunk(pos)
    Chunk::build_full(&perlin, &device)  chunk.rs:266     -> Chunk.build_full(builder)
    Chunk::build_partial(..) -> bool     chunk.rs:270     -> Chunk.build_partial(builder)
    Chunk::not_blank()                   chunk.rs:344     -> Chunk.not_blank()
    Chunk::verts_buffer_slice()          chunk.rs:346     -> Chunk.verts_buffer_slice()
    Chunk::inds_buffer_slice()           chunk.rs:347     -> Chunk.inds_buffer_slice()
    Chunk::num_inds()                    chunk.rs:348     -> Chunk.num_inds()
    noise::Perlin::new(seed)             state.rs:359     -> Perlin(seed)

plus the batched entry the GPU needs (``ChunkBuilder.build(positions)``): the reference builds one
chunk per frame (world.rs:113-145); a GPU wants thousands per call.

Everything computes through libuwcuda.so.  There is no CPU fallback: constructing a
``ChunkBuilder`` without the built library or without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Iterable, Optional, Sequence

import numpy as np

from . import _ffi
from ._ffi import (CHUNK_BLANK_EARLY, CHUNK_HAS_MESH, CHUNK_U16_OVERFLOW, DESC_DTYPE, TRI_DTYPE, VERT_DTYPE, UwError)

# reference constants, chunk.rs:5-12
CHUNK_SIZE = 16
INTERNAL_SIZE = 12
PERLIN_OCTAVES = 3
ISO_LEVEL = -0.1


class Perlin:
    """noise::Perlin::new(seed) (state.rs:359).  Only the seed matters: the permutation table is
    rebuilt inside the library (uw_perm_table)."""

    DEFAULT_SEED = 0

    def __init__(self, seed: int = 0):
        self._seed = int(seed) & 0xFFFFFFFF

    def seed(self) -> int:   # noise::Seedable::seed
        return self._seed


def _as_positions(positions) -> np.ndarray:
    p = np.ascontiguousarray(positions, dtype=np.int32)
    if p.ndim == 1:
        if p.size % 3:
            raise ValueError("positions must be n x 3")
        p = p.reshape(-1, 3)
    if p.ndim != 2 or p.shape[1] != 3:
        raise ValueError("positions must be n x 3")
    return p


@dataclass
class ChunkMesh:
    """One chunk's result: what Chunk::build_full leaves behind (chunk.rs:291-305)."""
    pos: tuple
    flags: int
    verts: np.ndarray          # VERT_DTYPE [vert_count]  == draw::VertColor
    inds: np.ndarray           # uint16 (or uint32) [index_count], chunk-local
    tris: Optional[np.ndarray] = None            # TRI_DTYPE [index_count / 3]   (UW_FLAG_TRIS)
    tri_cell_start: Optional[np.ndarray] = None  # uint16 [S^3 + 1], first triangle of every cell (scan order)

    @property
    def blank_early(self) -> bool:
        return bool(self.flags & CHUNK_BLANK_EARLY)

    def not_blank(self) -> bool:
        return bool(self.flags & CHUNK_HAS_MESH)

    def num_inds(self) -> int:
        return int(self.inds.shape[0])


class Batch:
    """Finished batch (host copies of the packed buffers)."""

    def __init__(self, descs: np.ndarray, verts: np.ndarray, inds: np.ndarray, tris=None, tri_cell_start=None):
        self.descs, self.verts, self.inds = descs, verts, inds
        self.tris, self.tri_cell_start = tris, tri_cell_start

    def __len__(self) -> int:
        return int(self.descs.shape[0])

    @property
    def n_verts(self) -> int:
        return int(self.verts.shape[0])

    @property
    def n_inds(self) -> int:
        return int(self.inds.shape[0])

    def compact(self):
        """(verts, inds) with every chunk's buffers concatenated tightly in chunk order.  The packed arrays themselves
        start every chunk on a 16-byte boundary (allocations are padded to an even vertex count / a multiple of 16
        bytes of indices), so they hold a few unused pad entries between chunks."""
        def ranges(off, cnt):
            cnt = cnt.astype(np.int64)
            tot = int(cnt.sum())
            start = np.repeat(off.astype(np.int64) - (np.cumsum(cnt) - cnt), cnt)
            return start + np.arange(tot, dtype=np.int64)
        d = self.descs
        return self.verts[ranges(d["vert_offset"], d["vert_count"])], self.inds[ranges(d["index_offset"], d["index_count"])]

    def chunk(self, i: int) -> ChunkMesh:
        d = self.descs[i]
        vo, vc, io, ic = int(d["vert_offset"]), int(d["vert_count"]), int(d["index_offset"]), int(d["index_count"])
        tris = tcs = None
        if self.tris is not None:
            tris, tcs = self.tris[io // 3:(io + ic) // 3], self.tri_cell_start[i]
        return ChunkMesh(tuple(int(v) for v in d["pos"]), int(d["flags"]), self.verts[vo:vo + vc], self.inds[io:io + ic], tris, tcs)

    def __iter__(self):
        return (self.chunk(i) for i in range(len(self)))


class ChunkBuilder:
    """Owns one uw_ctx (one CUDA device, one stream).  Not thread-safe (like the reference)."""

    def __init__(self, perlin: Optional[Perlin] = None, *, internal_size: int = INTERNAL_SIZE, device: int = -1,
                 exact_f64: bool = False, index32: bool = False, keep_densities: bool = False,
                 staged: bool = False, ordered: bool = False, tris: bool = False, analytic_skip: bool = False, exportable: bool = False,
                 guard_eps: float = 0.0, **consts):
        self._lib = _ffi.load_library()
        cfg = _ffi.UwConfig()
        self._lib.uw_config_default(C.byref(cfg))
        cfg.internal_size = internal_size
        cfg.seed = (perlin or Perlin()).seed()
        cfg.device = device
        cfg.guard_eps = guard_eps
        cfg.flags = ((_ffi.FLAG_EXACT_F64 if exact_f64 else 0) | (_ffi.FLAG_INDEX32 if index32 else 0)
                     | (_ffi.FLAG_KEEP_DENSITIES if keep_densities else 0) | (_ffi.FLAG_STAGED if staged else 0)
                     | (_ffi.FLAG_ORDERED if ordered else 0) | (_ffi.FLAG_TRIS if tris else 0)
                     | (_ffi.FLAG_ANALYTIC_SKIP if analytic_skip else 0) | (_ffi.FLAG_EXPORTABLE if exportable else 0))
        for k, v in consts.items():
            if not hasattr(cfg, k):
                raise TypeError(f"unknown config field {k!r}")
            setattr(cfg, k, v)
        self.cfg = cfg
        self._ctx = C.c_void_p()
        st = self._lib.uw_create(C.byref(cfg), C.byref(self._ctx))
        if st != _ffi.UW_OK:
            raise UwError(st, (self._lib.uw_last_error(None) or b"").decode())
        self.S = internal_size
        self.L = internal_size + 1
        self.index32 = index32

    # -- plumbing --------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.uw_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, st: int):
        if st != _ffi.UW_OK:
            raise UwError(st, (self._lib.uw_last_error(self._ctx) or b"").decode())

    def perm_table(self) -> np.ndarray:
        out = np.zeros(256, dtype=np.uint8)
        self._check(self._lib.uw_perm_table(self._ctx, out.ctypes.data))
        return out

    def set_stream(self, cuda_stream: int):
        self._check(self._lib.uw_set_stream(self._ctx, C.c_void_p(cuda_stream)))

    def set_profiling(self, enabled: bool):
        self._check(self._lib.uw_set_profiling(self._ctx, int(enabled)))

    def stage_times(self) -> dict:
        t = _ffi.UwStageTimes()
        self._check(self._lib.uw_get_stage_times(self._ctx, C.byref(t)))
        return {k: getattr(t, k) for k, _ in t._fields_}

    def ffma_peak_tflops(self) -> float:
        """Measured FFMA-chain throughput of the device (measurement aid for the FP32 roofline)."""
        v = C.c_double()
        self._check(self._lib.uw_debug_ffma_peak(self._ctx, C.byref(v)))
        return v.value

    def guard_count(self) -> int:
        v = C.c_uint64()
        self._check(self._lib.uw_get_guard_count(self._ctx, C.byref(v)))
        return v.value

    # -- the hot path ------------------------------------------------------------------------
    def _collect(self, handle) -> Batch:
        view = _ffi.UwBatchView()
        try:
            self._check(self._lib.uw_batch_view_get(handle, C.byref(view)))
            n, nv, ni = view.n_chunks, view.n_verts, view.n_inds
            descs = np.empty(n, dtype=DESC_DTYPE)
            verts = np.empty(nv, dtype=VERT_DTYPE)
            if n:
                C.memmove(descs.ctypes.data, view.descs, n * DESC_DTYPE.itemsize)
            if nv:
                C.memmove(verts.ctypes.data, view.verts, nv * VERT_DTYPE.itemsize)
            if view.inds32:
                inds = np.empty(ni, dtype=np.uint32)
                if ni:
                    C.memmove(inds.ctypes.data, view.inds32, ni * 4)
            else:
                inds = np.empty(ni, dtype=np.uint16)
                if ni:
                    C.memmove(inds.ctypes.data, view.inds16, ni * 2)
            tris = tcs = None
            if view.tris:
                tris = np.empty(ni // 3, dtype=TRI_DTYPE)
                if ni:
                    C.memmove(tris.ctypes.data, view.tris, (ni // 3) * TRI_DTYPE.itemsize)
                tcs = np.empty((n, self.S ** 3 + 1), dtype=np.uint16)
                if n:
                    C.memmove(tcs.ctypes.data, view.tri_cell_start, tcs.nbytes)
        finally:
            self._lib.uw_batch_free(handle)
        return Batch(descs, verts, inds, tris, tcs)

    def build(self, positions) -> Batch:
        """Chunk::new + Chunk::build_full for every position (host in, host out)."""
        p = _as_positions(positions)
        h = C.c_void_p()
        self._check(self._lib.uw_build(self._ctx, p.ctypes.data, p.shape[0], C.byref(h)))
        return self._collect(h)

    def build_async(self, positions):
        p = _as_positions(positions)
        h = C.c_void_p()
        self._check(self._lib.uw_build_async(self._ctx, p.ctypes.data, p.shape[0], C.byref(h)))
        return h

    def wait(self, handle) -> Batch:
        self._check(self._lib.uw_batch_wait(handle))
        return self._collect(handle)

    def build_stream(self, batches):
        """Generator over an iterable of position arrays: yields one Batch per input, in order, keeping TWO batches
        in flight (uw_build_async for batch k+1 before uw_batch_wait for batch k), so that batch k's copy to the
        host runs underneath batch k+1's kernel.  The streaming form of World::build_full_step for a loader that
        hands over a whole region in slices."""
        prev = None
        for positions in batches:
            nxt = self.build_async(positions)
            if prev is not None:
                yield self.wait(prev)
            prev = nxt
        if prev is not None:
            yield self.wait(prev)

    def build_from_densities(self, positions, densities) -> Batch:
        p = _as_positions(positions)
        d = np.ascontiguousarray(densities, dtype=np.float32).reshape(p.shape[0], -1)
        if d.shape[1] != self.L ** 3:
            raise ValueError("densities must be n x L^3")
        h = C.c_void_p()
        self._check(self._lib.uw_build_from_densities(self._ctx, p.ctypes.data, d.ctypes.data, p.shape[0], C.byref(h)))
        return self._collect(h)

    def build_device(self, d_positions_ptr: int, n: int):
        """Device-resident build: positions at a device pointer (n x 3 int32); no sync."""
        self._check(self._lib.uw_build_device(self._ctx, C.c_void_p(d_positions_ptr), n))

    def export_arena_fd(self, which: int):
        """(fd, allocation bytes) of the packed vertex (0) / index (1) arena of the last device-resident build:
        a POSIX file descriptor a renderer can import (builder created with exportable=True)."""
        fd, nbytes = C.c_int(-1), C.c_uint64(0)
        self._check(self._lib.uw_export_arena_fd(self._ctx, which, C.byref(fd), C.byref(nbytes)))
        return fd.value, nbytes.value

    def sync(self):
        self._check(self._lib.uw_sync(self._ctx))

    # -- multi-GPU gather (include/uwcuda.h uw_gather_*; see gather.py) ---------------------------
    def gather_create(self, n_segments: int, n_chunks: int, seg_vcap: int = 0, seg_icap: int = 0) -> _ffi.UwGatherInfo:
        """This builder's GPU becomes the rendering side: arenas for n_chunks chunks in n_segments segments."""
        info = _ffi.UwGatherInfo()
        self._check(self._lib.uw_gather_create(self._ctx, n_segments, n_chunks, seg_vcap, seg_icap, C.byref(info)))
        return info

    def gather_destroy(self):
        self._check(self._lib.uw_gather_destroy(self._ctx))

    def gather_attach(self, info: _ffi.UwGatherInfo, segment: int):
        self._check(self._lib.uw_gather_attach(self._ctx, C.byref(info), segment))

    def gather_detach(self):
        self._check(self._lib.uw_gather_detach(self._ctx))

    def gather_build(self, positions, first_chunk: int):
        """Chunk::new + build_full for the positions (host) into the attached segment; asynchronous."""
        p = _as_positions(positions)
        self._gather_pos = p                      # the library copies into its pinned staging before returning
        self._check(self._lib.uw_gather_build(self._ctx, p.ctypes.data, p.shape[0], first_chunk))

    def gather_build_device(self, d_positions_ptr: int, n: int, first_chunk: int):
        self._check(self._lib.uw_gather_build_device(self._ctx, C.c_void_p(d_positions_ptr), n, first_chunk))

    def gather_wait(self, descs_to_host: bool = False, draw_to_host: bool = False):
        from .gather import GatherResult
        out = _ffi.UwGatherResult()
        flags = (_ffi.GATHER_DESCS_TO_HOST if descs_to_host else 0) | (_ffi.GATHER_DRAW_TO_HOST if draw_to_host else 0)
        self._check(self._lib.uw_gather_wait(self._ctx, flags, C.byref(out)))
        return GatherResult(out, 4 if self.index32 else 2)

    def device_view(self) -> _ffi.UwDeviceView:
        v = _ffi.UwDeviceView()
        self._check(self._lib.uw_device_view_get(self._ctx, C.byref(v)))
        return v

    # -- parity taps -------------------------------------------------------------------------
    def debug_densities(self, positions) -> np.ndarray:
        p = _as_positions(positions)
        out = np.empty((p.shape[0], self.L ** 3), dtype=np.float32)
        self._check(self._lib.uw_debug_densities(self._ctx, p.ctypes.data, p.shape[0], out.ctypes.data))
        return out

    def debug_vertex_colors(self, world_z, level) -> np.ndarray:
        """The kernels' vertex colour (chunk.rs:215-222) for (world z, corner_b index % 3) pairs -> (n, 3) float32."""
        z = np.ascontiguousarray(world_z, dtype=np.float32).reshape(-1)
        lv = np.ascontiguousarray(np.broadcast_to(np.asarray(level, dtype=np.uint32), z.shape))
        out = np.empty((z.shape[0], 3), dtype=np.float32)
        self._check(self._lib.uw_debug_vertex_colors(self._ctx, z.ctypes.data, lv.ctypes.data, z.shape[0], out.ctypes.data))
        return out

    def debug_cases(self, positions) -> np.ndarray:
        p = _as_positions(positions)
        out = np.empty((p.shape[0], self.S ** 3), dtype=np.uint8)
        self._check(self._lib.uw_debug_cases(self._ctx, p.ctypes.data, p.shape[0], out.ctypes.data))
        return out

    def raycast_tris(self, origins, dirs, wall_range: int = 3) -> np.ndarray:
        """Batched Tri::intersects (util.rs:22-59) the way boid.rs:175-240 uses it, against the collision triangles of
        this builder's LAST build (ChunkBuilder(tris=True)): per ray the smallest hit distance t, -1 where no candidate
        triangle is hit.  `heading for a collision` = (0 <= t < wall_range); `direction is safe` = (t == -1)."""
        o = np.ascontiguousarray(origins, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, 3)
        if o.shape != d.shape:
            raise ValueError("origins and dirs must have the same shape")
        out = np.empty(o.shape[0], dtype=np.float32)
        self._check(self._lib.uw_raycast_tris(self._ctx, o.ctypes.data, d.ctypes.data, o.shape[0], wall_range, out.ctypes.data))
        return out

    def iso_at(self, points) -> np.ndarray:
        """perlin_util::iso_at on n f64 points (perlin_util.rs:24-29)."""
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        out = np.empty(pts.shape[0], dtype=np.float32)
        self._check(self._lib.uw_iso_at(self._ctx, pts.ctypes.data, pts.shape[0], out.ctypes.data))
        return out


class Chunk:
    """Drop-in shape of the reference's `Chunk` (chunk.rs:80-349) for single-chunk callers.

    `build_full(builder)` replaces `build_full(&perlin, &device)`: the builder carries the seed
    (Perlin) and the device.  `build_partial` completes the chunk in one step and returns True
    (the reference's frame-slicing, chunk.rs:19-20, exists only to keep a single CPU thread
    responsive; a batched GPU build has no partial state).
    """

    def __init__(self, pos: Sequence[int]):
        self.pos = (int(pos[0]), int(pos[1]), int(pos[2]))
        self.chunk_offset = tuple(p * CHUNK_SIZE for p in self.pos)   # chunk.rs:90-94
        self._mesh: Optional[ChunkMesh] = None

    @classmethod
    def new(cls, pos):
        return cls(pos)

    def build_full(self, builder: ChunkBuilder) -> None:
        self._mesh = builder.build([self.pos]).chunk(0)

    def build_partial(self, builder: ChunkBuilder) -> bool:
        if self._mesh is None:
            self.build_full(builder)
        return True

    def _adopt(self, mesh: ChunkMesh) -> "Chunk":
        self._mesh = mesh
        return self

    def not_blank(self) -> bool:
        return self._mesh is not None and self._mesh.not_blank()

    def verts_buffer_slice(self) -> np.ndarray:
        if not self.not_blank():   # the reference unwraps a None buffer here -> panic (chunk.rs:346)
            raise RuntimeError("called verts_buffer_slice() on a blank chunk")
        return self._mesh.verts

    def inds_buffer_slice(self) -> np.ndarray:
        if not self.not_blank():
            raise RuntimeError("called inds_buffer_slice() on a blank chunk")
        return self._mesh.inds

    def num_inds(self) -> int:
        return 0 if self._mesh is None else self._mesh.num_inds()

    def tris_around(self, local_pos_percent: Sequence[float], rng: int) -> np.ndarray:
        """Chunk::tris_around (chunk.rs:315-342): the collision triangles of every cell within `rng` cells of
        the cell containing local_pos_percent (each component in [0,1)).  Needs ChunkBuilder(tris=True)."""
        m = self._mesh
        if m is None or m.tris is None:
            raise RuntimeError("tris_around needs a chunk built with ChunkBuilder(tris=True)")
        S = round((len(m.tri_cell_start) - 1) ** (1.0 / 3.0))
        mid = [int(np.floor(np.float32(v) * np.float32(S))) for v in local_pos_percent]
        lo = [max(c - rng, 0) for c in mid]
        hi = [min(c + rng, S) for c in mid]                      # inclusive, like the reference (cells == S hold nothing)
        out = []
        for x in range(lo[0], hi[0] + 1):
            for y in range(lo[1], hi[1] + 1):
                for z in range(lo[2], hi[2] + 1):
                    if x < S and y < S and z < S:
                        c = (x * S + y) * S + z
                        out.append(m.tris[int(m.tri_cell_start[c]):int(m.tri_cell_start[c + 1])])
        return np.concatenate(out) if out else np.zeros(0, dtype=TRI_DTYPE)


def build_chunks(builder: ChunkBuilder, positions: Iterable[Sequence[int]]) -> list:
    """Batched `World::build_full_step` (world.rs:113-123): one call, many chunks."""
    pos = _as_positions(list(positions))
    batch = builder.build(pos)
    return [Chunk(tuple(int(v) for v in pos[i]))._adopt(batch.chunk(i)) for i in range(len(batch))]
