"""Multi-GPU region builds: slabs of the position list, meshes gathered on the rendering GPU.

Chunks are independent (chunk.rs:89-129), so the compute path has NO collective: rank r builds slab r of the
request.  The only cross-GPU traffic is the optional gather of the finished meshes to the GPU that draws, and it is
fused into the build itself (include/uwcuda.h, uw_gather_*): the rendering GPU owns the arenas, every producer's
fused kernel stores its vertices / indices / descriptors straight into its segment through NVLink peer addresses
and publishes a per-segment head; the consumer waits for the heads on its own stream.  One-sided -- no NCCL, no
rendezvous on the data path.

Two ways to drive it:

  * one process per GPU (torchrun): `RegionGather` -- rank `dst` creates the arena, the 168-byte `uw_gather_info`
    travels to the other ranks once through torch.distributed (any backend; setup only), every rank attaches its
    segment (CUDA IPC), then per build: `build()` on every rank, `wait()` on the rendering rank.
  * one process, G GPUs: `MultiBuilder` (uw_multi_*), what a single-process caller such as the reference's
    `World` (world.rs:113-123) would use.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _ffi
from ._ffi import DESC_DTYPE, VERT_DTYPE, UwError
from .chunk import ChunkBuilder, Perlin, _as_positions, INTERNAL_SIZE


def slab_bounds(n: int, parts: int, part: int) -> tuple:
    """(first, count) of contiguous slab `part` of `parts` of an n-chunk request; the remainder is spread one per
    slab.  Mirror of uw_slab_bounds (the C function is what uw_multi_build uses; tests compare the two)."""
    parts = max(1, parts)
    base, rem = divmod(n, parts)
    return part * base + min(part, rem), base + (1 if part < rem else 0)


def slab_bounds_weighted(n: int, parts: int, part: int, render_part: int, render_permille: int) -> tuple:
    """Gather-aware partition (mirror of uw_slab_bounds_weighted): slab `render_part` holds render_permille / 1000 of
    the request (never less than an even share), the other slabs share the rest evenly, contiguous and in part order."""
    if parts <= 1 or render_part >= parts or render_permille * parts <= 1000:
        return slab_bounds(n, parts, part)
    render_permille = min(render_permille, 1000)
    cr = (n * render_permille + 500) // 1000
    base, rem = divmod(n - cr, parts - 1)
    lo = cnt = k = 0
    for p in range(part + 1):
        lo += cnt
        if p == render_part:
            cnt = cr
        else:
            cnt = base + (1 if k < rem else 0)
            k += 1
    return lo, cnt


def balance_share(share: float, parts: int, t_render: float, t_other: float) -> float:
    """One step of the gather-aware balance (mirror of the library's): the rendering GPU's kernel ran t_render, the
    slowest other producer t_other (a producer whose stores wait for the rendering GPU's NVLink ingress runs longer than
    its compute alone).  Moves the share by the square root of the ratio; even split <= share <= 1/2."""
    even = 1.0 / parts
    if parts < 2 or not (t_render > 0.0) or not (t_other > 0.0):
        return share
    s = share if share > 0.0 else even
    ratio = t_other / t_render
    if 0.95 < ratio < 1.05:
        return s
    return min(0.5, max(even, s * ratio ** 0.5))


def share_search_next(state: dict, parts: int, cost: float) -> float:
    """Mirror of uw_share_search_next (what uw_multi_build runs between requests): damped hill climb on the measured
    cost of a whole request; settles on the cheapest share seen, starts over when the cost there rises by > 10 %.
    `state` = {} to start; returns the share for the next request."""
    if parts < 2:
        state["share"] = 0.0
        return 0.0
    even = 1.0 / parts
    min_step = even / 64.0
    clamp = lambda x: even if x < even else 0.5 if x > 0.5 else x

    def restart(frm):
        state.update(share=frm, step=even / 4.0, dir=1, last_cost=0.0, best_share=frm, best_cost=0.0, settled=0)

    if not state.get("share", 0.0) > 0.0:
        restart(even)
        state["moves"] = 0
    if not cost > 0.0:
        return state["share"]
    state["moves"] += 1
    if state["moves"] <= 2:                                # cold requests are not evidence
        return state["share"]
    if state["settled"]:
        if cost > state["best_cost"] * 1.10:
            restart(state["share"])
            state["last_cost"] = state["best_cost"] = cost
            state["share"] = clamp(state["share"] + state["step"])
        elif cost < state["best_cost"]:
            state["best_cost"] = cost
        return state["share"]
    if not state["best_cost"] > 0.0 or cost < state["best_cost"]:
        state["best_cost"], state["best_share"] = cost, state["share"]
    if state["last_cost"] > 0.0:
        if cost > state["last_cost"] * 1.01:
            state["dir"], state["step"] = -state["dir"], state["step"] * 0.5
        elif not cost < state["last_cost"] * 0.99:
            state["step"] *= 0.5
    state["last_cost"] = cost
    if state["step"] < min_step:
        state["settled"], state["share"] = 1, state["best_share"]
        return state["share"]
    nxt = clamp(state["share"] + state["dir"] * state["step"])
    if nxt == state["share"]:
        state["dir"] = -state["dir"]
        nxt = clamp(state["share"] + state["dir"] * state["step"])
    state["share"] = nxt
    return nxt


def info_to_bytes(info: _ffi.UwGatherInfo) -> bytes:
    return bytes(C.string_at(C.addressof(info), C.sizeof(info)))


def info_from_bytes(raw: bytes) -> _ffi.UwGatherInfo:
    if len(raw) != C.sizeof(_ffi.UwGatherInfo):
        raise ValueError("uw_gather_info blob has the wrong size")
    return _ffi.UwGatherInfo.from_buffer_copy(raw)


def broadcast_info(info: Optional[_ffi.UwGatherInfo], src: int = 0, group=None, device=None) -> _ffi.UwGatherInfo:
    """Hand rank `src`'s uw_gather_info to every rank (setup, not data path).  Works over gloo (CPU tensor) and
    NCCL (pass device="cuda")."""
    import torch
    import torch.distributed as dist
    n = C.sizeof(_ffi.UwGatherInfo)
    if dist.get_rank(group) == src:
        t = torch.frombuffer(bytearray(info_to_bytes(info)), dtype=torch.uint8).clone()
    else:
        t = torch.zeros(n, dtype=torch.uint8)
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=src, group=group)
    return info_from_bytes(t.cpu().numpy().tobytes())


class _DevMem:
    """Zero-copy view of raw device memory for torch.as_tensor (__cuda_array_interface__)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}


def device_batch_tensors(builder):
    """(descs, verts, inds) of the builder's last device-resident build as uint8 CUDA tensors that alias the
    library's arenas (no copy).  Call after builder.sync(); valid until the next build on that builder."""
    import torch
    v = builder.device_view()
    isz = 4 if v.d_inds32 else 2
    iptr = v.d_inds32 or v.d_inds16

    def wrap(ptr, nbytes):
        if nbytes == 0 or not ptr:
            return torch.empty(0, dtype=torch.uint8, device="cuda")
        return torch.as_tensor(_DevMem(ptr, nbytes), device="cuda")

    return (wrap(v.d_descs, v.n_chunks * DESC_DTYPE.itemsize), wrap(v.d_verts, v.n_verts * VERT_DTYPE.itemsize),
            wrap(iptr, v.n_inds * isz))


class GatherResult:
    """uw_gather_result with numpy conveniences."""

    def __init__(self, raw: _ffi.UwGatherResult, index_bytes: int):
        self.raw = raw
        self.n_segments, self.epoch = raw.n_segments, raw.epoch
        self.n_chunks, self.n_verts, self.n_inds, self.n_draw = raw.n_chunks, raw.n_verts, raw.n_inds, raw.n_draw
        self.seg_vcap, self.seg_icap = raw.seg_vcap, raw.seg_icap
        self.index_bytes = index_bytes
        self.segments = [dict(first_chunk=s.first_chunk, n_chunks=s.n_chunks, n_mesh=s.n_mesh, n_blank=s.n_blank,
                              n_verts=s.n_verts, n_inds=s.n_inds, guard=s.guard, overflow=s.overflow)
                         for s in raw.seg[:raw.n_segments]]

    def host_descs(self) -> Optional[np.ndarray]:
        """The pinned host copy of the descriptors (UW_GATHER_DESCS_TO_HOST) as a structured array, or None."""
        if not self.raw.h_descs:
            return None
        hi = max((s["first_chunk"] + s["n_chunks"] for s in self.segments if s["n_chunks"]), default=0)
        buf = (C.c_uint8 * (hi * DESC_DTYPE.itemsize)).from_address(self.raw.h_descs)
        return np.frombuffer(buf, dtype=DESC_DTYPE)

    def host_draw(self) -> Optional[np.ndarray]:
        """The pinned host copy of the draw list (UW_GATHER_DRAW_TO_HOST): descriptors of the chunks that ended with a
        mesh, segment by segment, completion order inside a segment.  None if it was not requested (or is empty)."""
        if not self.raw.h_draw:
            return None
        buf = (C.c_uint8 * (self.n_draw * DESC_DTYPE.itemsize)).from_address(self.raw.h_draw)
        return np.frombuffer(buf, dtype=DESC_DTYPE)

    def download(self):
        """(descs, verts, inds) of the whole arena on the host -- verification only (plain cudaMemcpy of every
        segment).  descs[i].vert_offset / index_offset index the returned verts / inds arrays."""
        lib = _ffi.load_library()
        hi = max((s["first_chunk"] + s["n_chunks"] for s in self.segments if s["n_chunks"]), default=0)
        nseg = self.n_segments

        def pull(ptr, count, dtype):
            out = np.zeros(count, dtype=dtype)
            if count:
                st = lib.uw_debug_copy_to_host(C.c_void_p(ptr), out.nbytes, out.ctypes.data)
                if st != _ffi.UW_OK:
                    raise UwError(st, "uw_debug_copy_to_host failed")
            return out

        descs = pull(self.raw.d_descs, hi, DESC_DTYPE)
        verts = pull(self.raw.d_verts, nseg * self.seg_vcap, VERT_DTYPE)
        inds = pull(self.raw.d_inds, nseg * self.seg_icap, np.uint32 if self.index_bytes == 4 else np.uint16)
        return descs, verts, inds


class RegionGather:
    """One process per GPU.  Every rank: `RegionGather(builder, rank, world, n_chunks, dst)` after
    torch.distributed is initialised (collective: broadcasts the arena info), then `build(positions_of_my_slab,
    first_chunk)` on every rank and `wait()` on rank `dst`."""

    def __init__(self, builder: ChunkBuilder, rank: int, world: int, n_chunks: int, dst: int = 0,
                 seg_vcap: int = 0, seg_icap: int = 0, group=None, bcast_device=None):
        self.builder, self.rank, self.world, self.dst = builder, rank, world, dst
        info = builder.gather_create(world, n_chunks, seg_vcap, seg_icap) if rank == dst else None
        if world > 1:
            info = broadcast_info(info, src=dst, group=group, device=bcast_device)
        self.info = info
        builder.gather_attach(info, rank)
        self.group = group
        self.render_share = 0.0                           # the rendering rank's fraction of a request; 0 = even split

    def plan(self, n: int) -> tuple:
        """(first, count) of this rank's slab of an n-chunk request under the current render share."""
        return slab_bounds_weighted(n, self.world, self.rank, self.dst, int(self.render_share * 1000.0 + 0.5))

    def feedback(self, my_kernel_seconds: float, device=None) -> float:
        """Collective (control plane, between builds): every rank reports how long its last build took; the rendering
        rank's share moves towards the point where it finishes together with the slowest producer."""
        if self.world < 2:
            return self.render_share
        import torch
        import torch.distributed as dist
        t = torch.zeros(self.world, dtype=torch.float64, device=device or "cpu")
        t[self.rank] = my_kernel_seconds
        dist.all_reduce(t, group=self.group)
        ts = t.cpu().tolist()
        others = max(x for r, x in enumerate(ts) if r != self.dst)
        self.render_share = balance_share(self.render_share, self.world, ts[self.dst], others)
        return self.render_share

    def candidate_shares(self, steps: int = 7, growth: float = 0.25) -> list:
        """Render shares worth measuring: the even split and `steps - 1` larger ones (the rendering rank's own output does
        not cross NVLink, so its slab only ever grows), capped at one half."""
        even = 1.0 / max(self.world, 1)
        return sorted({min(0.5, even * (1.0 + growth * k)) for k in range(steps)})

    def tune(self, run_step, candidates=None, repeats: int = 2, cold: int = 2, device=None, refine: bool = True) -> float:
        """Collective (control plane, warm-up only): measure whole steps under each candidate render share and keep the
        fastest.  `run_step()` runs ONE complete request under the current share on this rank (plan, build, the
        rendering rank's wait, sync) and returns its seconds; the step time is the maximum over ranks.  The one-step
        controller (`feedback`) equalises kernel times, which stops short of the optimum when the rendering GPU's own
        kernel is slowed by the traffic arriving over NVLink; a direct search over a handful of shares does not care."""
        if self.world < 2:
            return self.render_share
        import torch
        import torch.distributed as dist
        cands = list(candidates) if candidates is not None else self.candidate_shares()
        best, best_t = self.render_share, float("inf")
        self.tune_trace = []                                   # [(share, seconds)]: what the search saw
        self.render_share = cands[0]
        for _ in range(cold):
            run_step()

        def measure(sh):
            nonlocal best, best_t
            self.render_share = sh
            t_min = float("inf")
            for _ in range(repeats):
                t = torch.tensor([run_step()], dtype=torch.float64, device=device or "cpu")
                dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
                t_min = min(t_min, float(t.cpu()[0]))
            self.tune_trace.append((sh, t_min))
            if t_min < best_t:
                best, best_t = sh, t_min

        for sh in cands:
            measure(sh)
        if refine and len(cands) > 1:                          # one refinement pass: half a grid step either side of the best
            half = 0.5 * min(b - a for a, b in zip(cands, cands[1:]))
            centre = best
            for sh in (centre - half, centre + half):
                if cands[0] <= sh <= 0.5 and all(abs(sh - c) > 1e-9 for c in cands):
                    measure(sh)
        self.render_share = best
        return best

    def build(self, positions, first_chunk: int):
        self.builder.gather_build(positions, first_chunk)

    def build_device(self, d_positions_ptr: int, n: int, first_chunk: int):
        self.builder.gather_build_device(d_positions_ptr, n, first_chunk)

    def wait(self, descs_to_host: bool = False, draw_to_host: bool = False) -> GatherResult:
        if self.rank != self.dst:
            raise RuntimeError("RegionGather.wait() belongs to the rendering rank")
        return self.builder.gather_wait(descs_to_host, draw_to_host)

    def close(self):
        self.builder.gather_detach()
        if self.rank == self.dst:
            self.builder.gather_destroy()


class MultiBuilder:
    """uw_multi_*: one process drives G GPUs; devices[0] renders."""

    def __init__(self, perlin: Optional[Perlin] = None, devices: Sequence[int] = (0,), *, internal_size: int = INTERNAL_SIZE,
                 index32: bool = False):
        self._lib = _ffi.load_library()
        cfg = _ffi.UwConfig()
        self._lib.uw_config_default(C.byref(cfg))
        cfg.internal_size = internal_size
        cfg.seed = (perlin or Perlin()).seed()
        cfg.flags = _ffi.FLAG_INDEX32 if index32 else 0
        self.index_bytes = 4 if index32 else 2
        devs = (C.c_int32 * len(devices))(*devices)
        self._m = C.c_void_p()
        st = self._lib.uw_multi_create(C.byref(cfg), devs, len(devices), C.byref(self._m))
        if st != _ffi.UW_OK:
            raise UwError(st, (self._lib.uw_multi_last_error(None) or b"").decode())
        self.devices = list(devices)

    def build(self, positions, descs_to_host: bool = False, draw_to_host: bool = False) -> GatherResult:
        p = _as_positions(positions)
        self._keep = p
        out = _ffi.UwGatherResult()
        flags = (_ffi.GATHER_DESCS_TO_HOST if descs_to_host else 0) | (_ffi.GATHER_DRAW_TO_HOST if draw_to_host else 0)
        st = self._lib.uw_multi_build(self._m, p.ctypes.data, p.shape[0], flags, C.byref(out))
        if st != _ffi.UW_OK:
            raise UwError(st, (self._lib.uw_multi_last_error(self._m) or b"").decode())
        return GatherResult(out, self.index_bytes)

    def close(self):
        if getattr(self, "_m", None) and self._m.value:
            self._lib.uw_multi_destroy(self._m)
            self._m = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
