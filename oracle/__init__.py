"""CPU oracle for the chunk-build hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``underwaterworld_b200``) never does; it fails loudly when its CUDA library is missing.

``Oracle`` is a thin ctypes binding over ``oracle/uw_oracle.cpp`` (see that file's header
for what it restates and how it is pinned).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libuw_oracle.so")

FLAG_BLANK_EARLY = 1
FLAG_HAS_MESH = 2
FLAG_U16_OVERFLOW = 4

MODE_FAITHFUL = 0
MODE_FAST = 1
MODE_RULE = 2

VERT_DTYPE = np.dtype([("pos", "<f4", (3,)), ("color", "<f4", (3,))])
TRI_DTYPE = np.dtype([("verts", "<f4", (3, 3)), ("normal", "<f4", (3,))])


class OracleConfig(C.Structure):
    _fields_ = [
        ("internal_size", C.c_int32),
        ("chunk_size", C.c_int32),
        ("octaves", C.c_uint32),
        ("iso_level", C.c_float),
        ("max_height", C.c_float),
        ("adj_z_mod", C.c_float),
        ("min_hue", C.c_float),
        ("max_hue", C.c_float),
        ("saturation", C.c_float),
        ("base_value", C.c_float),
        ("min_z", C.c_float),
        ("max_z", C.c_float),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with g++ (no GPU needed).  Returns the library path."""
    src = os.path.join(_HERE, "uw_oracle.cpp")
    tab = os.path.join(_HERE, "mc_tables_oracle.h")
    stale = (
        force
        or not os.path.exists(_LIB_PATH)
        or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(tab))
    )
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


class Oracle:
    def __init__(self, internal_size: int = 12, **overrides):
        self.lib = C.CDLL(build())
        L = self.lib
        L.uwo_config_default.argtypes = [C.POINTER(OracleConfig)]
        L.uwo_perm_table.argtypes = [C.c_uint32, C.c_void_p]
        L.uwo_perlin3.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        L.uwo_perlin3.restype = C.c_double
        L.uwo_perlin3_octaves.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_uint32]
        L.uwo_perlin3_octaves.restype = C.c_double
        L.uwo_iso_at.argtypes = [C.POINTER(OracleConfig), C.c_void_p, C.c_double, C.c_double, C.c_double]
        L.uwo_iso_at.restype = C.c_float
        L.uwo_densities.argtypes = [C.POINTER(OracleConfig), C.c_void_p, C.c_void_p, C.c_void_p]
        L.uwo_hsv_to_rgb.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.uwo_to_srgb.argtypes = [C.c_void_p, C.c_void_p]
        L.uwo_vertex_color.argtypes = [C.POINTER(OracleConfig), C.c_float, C.c_uint32, C.c_void_p]
        L.uwo_build_chunk.argtypes = [
            C.POINTER(OracleConfig), C.c_void_p, C.c_void_p, C.c_int,
            C.c_void_p, C.c_void_p, C.c_void_p,
            C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
            C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
            C.c_void_p, C.c_uint32, C.c_void_p,
        ]
        L.uwo_build_chunk.restype = C.c_int
        L.uwo_vertex_pairs.argtypes = [C.POINTER(OracleConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.uwo_vertex_pairs.restype = C.c_int
        L.uwo_raycast.argtypes = [C.POINTER(OracleConfig), C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                  C.c_uint32, C.c_int32, C.c_void_p]
        L.uwo_raycast.restype = None
        L.uwo_build_batch_timed.argtypes = [
            C.POINTER(OracleConfig), C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int,
            C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
            C.POINTER(C.c_double),
        ]
        L.uwo_build_batch_timed.restype = C.c_double
        self.cfg = OracleConfig()
        L.uwo_config_default(C.byref(self.cfg))
        self.cfg.internal_size = internal_size
        for k, v in overrides.items():
            setattr(self.cfg, k, v)

    # -- noise -------------------------------------------------------------
    def perm_table(self, seed: int) -> np.ndarray:
        out = np.zeros(256, dtype=np.uint8)
        self.lib.uwo_perm_table(seed & 0xFFFFFFFF, out.ctypes.data)
        return out

    def perlin3(self, perm: np.ndarray, x: float, y: float, z: float) -> float:
        perm = np.ascontiguousarray(perm, dtype=np.uint8)
        return self.lib.uwo_perlin3(perm.ctypes.data, x, y, z)

    def perlin3_octaves(self, perm, x, y, z, octaves=3) -> float:
        perm = np.ascontiguousarray(perm, dtype=np.uint8)
        return self.lib.uwo_perlin3_octaves(perm.ctypes.data, x, y, z, octaves)

    def iso_at(self, perm, x, y, z) -> float:
        perm = np.ascontiguousarray(perm, dtype=np.uint8)
        return self.lib.uwo_iso_at(C.byref(self.cfg), perm.ctypes.data, x, y, z)

    def densities(self, perm, pos) -> np.ndarray:
        perm = np.ascontiguousarray(perm, dtype=np.uint8)
        p = np.asarray(pos, dtype=np.int32)
        Ls = self.cfg.internal_size + 1
        out = np.empty(Ls ** 3, dtype=np.float32)
        self.lib.uwo_densities(C.byref(self.cfg), perm.ctypes.data, p.ctypes.data, out.ctypes.data)
        return out

    # -- colour --------------------------------------------------------------
    def hsv_to_rgb(self, h, s, v) -> np.ndarray:
        out = np.empty(3, dtype=np.float32)
        self.lib.uwo_hsv_to_rgb(h, s, v, out.ctypes.data)
        return out

    def to_srgb(self, rgb) -> np.ndarray:
        a = np.ascontiguousarray(rgb, dtype=np.float32)
        out = np.empty(3, dtype=np.float32)
        self.lib.uwo_to_srgb(a.ctypes.data, out.ctypes.data)
        return out

    def vertex_color(self, world_z, corner_b_idx) -> np.ndarray:
        out = np.empty(3, dtype=np.float32)
        self.lib.uwo_vertex_color(C.byref(self.cfg), world_z, corner_b_idx, out.ctypes.data)
        return out

    # -- chunk ---------------------------------------------------------------
    def build_chunk(self, perm, pos, mode=MODE_FAITHFUL, isos=None, want_tris=False):
        """Chunk::build_full.  Returns dict(isos, cases, verts, inds (u32), flags, [tris, tri_cell_start])."""
        perm = np.ascontiguousarray(perm, dtype=np.uint8)
        p = np.asarray(pos, dtype=np.int32)
        S = self.cfg.internal_size
        Ls = S + 1
        isos_out = np.empty(Ls ** 3, dtype=np.float32)
        cases = np.zeros(S ** 3, dtype=np.uint8)
        vcap = 5 * S * Ls * Ls + 16
        icap = 15 * S ** 3
        verts = np.zeros(vcap, dtype=VERT_DTYPE)
        inds = np.zeros(icap, dtype=np.uint32)
        nv, ni, fl = C.c_uint32(), C.c_uint32(), C.c_uint32()
        isos_in = None
        if isos is not None:
            isos_in = np.ascontiguousarray(isos, dtype=np.float32)
            assert isos_in.size == Ls ** 3
        tris = tcs = None
        if want_tris:
            tris = np.zeros(5 * S ** 3, dtype=TRI_DTYPE)
            tcs = np.zeros(S ** 3 + 1, dtype=np.uint32)
        rc = self.lib.uwo_build_chunk(
            C.byref(self.cfg), perm.ctypes.data, p.ctypes.data, mode,
            isos_in.ctypes.data if isos_in is not None else None,
            isos_out.ctypes.data, cases.ctypes.data,
            verts.ctypes.data, vcap, inds.ctypes.data, icap,
            C.byref(nv), C.byref(ni), C.byref(fl),
            tris.ctypes.data if want_tris else None, 5 * S ** 3 if want_tris else 0,
            tcs.ctypes.data if want_tris else None,
        )
        assert rc == 0
        r = dict(isos=isos_out, cases=cases, verts=verts[: nv.value].copy(), inds=inds[: ni.value].copy(),
                 flags=fl.value)
        if want_tris:
            r["tri_cell_start"] = tcs
            r["tris"] = tris[: tcs[-1]].copy()
        return r

    def vertex_pairs(self, perm, pos, isos=None) -> np.ndarray:
        """(n_verts, 2) lattice indices (x*L*L + y*L + z) of every vertex's ordered corner pair (corner_a, corner_b),
        in vertex order -- Build::vert_pairs (chunk.rs:38)."""
        perm = np.ascontiguousarray(perm, dtype=np.uint8)
        p = np.asarray(pos, dtype=np.int32)
        S = self.cfg.internal_size
        cap = 5 * S * (S + 1) ** 2 + 16
        out = np.zeros((cap, 2), dtype=np.uint32)
        isos_in = None if isos is None else np.ascontiguousarray(isos, dtype=np.float32)
        n = self.lib.uwo_vertex_pairs(C.byref(self.cfg), perm.ctypes.data, p.ctypes.data,
                                      isos_in.ctypes.data if isos_in is not None else None, out.ctypes.data, cap)
        assert n >= 0
        return out[:n].copy()

    def raycast(self, perm, chunk_positions, origins, dirs, wall_range=3) -> np.ndarray:
        """Tri::intersects (util.rs:22-59) over the triangles boid.rs:175-208 gathers around every ray origin, out of
        the given chunks: the smallest hit distance per ray, -1 where nothing is hit."""
        perm = np.ascontiguousarray(perm, dtype=np.uint8)
        cp = np.ascontiguousarray(chunk_positions, dtype=np.int32).reshape(-1, 3)
        o = np.ascontiguousarray(origins, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, 3)
        out = np.empty(len(o), dtype=np.float32)
        self.lib.uwo_raycast(C.byref(self.cfg), perm.ctypes.data, cp.ctypes.data, len(cp), o.ctypes.data, d.ctypes.data,
                             len(o), wall_range, out.ctypes.data)
        return out

    def build_batch_timed(self, perm, positions, mode=MODE_FAITHFUL, nthreads=1):
        perm = np.ascontiguousarray(perm, dtype=np.uint8)
        p = np.ascontiguousarray(positions, dtype=np.int32).reshape(-1, 3)
        tv, ti, nb, nm = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        cs = C.c_double()
        secs = self.lib.uwo_build_batch_timed(
            C.byref(self.cfg), perm.ctypes.data, p.ctypes.data, p.shape[0], mode, nthreads,
            C.byref(tv), C.byref(ti), C.byref(nb), C.byref(nm), C.byref(cs))
        return dict(seconds=secs, n=p.shape[0], verts=tv.value, inds=ti.value, blank=nb.value,
                    mesh=nm.value, checksum=cs.value)
