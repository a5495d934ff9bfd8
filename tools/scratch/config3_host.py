"""BASELINE config 3 through the HOST path: 524288 chunks streamed to pinned host memory in slices (two in flight)."""
import sys, os, time, ctypes as C, numpy as np
sys.path.insert(0, os.getcwd())
import underwaterworld_b200 as uw
from underwaterworld_b200 import _ffi
lib = uw.load_library()
pos = uw.region.config_positions("large")
b = uw.ChunkBuilder(uw.Perlin(0))
ctx = b._ctx
view = _ffi.UwBatchView()
for sl in (2048, 8192, 32768):
    parts = [np.ascontiguousarray(pos[i:i + sl]) for i in range(0, len(pos), sl)]
    def run():
        nv = ni = 0
        prev = None
        for p in parts:
            h = C.c_void_p()
            assert lib.uw_build_async(ctx, p.ctypes.data, len(p), C.byref(h)) == 0, lib.uw_last_error(ctx)
            if prev is not None:
                assert lib.uw_batch_wait(prev) == 0
                lib.uw_batch_view_get(prev, C.byref(view)); nv += view.n_verts; ni += view.n_inds
                lib.uw_batch_free(prev)
            prev = h
        assert lib.uw_batch_wait(prev) == 0
        lib.uw_batch_view_get(prev, C.byref(view)); nv += view.n_verts; ni += view.n_inds
        lib.uw_batch_free(prev)
        return nv, ni
    run()
    t0 = time.perf_counter(); nv, ni = run(); dt = time.perf_counter() - t0
    mb = (nv * 24 + ni * 2 + len(pos) * 32) / 1e6
    print(f"slices of {sl:6d}: {dt*1e3:7.1f} ms  {len(pos)/dt/1e6:6.2f} M chunks/s  {mb:.0f} MB to host -> {mb/1e3/dt:.1f} GB/s  verts {nv} inds {ni}")
