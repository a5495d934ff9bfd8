"""Host-call latency of small batches (config 5 territory): blocking uw_build vs device-resident build+sync."""
import sys, os, time, ctypes as C, numpy as np, torch
sys.path.insert(0, os.getcwd())
import underwaterworld_b200 as uw
from underwaterworld_b200 import _ffi
lib = uw.load_library()
b = uw.ChunkBuilder(uw.Perlin(0))
ctx = b._ctx
allpos = uw.region.box_region((-8, 8), (-8, 8), (-1, 1))      # surface layers: every chunk has a mesh
view = _ffi.UwBatchView()
for n in (1, 8, 32, 128, 360):
    pos = np.ascontiguousarray(allpos[:n])
    d_pos = torch.from_numpy(pos).cuda()
    def host():
        h = C.c_void_p()
        st = lib.uw_build(ctx, pos.ctypes.data, n, C.byref(h))
        assert st == 0, (st, lib.uw_last_error(ctx))
        lib.uw_batch_view_get(h, C.byref(view)); lib.uw_batch_free(h)
    def dev():
        assert lib.uw_build_device(ctx, C.c_void_p(d_pos.data_ptr()), n) == 0
        assert lib.uw_sync(ctx) == 0
    for f in (host, dev):
        for i in range(20): f()
    th, td = [], []
    for i in range(200):
        t0 = time.perf_counter(); host(); th.append(time.perf_counter() - t0)
        t0 = time.perf_counter(); dev(); td.append(time.perf_counter() - t0)
    print(f"n={n:4d}  uw_build p50 {1e6*np.median(th):6.1f} us  p99 {1e6*np.quantile(th, 0.99):6.1f}   build_device+sync p50 {1e6*np.median(td):6.1f} us   out bytes {view.n_verts*24 + view.n_inds*2 + n*32}")
