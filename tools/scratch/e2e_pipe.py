import sys, os, time, ctypes as C, numpy as np, torch
sys.path.insert(0, os.getcwd())
import underwaterworld_b200 as uw
from underwaterworld_b200 import _ffi
lib = uw.load_library()
pos = uw.region.config_positions("spawn")
b = uw.ChunkBuilder(uw.Perlin(0))
ctx = b._ctx
view = _ffi.UwBatchView()
def single():
    h = C.c_void_p()
    assert lib.uw_build(ctx, pos.ctypes.data, len(pos), C.byref(h)) == 0
    lib.uw_batch_view_get(h, C.byref(view)); lib.uw_batch_free(h)
for i in range(10): single()
ts = []
for i in range(100):
    t0 = time.perf_counter(); single(); ts.append(time.perf_counter() - t0)
print(f"single-call e2e median {1e6*np.median(ts):.1f} us  min {1e6*min(ts):.1f}")
def pipelined(K):
    prev = C.c_void_p()
    assert lib.uw_build_async(ctx, pos.ctypes.data, len(pos), C.byref(prev)) == 0
    for k in range(1, K):
        nxt = C.c_void_p()
        assert lib.uw_build_async(ctx, pos.ctypes.data, len(pos), C.byref(nxt)) == 0, lib.uw_last_error(ctx)
        assert lib.uw_batch_wait(prev) == 0
        lib.uw_batch_view_get(prev, C.byref(view)); lib.uw_batch_free(prev)
        prev = nxt
    assert lib.uw_batch_wait(prev) == 0
    lib.uw_batch_view_get(prev, C.byref(view)); lib.uw_batch_free(prev)
pipelined(20)
for K in (50, 200):
    t0 = time.perf_counter(); pipelined(K); dt = time.perf_counter() - t0
    print(f"pipelined K={K}: {1e6*dt/K:.1f} us/step  {len(pos)*K/dt/1e6:.2f} M chunks/s  nv={view.n_verts}")
