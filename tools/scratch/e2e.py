import sys, os, time, ctypes as C, numpy as np, torch
sys.path.insert(0, os.getcwd())
import underwaterworld_b200 as uw
from underwaterworld_b200 import _ffi
lib = uw.load_library()
pos = uw.region.config_positions("spawn")
b = uw.ChunkBuilder(uw.Perlin(0))
ctx = b._ctx
view = _ffi.UwBatchView(); h = C.c_void_p()
def step():
    st = lib.uw_build(ctx, pos.ctypes.data, len(pos), C.byref(h))
    assert st == 0, lib.uw_last_error(ctx)
    lib.uw_batch_view_get(h, C.byref(view)); nv, ni = view.n_verts, view.n_inds
    # touch the data like a consumer would (checksum of first/last bytes)
    lib.uw_batch_free(h)
    return nv, ni
for i in range(10): step()
ts = []
for i in range(100):
    t0 = time.perf_counter(); nv, ni = step(); ts.append(time.perf_counter() - t0)
print(f"zero_copy={os.environ.get('UW_ZERO_COPY')}  e2e median {1e6*np.median(ts):.1f} us  p10 {1e6*np.quantile(ts,0.1):.1f}  min {1e6*min(ts):.1f}  nv={nv} ni={ni}")
ref = uw.ChunkBuilder(uw.Perlin(0), ordered=True).build(pos)
got = b.build(pos)
ok = all(np.array_equal(got.chunk(i).inds, ref.chunk(i).inds) and np.array_equal(got.chunk(i).verts.view(np.uint8), ref.chunk(i).verts.view(np.uint8)) for i in range(len(pos)))
print("per-chunk identical to ordered device path:", ok)
