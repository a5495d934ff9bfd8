import sys, os, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
import underwaterworld_b200 as uw
b = uw.ChunkBuilder(uw.Perlin(0))
st = torch.cuda.current_stream(); b.set_stream(st.cuda_stream)
pos = uw.region.config_positions("large")
print("chunks", len(pos))
d_pos = torch.from_numpy(pos).cuda()
for sub in (65536, 131072, 524288):
    def run():
        tv = ti = 0
        for s0 in range(0, len(pos), sub):
            n = min(sub, len(pos) - s0)
            b.build_device(d_pos.data_ptr() + s0 * 12, n)
            b.sync()
            v = b.device_view(); tv += v.n_verts; ti += v.n_inds
        return tv, ti
    run()
    torch.cuda.synchronize(); t0 = time.perf_counter(); tv, ti = run(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"sub-batch {sub:7d}: {dt*1e3:8.2f} ms  {len(pos)/dt/1e6:7.2f} M chunks/s  {len(pos)*1728/dt/1e9:7.1f} G voxels/s  verts {tv} inds {ti}")
print("mem GB", torch.cuda.mem_get_info())
