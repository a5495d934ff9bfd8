import sys, os, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
import underwaterworld_b200 as uw
pos = uw.region.config_positions("spawn")
b = uw.ChunkBuilder(uw.Perlin(0), internal_size=64)
st = torch.cuda.current_stream(); b.set_stream(st.cuda_stream)
d_pos = torch.from_numpy(pos).cuda()
b.set_profiling(True)
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    b.build_device(d_pos.data_ptr(), len(pos)); b.sync()
    dt = time.perf_counter() - t0
    v = b.device_view(); t = b.stage_times()
    print(f"config 4 (2048 x 64^3): {dt*1e3:.2f} ms  {len(pos)/dt:.0f} chunks/s  {len(pos)*64**3/dt/1e9:.2f} G voxels/s  verts {v.n_verts} inds {v.n_inds}  stages {t}")
print(torch.cuda.mem_get_info())
