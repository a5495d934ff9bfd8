import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
import underwaterworld_b200 as uw
b = uw.ChunkBuilder(uw.Perlin(0))
st = torch.cuda.current_stream(); b.set_stream(st.cuda_stream)
allpos = uw.region.box_region((-16, 16), (-16, 16), (-4, 4))
for n in (1, 148, 592, 1024, 2048, 4096, 8192):
    pos = allpos[:n] if n != 2048 else uw.region.config_positions("spawn")
    d_pos = torch.from_numpy(np.ascontiguousarray(pos)).cuda()
    for i in range(5): b.build_device(d_pos.data_ptr(), n)
    b.sync()
    # single-step events
    ts = []
    for i in range(30):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st); b.build_device(d_pos.data_ptr(), n); e1.record(st); b.sync(); ts.append(e0.elapsed_time(e1))
    # 20 steps between one event pair
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for i in range(20): b.build_device(d_pos.data_ptr(), n)
    e1.record(st); b.sync()
    print(f"n={n:5d}  single-step median {1e3*np.median(ts):7.2f} us   back-to-back x20 avg {1e3*e0.elapsed_time(e1)/20:7.2f} us")
