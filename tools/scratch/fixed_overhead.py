"""Device time of the fused kernel on batches that isolate its fixed costs: events around 50 back-to-back launches."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
import underwaterworld_b200 as uw
b = uw.ChunkBuilder(uw.Perlin(0))
st = torch.cuda.current_stream(); b.set_stream(st.cuda_stream)
def run(name, pos):
    pos = np.ascontiguousarray(pos.astype(np.int32)); d = torch.from_numpy(pos).cuda()
    for i in range(5): b.build_device(d.data_ptr(), len(pos))
    b.sync(); res = []
    for g in range(6):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for i in range(50): b.build_device(d.data_ptr(), len(pos))
        e1.record(st); b.sync(); torch.cuda.synchronize(); res.append(e0.elapsed_time(e1) / 50 * 1e3)
    v = b.device_view()
    print(f"{name:34s} n={len(pos):5d}  {np.median(res):7.2f} us/launch   inds {int(v.n_inds)}")
R = uw.region.box_region
run("1 blank chunk (z=3)", R((0, 1), (0, 1), (3, 4)))
run("148 blank chunks (z=3)", R((0, 148), (0, 1), (3, 4)))
run("592 blank chunks (z=3)", R((0, 148), (0, 4), (3, 4)))
run("1184 blank chunks (z=3)", R((0, 148), (0, 8), (3, 4)))
run("2368 blank chunks (z=3)", R((0, 148), (0, 16), (3, 4)))
run("592 surface chunks (z=-1)", R((0, 148), (0, 4), (-1, 0)))
run("1184 surface chunks (z=-1)", R((0, 148), (0, 8), (-1, 0)))
run("592 solid chunks (z=-4)", R((0, 148), (0, 4), (-4, -3)))
run("config 2", uw.region.config_positions("spawn"))
