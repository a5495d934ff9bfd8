"""Fine-grained A/B of the device-resident config-2 step: 50 launches between one event pair (L2 flushed before the
group), several groups, per-library subprocesses interleaved.  python tools/scratch/ab_fine.py libA.so libB.so"""
import os, subprocess, sys
WORKER = r'''
import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
import underwaterworld_b200 as uw
pos = uw.region.config_positions("spawn")
d_pos = torch.from_numpy(pos).cuda()
b = uw.ChunkBuilder(uw.Perlin(0))
st = torch.cuda.current_stream(); b.set_stream(st.cuda_stream)
for i in range(10): b.build_device(d_pos.data_ptr(), len(pos))
b.sync()
res = []
for g in range(8):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for i in range(50): b.build_device(d_pos.data_ptr(), len(pos))
    e1.record(st); b.sync(); torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / 50 * 1e3)
print("%.2f us/launch back-to-back (median of 8 groups), min %.2f" % (float(np.median(res)), min(res)))
'''
for rnd in range(3):
    for l in sys.argv[1:]:
        env = dict(os.environ); env["UWCUDA_LIB"] = os.path.abspath(l)
        out = subprocess.run([sys.executable, "-c", WORKER], env=env, capture_output=True, text=True)
        print(os.path.basename(l), out.stdout.strip() or out.stderr[-300:])
