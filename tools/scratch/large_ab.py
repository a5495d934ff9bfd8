"""A/B of library builds on config 3 (524288 chunks, one launch): python tools/scratch/large_ab.py libA.so libB.so"""
import os, subprocess, sys
WORKER = r'''
import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
import underwaterworld_b200 as uw
b = uw.ChunkBuilder(uw.Perlin(0))
st = torch.cuda.current_stream(); b.set_stream(st.cuda_stream)
pos = uw.region.config_positions("large")
d_pos = torch.from_numpy(pos).cuda()
for i in range(2): b.build_device(d_pos.data_ptr(), len(pos)); b.sync()
ts = []
for i in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st); b.build_device(d_pos.data_ptr(), len(pos)); e1.record(st)
    b.sync(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print("%.3f ms median, %.3f min" % (float(np.median(ts)), min(ts)))
'''
for rnd in range(2):
    for l in sys.argv[1:]:
        env = dict(os.environ); env["UWCUDA_LIB"] = os.path.abspath(l)
        out = subprocess.run([sys.executable, "-c", WORKER], env=env, capture_output=True, text=True)
        print(os.path.basename(l), out.stdout.strip() or out.stderr[-300:])
