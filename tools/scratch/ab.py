"""A/B timing of kernel variants in ONE gpurun call: python tools/scratch/ab.py libA.so libB.so ...
Each variant runs in its own subprocess (UWCUDA_LIB), interleaved over several rounds."""
import os, subprocess, sys, json
import numpy as np
WORKER = r'''
import sys, os, json, numpy as np, torch
sys.path.insert(0, os.getcwd())
import underwaterworld_b200 as uw
kw = json.loads(os.environ.get("UW_KW", "{}"))
pos = uw.region.config_positions("spawn")
d_pos = torch.from_numpy(pos).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
b = uw.ChunkBuilder(uw.Perlin(0), **kw)
st = torch.cuda.current_stream(); b.set_stream(st.cuda_stream)
for i in range(5): b.build_device(d_pos.data_ptr(), len(pos))
b.sync()
ts = []
for i in range(40):
    flush.fill_(i)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st); b.build_device(d_pos.data_ptr(), len(pos)); e1.record(st)
    b.sync(); ts.append(e0.elapsed_time(e1))
v = b.device_view()
print(json.dumps({"median_us": 1e3 * float(np.median(ts)), "min_us": 1e3 * min(ts), "nv": int(v.n_verts), "ni": int(v.n_inds)}))
'''
libs = sys.argv[1:]
res = {l: [] for l in libs}
for rnd in range(3):
    for l in libs:
        env = dict(os.environ)
        name, _, kw = l.partition("@")
        env["UWCUDA_LIB"] = os.path.abspath(name)
        env["UW_KW"] = kw or "{}"
        out = subprocess.run([sys.executable, "-c", WORKER], env=env, capture_output=True, text=True)
        try:
            res[l].append(json.loads(out.stdout.strip().splitlines()[-1]))
        except Exception:
            print(l, "FAILED", out.stderr[-500:])
for l in libs:
    if res[l]:
        print(f"{l:60s} median {np.median([r['median_us'] for r in res[l]]):7.2f} us  min {min(r['min_us'] for r in res[l]):7.2f} us  nv={res[l][0]['nv']} ni={res[l][0]['ni']}")
