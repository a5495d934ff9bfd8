#!/usr/bin/env python3
"""A/B timing of kernel variants in ONE gpurun call.

    python tools/ab.py libA.so libB.so ...        (a name may carry builder kwargs and environment settings:
                                                   lib.so@{"staged":true}   lib.so@@UW_FUSED_PIPE=0,UW_X=1)

Each variant runs in its own subprocess (UWCUDA_LIB), interleaved over several rounds, on three workloads:
config 2 (2048 chunks, L2 flushed, per launch), 32 768 chunks (back to back) and config 3 (524 288 chunks, one
launch).  Variants are built with tools/build_variant.sh."""
import json
import os
import subprocess
import sys

import numpy as np

WORKER = r'''
import sys, os, json, numpy as np, torch
sys.path.insert(0, os.getcwd())
import underwaterworld_b200 as uw
kw = json.loads(os.environ.get("UW_KW", "{}"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
b = uw.ChunkBuilder(uw.Perlin(0), **kw)
st = torch.cuda.current_stream(); b.set_stream(st.cuda_stream)
out = {}
pos = uw.region.config_positions("spawn")
d_pos = torch.from_numpy(pos).cuda()
for i in range(5): b.build_device(d_pos.data_ptr(), len(pos))
b.sync()
ts = []
for i in range(40):
    flush.fill_(i)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st); b.build_device(d_pos.data_ptr(), len(pos)); e1.record(st)
    b.sync(); ts.append(e0.elapsed_time(e1))
v = b.device_view()
out["c2_median_us"] = 1e3 * float(np.median(ts)); out["c2_min_us"] = 1e3 * min(ts); out["c2_ni"] = int(v.n_inds)
for name, box, reps in (("n32768", ((-32, 32), (-32, 32), (-4, 4)), 10), ("c3", ((-64, 64), (-64, 64), (-16, 16)), 5)):
    if kw.get("staged") and name == "c3": continue
    p = uw.region.box_region(*box)
    d = torch.from_numpy(p).cuda()
    for i in range(2): b.build_device(d.data_ptr(), len(p))
    b.sync()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for i in range(reps): b.build_device(d.data_ptr(), len(p))
    e1.record(st); b.sync()
    out[name + "_us"] = 1e3 * e0.elapsed_time(e1) / reps
    out[name + "_ni"] = int(b.device_view().n_inds)
print(json.dumps(out))
'''


def main():
    libs = sys.argv[1:]
    res = {l: [] for l in libs}
    for rnd in range(3):
        for l in libs:
            env = dict(os.environ)
            name, _, rest = l.partition("@")
            kw, _, envs = rest.partition("@")
            env["UWCUDA_LIB"] = os.path.abspath(name)
            env["UW_KW"] = kw or "{}"
            for a in filter(None, envs.split(",")):
                k, _, v = a.partition("=")
                env[k] = v
            out = subprocess.run([sys.executable, "-c", WORKER], env=env, capture_output=True, text=True)
            try:
                res[l].append(json.loads(out.stdout.strip().splitlines()[-1]))
            except Exception:
                print(l, "FAILED", out.stderr[-800:])
    for l in libs:
        r = res[l]
        if not r:
            continue
        med = lambda k: float(np.median([x[k] for x in r if k in x])) if any(k in x for x in r) else float("nan")
        print(f"{os.path.basename(l.split('@')[0]) + ' ' + '@'.join(l.split('@')[1:]):44s} config2 {med('c2_median_us'):7.2f} us (min {min(x['c2_min_us'] for x in r):6.2f})   "
              f"32768 {med('n32768_us'):8.1f} us   config3 {med('c3_us'):9.1f} us   ni={r[0]['c2_ni']}/{r[0].get('n32768_ni')}/{r[0].get('c3_ni')}")


if __name__ == "__main__":
    main()
