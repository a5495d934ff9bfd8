"""Fused-path vertex tiling: a build with a tiny vertex-list tile (-DUW_VLIST_CAP=256: every surface chunk takes the
multi-tile path) must produce byte-identical per-chunk buffers.  python tools/scratch/cap_check.py small_cap_lib.so"""
import os, subprocess, sys, hashlib
WORKER = r'''
import sys, os, hashlib, numpy as np
sys.path.insert(0, os.getcwd())
import underwaterworld_b200 as uw
pos = uw.region.config_positions("spawn")
h = hashlib.sha256()
for kw in (dict(), dict(ordered=True), dict(tris=True)):
    with uw.ChunkBuilder(uw.Perlin(0), **kw) as b:
        batch = b.build(pos)
        for i in range(len(pos)):
            m = batch.chunk(i)
            h.update(m.verts.tobytes()); h.update(m.inds.tobytes())
        print(kw, batch.n_verts, batch.n_inds, max(len(batch.chunk(i).verts) for i in range(len(pos))))
print("digest", h.hexdigest())
'''
outs = []
for lib in (None, sys.argv[1]):
    env = dict(os.environ)
    if lib: env["UWCUDA_LIB"] = os.path.abspath(lib)
    r = subprocess.run([sys.executable, "-c", WORKER], env=env, capture_output=True, text=True)
    print(lib or "default", r.stdout.strip() or r.stderr[-400:])
    outs.append(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else None)
print("IDENTICAL" if outs[0] and outs[0] == outs[1] else "DIFFERENT")
