"""Max |FP32 path density - f64 oracle density| over a sample of config-2 chunks (and the guard-band count)."""
import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
import underwaterworld_b200 as uw
from oracle import Oracle
o = Oracle(12); perm = o.perm_table(0)
pos = uw.region.config_positions("spawn")[::7]
with uw.ChunkBuilder(uw.Perlin(0)) as b:
    d = b.debug_densities(pos)
ref = np.stack([o.densities(perm, tuple(int(v) for v in p)) for p in pos])
err = np.abs(d.astype(np.float64) - ref.astype(np.float64))
print(f"{len(pos)} chunks, {err.size} samples: max |err| {err.max():.3e}, mean {err.mean():.3e}, p99.9 {np.quantile(err, 0.999):.3e}")
for seed in (1, 42):
    perm = o.perm_table(seed)
    with uw.ChunkBuilder(uw.Perlin(seed)) as b:
        d = b.debug_densities(pos[:60])
    ref = np.stack([o.densities(perm, tuple(int(v) for v in p)) for p in pos[:60]])
    print(f"seed {seed}: max |err| {np.abs(d.astype(np.float64) - ref).max():.3e}")
