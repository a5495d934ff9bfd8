import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
import underwaterworld_b200 as uw
from underwaterworld_b200 import region
n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 8
pos = region.box_region((-n_side, n_side), (-n_side, n_side), (-4, 4))
d_pos = torch.from_numpy(pos).cuda()
with uw.ChunkBuilder(uw.Perlin(0)) as b:
    for i in range(4):
        b.build_device(d_pos.data_ptr(), len(pos)); b.sync()
print("ok", len(pos))
