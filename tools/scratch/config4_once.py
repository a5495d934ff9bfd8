import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
import underwaterworld_b200 as uw
pos = uw.region.config_positions("spawn")
d_pos = torch.from_numpy(pos).cuda()
with uw.ChunkBuilder(uw.Perlin(0), internal_size=64) as b:
    for i in range(2):
        b.build_device(d_pos.data_ptr(), len(pos)); b.sync()
print("ok")
