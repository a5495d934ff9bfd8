#!/usr/bin/env python3
"""Runs one configuration a few times and exits -- the command ncu captures are taken from.

    python tools/once.py fused 8          config 2 shape, 16x16x8 = 2048 chunks (n_side 8), default fused kernel
    python tools/once.py fused 32         32 768 chunks
    python tools/once.py config3          524 288 chunks, one launch
    python tools/once.py staged 32        UW_FLAG_STAGED pipeline (k_noise_spec, k_classify_spec, k_scan_chunks, k_emit_small)
    python tools/once.py config4          2048 chunks of 64^3
"""
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
import underwaterworld_b200 as uw  # noqa: E402
from underwaterworld_b200 import region  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "fused"
n_side = int(sys.argv[2]) if len(sys.argv) > 2 else 8
kw, S = {}, 12
if mode == "config3":
    pos = region.config_positions("large")
elif mode == "config4":
    pos, S = region.config_positions("spawn"), 64
else:
    pos = region.box_region((-n_side, n_side), (-n_side, n_side), (-4, 4))
    kw = {"staged": True} if mode == "staged" else {}
d_pos = torch.from_numpy(pos).cuda()
with uw.ChunkBuilder(uw.Perlin(0), internal_size=S, **kw) as b:
    for i in range(4):
        b.build_device(d_pos.data_ptr(), len(pos))
        b.sync()
print("ok", mode, len(pos))
