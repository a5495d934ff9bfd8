#!/usr/bin/env python3
"""uw_multi_build on every visible GPU over the 524 288-chunk region: the render share the library's search
(uw_share_search_next) picks request by request, with the wall time of each request.  usage: python tools/multi_share_trace.py [requests]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import underwaterworld_b200 as uw
reqs = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ndev = torch.cuda.device_count()
pos = uw.region.config_positions("large")
pin = torch.from_numpy(pos).pin_memory().numpy()
with uw.MultiBuilder(uw.Perlin(0), devices=list(range(ndev))) as mb:
    n_inds = None
    for k in range(reqs):
        share = mb._lib.uw_multi_render_share(mb._m)
        t0 = time.perf_counter()
        res = mb.build(pin, draw_to_host=True)
        dt = time.perf_counter() - t0
        assert n_inds in (None, res.n_inds)
        n_inds = res.n_inds
        print(f"request {k:2d}  {ndev} GPUs  render share {share / 1000:.3f}  {1e3 * dt:7.3f} ms  n_inds {res.n_inds}")
