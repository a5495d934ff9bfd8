import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo')
import underwaterworld_b200 as uw
pos = uw.region.config_positions('spawn')
d_pos = torch.from_numpy(pos).cuda()
flush = torch.empty(256<<20, dtype=torch.uint8, device='cuda')
ref=None
for name,kw in [('default(completion order)',{}),('ordered',dict(ordered=True)),('staged',dict(staged=True)),('tris',dict(tris=True)),('exportable',dict(exportable=True))]:
    b = uw.ChunkBuilder(uw.Perlin(0), **kw)
    st = torch.cuda.current_stream(); b.set_stream(st.cuda_stream)
    for flushit in (True, False):
        for i in range(5): b.build_device(d_pos.data_ptr(), len(pos))
        b.sync()
        ts=[]
        for i in range(30):
            if flushit: flush.fill_(i)
            e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
            e0.record(st); b.build_device(d_pos.data_ptr(), len(pos)); e1.record(st)
            b.sync(); ts.append(e0.elapsed_time(e1))
        print(name, 'flush' if flushit else 'noflush', 'ms/step median', np.median(ts), 'min', min(ts))
    got = b.build(pos)
    per = [(got.chunk(i).inds.tobytes(), got.chunk(i).verts.tobytes(), got.chunk(i).flags) for i in range(len(got))]
    if ref is None: ref=per
    else: print(name,'per-chunk identical to ordered:', per==ref, 'totals', got.n_verts, got.n_inds)
    b.close()
