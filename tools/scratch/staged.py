import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
import underwaterworld_b200 as uw
from underwaterworld_b200 import region
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, pos in (("2048", region.config_positions("spawn")), ("32768", region.box_region((-32, 32), (-32, 32), (-4, 4)))):
    d_pos = torch.from_numpy(pos).cuda()
    stream = torch.cuda.Stream()
    with uw.ChunkBuilder(uw.Perlin(0), staged=True) as b:
        b.set_stream(stream.cuda_stream); b.set_profiling(True)
        acc = {}
        reps = 10
        for i in range(reps + 2):
            flush.fill_(i & 0xFF); torch.cuda.synchronize()
            b.build_device(d_pos.data_ptr(), len(pos)); b.sync()
            if i >= 2:
                t = b.stage_times()
                for k in ("noise_ms", "classify_ms", "scan_ms", "emit_ms", "total_ms"):
                    acc[k] = acc.get(k, 0) + t[k] / reps
        dv = b.device_view()
        nv, ni = int(dv.n_verts), int(dv.n_inds)
    n = len(pos); L3 = 2197
    hb = None
    alg = n * 4 * L3 + n * 44 + 24 * nv + 2 * ni
    ext = acc["classify_ms"] + acc["scan_ms"] + acc["emit_ms"]
    print(name, {k: round(v * 1e3, 1) for k, v in acc.items()}, f"nv={nv} ni={ni} ext={ext*1e3:.1f}us alg={alg/1e6:.1f}MB -> {alg/ext/1e6:.0f} GB/s ({alg/ext/1e6/6459.3:.3f} of HBM)")
