import sys, os, ctypes as C, numpy as np, torch
sys.path.insert(0, os.getcwd())
os.environ["UWCUDA_LIB"] = os.path.abspath(sys.argv[1])
import underwaterworld_b200 as uw
lib = uw.load_library()
pos = uw.region.config_positions("spawn")
N = int(os.environ.get("N", "2048"))
ZR = int(os.environ.get("ZR", "4"))
if N != 2048: pos = np.ascontiguousarray(uw.region.box_region((-16, 16), (-16, 16), (-ZR, ZR))[:N])
d_pos = torch.from_numpy(pos).cuda()
b = uw.ChunkBuilder(uw.Perlin(0))
for i in range(3): b.build_device(d_pos.data_ptr(), len(pos))
b.sync()
out = (C.c_ulonglong * 16)()
lib.uw_debug_phase_cycles(out, 1)
R = 10
for i in range(R): b.build_device(d_pos.data_ptr(), len(pos))
b.sync()
lib.uw_debug_phase_cycles(out, 0)
v = np.array(list(out), dtype=np.float64)
n, nact, nany = v[8], v[9], v[10]
names = ["ticket+pos", "K1 noise", "K2 prepare", "K3 offsets", "K4 D1 fill", "K4 D2+E", "tail/barrier"]
tot = v[:7].sum() + v[11] + v[12]
print(f"chunks {n/R:.0f} per build, active {nact/R:.0f}, any_lt {nany/R:.0f}; total cycles/chunk {tot/n:.0f}")
print(f"  K1 split: H {v[11]/n:.0f}  X {v[12]/n:.0f}  YZ {v[1]/n:.0f} cycles/chunk")
print(f"  end of stage YZ: spare warp last in {100*v[15]/n:.1f}% of the chunks, by {v[13]/max(v[15],1):.0f} cycles on average; otherwise the columns are last by {v[14]/max(n-v[15],1):.0f}")
for i, nm in enumerate(names):
    denom = nact if i in (4, 5) else (nany if i == 2 else n)
    print(f"  {nm:14s} {100*v[i]/tot:5.1f}%   avg cycles per (relevant) chunk {v[i]/max(denom,1):8.0f}")
cta = (C.c_ulonglong * 4096)()
b.build_device(d_pos.data_ptr(), len(pos)); b.sync()
lib.uw_debug_cta_times(cta)
a = np.array(list(cta), dtype=np.float64).reshape(1024, 4)[:min(592, N)]
t0 = a[:, 0].min()
st, en = (a[:, 0] - t0) / 1e3, (a[:, 1] - t0) / 1e3
print(f"CTA start us: min {st.min():.1f} max {st.max():.1f};  end us: min {en.min():.1f} median {np.median(en):.1f} max {en.max():.1f}")
print(f"chunks per CTA: min {a[:,2].min():.0f} mean {a[:,2].mean():.2f} max {a[:,2].max():.0f}; heavy per CTA: min {a[:,3].min():.0f} mean {a[:,3].mean():.2f} max {a[:,3].max():.0f}")
print("end-time histogram (us):", np.histogram(en, bins=8)[0].tolist(), np.round(np.histogram(en, bins=8)[1], 1).tolist())
order = np.argsort(-en)[:12]
print("latest CTAs (end us, start us, chunks, mesh chunks):", [(round(float(en[i]), 1), round(float(st[i]), 1), int(a[i, 2]), int(a[i, 3])) for i in order])
one = en[a[:, 2] == 1]
print("CTAs with exactly 1 chunk:", len(one), "end us max", one.max() if len(one) else None)
for k in (1, 2, 3, 4, 5):
    sel = a[:, 2] == k
    if sel.any(): print(f"  CTAs with {k} chunks: {int(sel.sum())}, end us mean {en[sel].mean():.1f} max {en[sel].max():.1f}, mesh chunks mean {a[sel, 3].mean():.2f}")
