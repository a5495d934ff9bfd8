import torch, time, numpy as np
for mb in (0.064, 2.6, 9.6, 12.3):
    n = int(mb * 1e6)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    for _ in range(5): h.copy_(d, non_blocking=True); torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        torch.cuda.synchronize(); t0 = time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(f"D2H {mb:6.3f} MB: {1e6*np.median(ts):7.1f} us  -> {n/np.median(ts)/1e9:5.1f} GB/s")
# 4 copies of a quarter each, back to back
n = int(12.3e6 / 4)
d = torch.empty(4 * n, dtype=torch.uint8, device="cuda"); h = torch.empty(4 * n, dtype=torch.uint8, pin_memory=True)
ts = []
for _ in range(20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(4): h[k*n:(k+1)*n].copy_(d[k*n:(k+1)*n], non_blocking=True)
    torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print(f"4 x quarter copies: {1e6*np.median(ts):7.1f} us")
