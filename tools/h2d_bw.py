import time, numpy as np, torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcqueen_b200 import api
n = 256 << 20
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
pin = torch.empty(n, dtype=torch.uint8).pin_memory()
arr = np.zeros(n, np.uint8); arr[:] = 1
pageable = torch.from_numpy(arr)
def bw(src, label):
    torch.cuda.synchronize()
    for _ in range(2): dev.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(5): dev.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
    print(f"{label}: {n/dt/1e9:.1f} GB/s")
bw(pin, "torch pinned")
bw(pageable, "pageable numpy")
api.host_register(arr)
bw(pageable, "cudaHostRegister'ed numpy")
api.host_unregister(arr)
h = torch.empty(n, dtype=torch.uint8).pin_memory()
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(5): h.copy_(dev, non_blocking=True)
torch.cuda.synchronize(); print(f"D2H pinned: {n/((time.perf_counter()-t)/5)/1e9:.1f} GB/s")
