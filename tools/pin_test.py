import time, torch, numpy as np
torch.cuda.init(); torch.zeros(1, device="cuda")
for mb in (12, 74, 74, 300):
    n = mb * 1024 * 1024 // 8
    t = time.perf_counter(); a = torch.empty(n, dtype=torch.float64, pin_memory=True); t1 = time.perf_counter()
    print(f"pinned alloc {mb} MB: {1e3*(t1-t):.2f} ms")
    del a
src = torch.empty(74 * 1024 * 1024 // 8, dtype=torch.float64, pin_memory=True).numpy()
for rep in range(3):
    t = time.perf_counter(); dst = np.empty_like(src); dst[:] = src; t1 = time.perf_counter()
    print(f"copy 74 MB pinned -> fresh pageable: {1e3*(t1-t):.2f} ms")
    t = time.perf_counter(); dst[:] = src; t1 = time.perf_counter()
    print(f"copy 74 MB pinned -> touched pageable: {1e3*(t1-t):.2f} ms")
