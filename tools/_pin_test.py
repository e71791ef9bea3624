import time, torch, numpy as np
torch.cuda.init()
x = torch.empty(int(112e6 // 8), dtype=torch.float64, device="cuda").normal_()
torch.cuda.synchronize()
for mb in (28, 112):
    n = int(mb * 1e6 // 8)
    for rep in range(3):
        t0 = time.perf_counter(); h = torch.empty(n, dtype=torch.float64, pin_memory=True); t1 = time.perf_counter()
        h.copy_(x[:n], non_blocking=True); torch.cuda.synchronize(); t2 = time.perf_counter()
        a = np.array(h.numpy()); t3 = time.perf_counter()
        p = x[:n].cpu(); t4 = time.perf_counter()
        print(f"{mb} MB rep {rep}: pinned alloc {1e3*(t1-t0):.2f} ms, D2H pinned {1e3*(t2-t1):.2f} ms, host copy {1e3*(t3-t2):.2f} ms, pageable .cpu() {1e3*(t4-t3):.2f} ms")
        del h
