import os, time, torch
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    try: print(p, open(p).read().strip())
    except Exception as e: print(p, "n/a")
print("mem", os.popen("free -g | head -2").read())
a = torch.randn(4096, 4096); b = torch.randn(4096, 4096)
for nt in (8, 16, 32, 64, 128):
    torch.set_num_threads(nt)
    a @ b
    t = time.time()
    for _ in range(3): a @ b
    dt = (time.time() - t) / 3
    print(nt, "threads: %.3f s, %.1f GFLOP/s" % (dt, 2 * 4096**3 / dt / 1e9), flush=True)
w = torch.randn(32000, 4096); x = torch.randn(1, 4096)
for nt in (8, 32, 128):
    torch.set_num_threads(nt)
    x @ w.t()
    t = time.time()
    for _ in range(5): x @ w.t()
    dt = (time.time() - t) / 5
    print(nt, "gemv: %.4f s %.1f GB/s" % (dt, w.numel() * 4 / dt / 1e9), flush=True)
