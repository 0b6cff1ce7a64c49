"""Host CPU diagnostics for the CPU-baseline leg (development aid)."""
import os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    if os.path.exists(f): print(f, open(f).read().strip())
print(subprocess.run("lscpu | grep -E 'Model name|Socket|Thread|Core|MHz' | head -8; cat /proc/loadavg", shell=True, capture_output=True, text=True).stdout)
import numpy as np
import bench
from oracle.oracle import Oracle, available
if available("ref_fast"):
    path = bench.script_net_file(1, 6)
    R = Oracle("ref_fast")
    for T, n in [(1, 2), (8, 16), (32, 64), (64, 128), (128, 256)]:
        b = bench.workload_beliefs(n, 6, 0)
        t = R.bench_solve(1, 6, n, script_path=path, threads=T, num_iters=256, beliefs=b)
        print(f"threads {T:4d}: {n} subgames x 256 iters in {t:.2f} s -> {n*256/t:.0f} subgame-iters/s ({n*256/t/T:.0f} per thread)", flush=True)
