"""Why is bench.py's device-resident loop (reset + run) slower than its e2e loop (begin + run)?  (development aid)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rebel_b200 as rb
from rebel_b200.models import make_selfplay_net, flatten_state_dict
D, F, K, iters = 1, 6, 8192, 1024
H = 6
w = flatten_state_dict(make_selfplay_net(D, F, seed=0).state_dict())
b = np.random.RandomState(1).rand(K, 2, H); b /= b.sum(-1, keepdims=True)
lb = np.full(K, -1, np.int32); pl = np.zeros(K, np.int32)
act = np.random.RandomState(0).randint(0, iters + 1, size=K).astype(np.int32)
S = rb.WaveSolver(D, F, K, num_iters=iters, net_mode=rb.NET_TC_F16X2)
S.set_weights(w)
stream = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
S.begin(lb, pl, b, act)
def timed(name, fn, n=3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:34s} {e0.elapsed_time(e1) / n:8.1f} ms/step (wall {1e3 * (time.time() - t0) / n:8.1f})", flush=True)
def reset_run(use_stream=True, flush_it=True, sync=True):
    if flush_it: flush.fill_(1)
    S.reset(stream if use_stream else None); S.run(iters, stream if use_stream else None)
    if sync: S.sync() if not use_stream else torch.cuda.synchronize()
def begin_run():
    flush.fill_(1); S.begin(lb, pl, b, act); S.run(iters, stream); torch.cuda.synchronize()
def begin_noact_run():
    flush.fill_(1); S.begin(lb, pl, b, None); S.run(iters, stream); torch.cuda.synchronize()
for rep in range(2):
    timed("reset+run (torch stream)", reset_run)
    timed("begin+run", begin_run)
    timed("reset+run no flush", lambda: reset_run(flush_it=False))
    timed("reset+run own stream", lambda: reset_run(use_stream=False))
    timed("begin(no act)+run", begin_noact_run)
    timed("reset+run after no-act begin", reset_run)
    S.begin(lb, pl, b, act)
