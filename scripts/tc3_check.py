"""Development check of the value-net kernel (leaf_mlp_tc3.cuh): outputs vs the fp32 oracle net for several wave sizes, then time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rebel_b200 as rb
from oracle.oracle import Oracle, game_dims
from rebel_b200.models import flatten_state_dict, make_selfplay_net

D, F = int(os.environ.get("D", 1)), int(os.environ.get("F", 6))
A, H, Q = game_dims(D, F)
port = Oracle("port")
w = flatten_state_dict(make_selfplay_net(D, F, seed=0).state_dict())
rng = np.random.RandomState(0)
for mode_name in ("NET_TC_F16X2", "NET_TC_F16"):
    for K in (1, 2, 3, 149, 1300):
        b = rng.rand(K, 2, H); b /= b.sum(-1, keepdims=True)
        S = rb.WaveSolver(D, F, K, net_mode=getattr(rb, mode_name))
        S.set_weights(w)
        S.begin(np.full(K, -1, np.int32), np.zeros(K, np.int32), b)
        S.run(2)
        q, o, sc = S.leaf_io()
        want = port.net2_forward(w, Q, 256, H, q)
        err = np.abs(o - want)
        print(mode_name, "K", K, "rows", len(q), "max err", float(err.max()), "mean err", float(err.mean()), "rel rms", float(np.sqrt((err ** 2).mean() / (want ** 2).mean())), flush=True)
        S.close()
K = 8192
b = rng.rand(K, 2, H); b /= b.sum(-1, keepdims=True)
for mode_name in ("NET_TC_F16X2", "NET_TC_F16"):
    S = rb.WaveSolver(D, F, K, net_mode=getattr(rb, mode_name))
    S.set_weights(w)
    S.begin(np.full(K, -1, np.int32), np.zeros(K, np.int32), b)
    S.set_profiling(8)
    for _ in range(3):
        S.reset(); S.run(256); S.sync()
    tot, net = S.last_run_ms()
    print(mode_name, "K=8192 root wave: ms per iteration", tot / 256, "value net us per launch", net / 256 * 1e3, "rows", S.leaf_rows, flush=True)
    S.close()
