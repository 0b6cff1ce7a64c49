"""Value-net kernels in the SELF-PLAY regime (non-root subgames, concentrated beliefs): per mode the output error against the
fp32 oracle net on the kernel's own query rows, and the effect on the 1024-iteration root value means relative to the fp32 SIMT net."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rebel_b200 as rb
from oracle.oracle import Oracle, game_dims
from rebel_b200.models import flatten_state_dict, make_selfplay_net

D, F = int(os.environ.get("D", 1)), int(os.environ.get("F", 6))
K = 1024
A, H, Q = game_dims(D, F)
port = Oracle("port")
w = flatten_state_dict(make_selfplay_net(D, F, seed=0).state_dict())
np.set_printoptions(linewidth=250, precision=5, suppress=True)
# states of games after a few waves of fp32 self-play
S = rb.WaveSolver(D, F, K, num_iters=256, net_mode=rb.NET_FP32)
S.set_weights(w)
S.selfplay_create(np.arange(K, dtype=np.uint32) * 1000000 + 11)
for _ in range(5):
    S.selfplay_wave()
S.selfplay_wave(start_next=False)
S.sync()
lb, pl, bel = S.selfplay_state()
S.close()
print("last_bid histogram", np.bincount(lb + 1, minlength=A + 1))
mus = {}
for name in ("NET_FP32", "NET_TC_F16", "NET_TC_F16X2"):
    S = rb.WaveSolver(D, F, K, num_iters=1024, net_mode=getattr(rb, name))
    S.set_weights(w)
    S.begin(lb, pl, bel)
    S.run(9)
    q, o, sc = S.leaf_io()
    want = port.net2_forward(w, Q, 256, H, q)
    e = (o - want).astype(np.float64)
    rel = np.sqrt((e ** 2).mean() / (want.astype(np.float64) ** 2).mean())
    row_err = np.sqrt((e ** 2).mean(1)); row_mag = np.sqrt((want.astype(np.float64) ** 2).mean(1))
    worst = np.argsort(-row_err / (row_mag + 1e-12))[:5]
    print(f"{name}: rows {len(q)} rel rms {rel:.3e} max abs err {np.abs(e).max():.3e} (|out| rms {row_mag.mean():.3e}); bias per hand {e.mean(0)}")
    print(f"   rows with rel err > 1e-2: {(row_err > 1e-2 * row_mag).sum()}, > 1e-1: {(row_err > 1e-1 * row_mag).sum()}")
    for r in worst[:3]:
        print("   worst row", r, "rel", row_err[r] / row_mag[r], "q", q[r], "got", o[r], "want", want[r])
    S.run(1024 - 9)
    mus[name] = S.fetch(("root_means",))["root_means"].copy()
    S.close()
for name in ("NET_TC_F16", "NET_TC_F16X2"):
    d = np.abs(mus[name] - mus["NET_FP32"])
    per = d.reshape(K, -1).mean(1)
    print(f"{name}: mean |dmu| vs fp32 net {d.mean():.3e}, max {d.max():.3e}; per-subgame mean: median {np.median(per):.3e} p90 {np.percentile(per, 90):.3e} p99 {np.percentile(per, 99):.3e}")
    bad = np.argsort(-per)[:5]
    print("   worst subgames: last_bid", lb[bad], "player", pl[bad], "dmu", per[bad])
    for lbv in np.unique(lb):
        m = lb == lbv
        print(f"   last_bid {lbv:3d}: n {m.sum():4d} mean |dmu| {per[m].mean():.3e}")
