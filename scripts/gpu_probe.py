"""Verbose GPU-vs-oracle probe (development aid; the real checks live in tests/)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle.oracle import Oracle, game_dims
import rebel_b200 as rb
from rebel_b200.models import make_selfplay_net, flatten_state_dict

P = Oracle("port")


def dense_slice(x, N):
    return x[:, :N]


def compare(D, F, roots, iters_list, net_mode, max_depth=2, w=None):
    A, H, Q = game_dims(D, F)
    n = len(roots)
    beliefs = np.stack([P.synthetic_beliefs(H, 100 + i) for i in range(n)])
    lb = np.array([r[0] for r in roots], np.int32)
    pl = np.array([r[1] for r in roots], np.int32)
    S = rb.WaveSolver(D, F, n, max_depth=max_depth, net_mode=net_mode)
    if net_mode != rb.NET_ZERO:
        S.set_weights(w)
    S.begin(lb, pl, beliefs)
    done = 0
    ora = [P.cfr_solve(D, F, beliefs[i], iters_list, lb[i], pl[i], num_iters=max(iters_list), max_depth=max_depth,
                       net_w=(w if net_mode != rb.NET_ZERO else None)) for i in range(n)]
    for ci, it in enumerate(iters_list):
        S.run(it - done); done = it
        g = S.fetch(("root_means", "last", "avg", "sum", "regrets"))
        worst = {}
        for i in range(n):
            N = ora[i]["tree"].shape[0]
            for k in ("regrets", "last", "sum", "avg"):
                d = np.abs(g[k][i, :N] - ora[i][k][ci]).max()
                worst[k] = max(worst.get(k, 0), d)
            d = np.abs(g["root_means"][i] - ora[i]["root_means"][ci]).max()
            worst["mu"] = max(worst.get("mu", 0), d)
        print(f"  {D}x{F}f depth={max_depth} net={net_mode} iters={it:5d} maxabs:", {k: f"{v:.2e}" for k, v in worst.items()}, flush=True)
    if net_mode != rb.NET_ZERO:
        q, o, s = S.leaf_io()
        # queries/net of the LAST forward == iteration max-1
        net = torch_net(D, F)
        with torch.no_grad():
            ref = net(torch.from_numpy(q)).numpy()
        print("    leaf rows", q.shape, "net_out vs torch fp32 maxabs", np.abs(ref - o).max(), "out scale", np.abs(ref).max())
        row = 0
        for i in range(n):
            L = ora[i]["queries"].shape[1]
            dq = np.abs(q[row:row + L] - ora[i]["queries"][-1]).max() if L else 0
            print(f"    subgame {i}: query maxabs vs oracle {dq:.2e}")
            row += L
    S.close()

_nets = {}
def torch_net(D, F):
    if (D, F) not in _nets:
        _nets[(D, F)] = make_selfplay_net(D, F, seed=0)
    return _nets[(D, F)]


def main():
    print("devices:", rb.capi.lib().cfrb_device_count(), torch.cuda.get_device_name(0))
    for (D, F) in [(1, 4), (1, 6), (2, 3)]:
        A = 1 + 2 * D * F
        roots = [(-1, 0), (-1, 1), (0, 1), (3, 0), (A - 3, 1), (A - 2, 0)]
        print(f"== zero net {D}x{F}f")
        compare(D, F, roots, [1, 2, 3, 16, 128], rb.NET_ZERO)
        print(f"== fp32 net {D}x{F}f")
        w = flatten_state_dict(torch_net(D, F).state_dict())
        compare(D, F, roots, [1, 2, 3, 16, 128], rb.NET_FP32, w=w)
    print("== full tree, no net, 1x4f (CTA groups)")
    compare(1, 4, [(-1, 0)], [1, 2, 16, 256], rb.NET_ZERO, max_depth=100)
    print("== depth 3 zero net 1x4f")
    compare(1, 4, [(-1, 0), (2, 1)], [1, 2, 16], rb.NET_ZERO, max_depth=3)
    # timing
    for (D, F, K, iters) in [(1, 6, 1024, 64), (1, 6, 8192, 32)]:
        A, H, Q = game_dims(D, F)
        w = flatten_state_dict(torch_net(D, F).state_dict())
        for mode in (rb.NET_ZERO, rb.NET_FP32):
            S = rb.WaveSolver(D, F, K, net_mode=mode)
            if mode: S.set_weights(w)
            b = np.stack([P.synthetic_beliefs(H, i) for i in range(K)])
            S.begin(np.full(K, -1, np.int32), np.zeros(K, np.int32), b)
            S.run(4); S.sync()
            t = time.time(); S.run(iters); S.sync(); dt = time.time() - t
            ms, _ = S.last_run_ms()
            print(f"timing {D}x{F}f K={K} mode={mode}: {iters} iters wall {dt*1e3:.1f} ms, device {ms:.1f} ms -> {K*iters/(ms*1e-3)/1e6:.2f} M subgame-iters/s, rows={S.leaf_rows}")
            S.close()

if __name__ == "__main__":
    main()
