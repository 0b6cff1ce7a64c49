"""Two-tiles-in-flight value-net kernel vs the one-tile kernel: bit-identical outputs over wave sizes that give every CTA an
odd / even / single number of tiles, then time per launch (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rebel_b200 as rb
from rebel_b200.models import make_selfplay_net, flatten_state_dict


def outputs(D, F, K, mode, no_tc2, iters=3, timing=False):
    os.environ["CFRB_TC2"] = "0" if no_tc2 else "1"
    H = F ** D
    w = flatten_state_dict(make_selfplay_net(D, F, seed=0).state_dict())
    rng = np.random.RandomState(K)
    b = rng.rand(K, 2, H); b /= b.sum(-1, keepdims=True)
    S = rb.WaveSolver(D, F, K, net_mode=mode)
    S.set_weights(w)
    S.begin(np.full(K, -1, np.int32), np.zeros(K, np.int32), b)
    S.run(iters)
    q, o, sc = S.leaf_io()
    mu = S.fetch(("root_means",))["root_means"]
    t = None
    if timing:
        S.set_profiling(True); S.run(64); S.sync()
        tot, nt = S.last_run_ms(); t = (nt / 64 * 1e3, (tot - nt) / 65 * 1e3)
    S.close()
    return o, mu, t


def main():
    for (D, F) in [(1, 6), (1, 4)]:
        for K in (1, 2, 3, 5, 148, 149, 300, 1000, 4097, 8192):
            for mode in (rb.NET_TC_F16X2, rb.NET_TC_F16):
                a = outputs(D, F, K, mode, True)
                b = outputs(D, F, K, mode, False)
                ok = np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
                print(f"{D}x{F}f K={K:5d} mode={mode}: rows={a[0].shape[0]:7d} identical={ok} max|d|={np.abs(a[0] - b[0]).max():.3e}", flush=True)
                assert ok
    for (D, F, K) in [(1, 6, 8192), (1, 4, 4096)]:
        for no in (True, False):
            _, _, t = outputs(D, F, K, rb.NET_TC_F16X2, no, timing=True)
            print(f"{D}x{F}f K={K} {'one-tile' if no else 'two-tile'} kernel: net {t[0]:.1f} us/launch, cfr {t[1]:.1f} us/launch", flush=True)


if __name__ == "__main__":
    main()
