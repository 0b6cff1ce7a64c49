"""Kernel-level timing probe (development aid): value-net and CFR kernel time per launch for several configurations."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rebel_b200 as rb
from rebel_b200.models import make_selfplay_net, flatten_state_dict

def main():
    for (D, F, K) in [(1, 6, 8192), (1, 4, 4096), (2, 3, 8192)]:
        A = 1 + 2 * D * F; H = F ** D
        w = flatten_state_dict(make_selfplay_net(D, F, seed=0).state_dict())
        rng = np.random.RandomState(1)
        b = rng.rand(K, 2, H); b /= b.sum(-1, keepdims=True)
        for roots in ("root", "mixed"):
            lb = np.full(K, -1, np.int32) if roots == "root" else rng.randint(-1, A - 1, size=K).astype(np.int32)
            pl = np.zeros(K, np.int32) if roots == "root" else rng.randint(0, 2, size=K).astype(np.int32)
            for state in (rb.STATE_F64, rb.STATE_F32):
                for net in (rb.NET_TC_F16, rb.NET_ZERO):
                    S = rb.WaveSolver(D, F, K, net_mode=net, state_dtype=state)
                    if net: S.set_weights(w)
                    S.begin(lb, pl, b)
                    S.run(4); S.sync()
                    S.set_profiling(True)
                    S.run(64); S.sync()
                    tot, nt = S.last_run_ms()
                    print(f"{D}x{F}f K={K} {roots:5s} state={'f64' if state == 0 else 'f32'} net={'tc' if net else 'zero'}: 64 iters {tot:7.2f} ms; "
                          f"net {nt / 64 * 1e3:6.1f} us/launch, cfr {(tot - nt) / 65 * 1e3:6.1f} us/launch, rows {S.leaf_rows}; "
                          f"{K * 64 / (tot * 1e-3) / 1e6:6.2f} M subgame-iters/s", flush=True)
                    S.close()

if __name__ == "__main__":
    main()
