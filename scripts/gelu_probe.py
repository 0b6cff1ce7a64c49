"""A/B of the two tcgen05 epilogues (CFRB_NET_TC_F16 fp32 GELU vs CFRB_NET_TC_F16X2 packed-half GELU): accuracy of the raw net
outputs against torch fp32 on the same query rows, and kernel time per launch on the bench workload (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rebel_b200 as rb
from rebel_b200.models import make_selfplay_net, flatten_state_dict


def main():
    for (D, F, K) in [(1, 6, 8192), (2, 3, 8192), (1, 4, 4096)]:
        A = 1 + 2 * D * F; H = F ** D
        net = make_selfplay_net(D, F, seed=0)
        w = flatten_state_dict(net.state_dict())
        rng = np.random.RandomState(1)
        b = rng.rand(K, 2, H); b /= b.sum(-1, keepdims=True)
        lb = np.full(K, -1, np.int32); pl = np.zeros(K, np.int32)
        for mode, name in ((rb.NET_TC_F16, "tc f32-gelu"), (rb.NET_TC_F16X2, "tc f16x2-gelu")):
            S = rb.WaveSolver(D, F, K, net_mode=mode)
            S.set_weights(w)
            S.begin(lb, pl, b)
            S.run(5)
            q, o, sc = S.leaf_io()
            with torch.no_grad():
                ref = net(torch.from_numpy(q.astype(np.float32))).numpy()
            n = min(len(ref), 20000)
            err = np.abs(o[:n] - ref[:n])
            rel = np.sqrt((err ** 2).mean()) / np.sqrt((ref[:n] ** 2).mean())
            S.set_profiling(True)
            S.run(64); S.sync()
            tot, nt = S.last_run_ms()
            print(f"{D}x{F}f K={K} {name:14s}: net out vs torch fp32: max abs {err.max():.3e} rms-rel {rel:.3e} (|ref| rms {np.sqrt((ref[:n]**2).mean()):.3e}); "
                  f"net {nt / 64 * 1e3:6.1f} us/launch, cfr {(tot - nt) / 65 * 1e3:6.1f} us/launch; {K * 64 / (tot * 1e-3) / 1e6:6.2f} M subgame-iters/s", flush=True)
            S.close()


if __name__ == "__main__":
    main()
