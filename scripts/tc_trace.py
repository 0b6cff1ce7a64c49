"""Timeline of one CTA of the tcgen05 value-net kernel from clock64() stamps (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rebel_b200 as rb
from rebel_b200.models import make_selfplay_net, flatten_state_dict

EP = ["d1 ready", "epi1 done", "x restaged", "d2 ready", "epi2 done", "d3 ready", "out done",
      "e1:ld", "e1:pass1", "e1:bar", "e1:pass2", "e2:ld", "e2:pass1", "e2:bar", "e2:pass2"]
MM = ["x ready", "L1 issued", "a2 ready", "L2 issued", "a3 ready", "L3 issued"]


def main():
    D, F, K = 1, 6, 8192
    H = F ** D
    w = flatten_state_dict(make_selfplay_net(D, F, seed=0).state_dict())
    rng = np.random.RandomState(1)
    b = rng.rand(K, 2, H); b /= b.sum(-1, keepdims=True)
    for mode, name in ((rb.NET_TC_F16, "f32-gelu"), (rb.NET_TC_F16X2, "f16x2-gelu")):
        S = rb.WaveSolver(D, F, K, net_mode=mode)
        S.set_weights(w)
        S.begin(np.full(K, -1, np.int32), np.zeros(K, np.int32), b)
        S.run(3)
        S.net_trace()
        t = S.net_trace()
        e = t[:1024].reshape(64, 16); m = t[1024:1024 + 512].reshape(64, 8)
        n = int((e[:, 0] > 0).sum())
        t0 = e[0, 0]
        print(f"== {name}: {n} tiles on CTA 0; cycles per tile (median) {np.median(np.diff(e[:n, 0])):.0f}")
        order = [("E", 0), ("E", 7), ("E", 8), ("E", 9), ("E", 10), ("E", 1), ("M", 2), ("M", 3), ("E", 2), ("E", 3), ("E", 11), ("E", 12), ("E", 13),
                 ("E", 14), ("E", 4), ("M", 4), ("M", 5), ("E", 5), ("E", 6), ("M", 0), ("M", 1)]
        for it in (3, 4):
            base = e[it, 0]
            s = []
            for kind, i in order:
                v = (e[it, i] if kind == "E" else m[it, i]) - base
                s.append(f"{(EP[i] if kind == 'E' else 'MMA ' + MM[i])}={v}")
            print(f"  tile iter {it}: " + ", ".join(s))
        S.close()


if __name__ == "__main__":
    main()
