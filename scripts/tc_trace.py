"""Timeline of one CTA of the tcgen05 value-net kernel (leaf_mlp_tc3.cuh) from clock64() stamps (development aid).
Epilogue thread 0, per phase k: [0] accumulators ready, [1] in registers, [2] own sum of squares done, [3] row statistics
exchanged, [4] sub-chunk 0 (+ the previous tile's output rows) done, [5] all four sub-chunks issued, [6] A operand handed over.
MMA thread, per phase k: [0] D free, [1] layer 3 / refill issued and read back, [2] next accumulators issued."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rebel_b200 as rb
from rebel_b200.models import make_selfplay_net, flatten_state_dict


def main():
    D, F, K = 1, 6, 8192
    H = F ** D
    w = flatten_state_dict(make_selfplay_net(D, F, seed=0).state_dict())
    rng = np.random.RandomState(1)
    b = rng.rand(K, 2, H); b /= b.sum(-1, keepdims=True)
    for mode, name in ((rb.NET_TC_F16X2, "f16x2-gelu"), (rb.NET_TC_F16, "f32-gelu")):
        S = rb.WaveSolver(D, F, K, net_mode=mode)
        S.set_weights(w)
        S.begin(np.full(K, -1, np.int32), np.zeros(K, np.int32), b)
        S.run(3)
        S.net_trace()
        t = S.net_trace()
        e = t[:1024].reshape(128, 8); m = t[1024:2048].reshape(256, 4)
        n = int((e[:, 0] > 0).sum())
        d = np.diff(e[:n, 0])
        print(f"== {name}: {n} phases on CTA 0; cycles per phase median {np.median(d):.0f} (min {d.min()}, max {d.max()})")
        for k in range(8, 16):
            base = e[k, 0]
            prev_end = e[k - 1, 6] - base
            print(f"  phase {k}: prev end {prev_end}, ld {e[k,1]-base}, sumsq {e[k,2]-base}, stats {e[k,3]-base}, chunk0 {e[k,4]-base}, chunk3 {e[k,5]-base}, handover {e[k,6]-base}"
                  f" | MMA: Dfree {m[k,0]-base}, L3/readback {m[k,1]-base}, next issued {m[k,2]-base}")
        S.close()


if __name__ == "__main__":
    main()
