"""Probe of the tcgen05 value-net kernel: layer-by-layer taps vs numpy on fp16-rounded operands (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rebel_b200 as rb
from rebel_b200.models import make_selfplay_net, flatten_state_dict

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def h16(x):
    return x.astype(np.float16).astype(np.float32)


def gelu(x):
    return torch.nn.functional.gelu(torch.from_numpy(x)).numpy()


def ln(x, g, b):
    m = x.mean(-1, keepdims=True); v = x.var(-1, keepdims=True)
    return (x - m) / np.sqrt(v + 1e-5) * g + b


def main():
    for (D, F) in [(1, 6), (2, 3), (1, 4)]:
        A = 1 + 2 * D * F; H = F ** D; Q = 2 + A + 2 * H
        net = make_selfplay_net(D, F, seed=0)
        sd = {k: v.numpy() for k, v in net.state_dict().items()}
        w = flatten_state_dict(net.state_dict())
        K = 8
        rng = np.random.RandomState(0)
        b = rng.rand(K, 2, H); b /= b.sum(-1, keepdims=True)
        S = rb.WaveSolver(D, F, K, net_mode=rb.NET_TC_F16)
        S.set_weights(w)
        S.begin(np.full(K, -1, np.int32), np.zeros(K, np.int32), b)
        S.run(1)
        q, o, sc = S.leaf_io()
        d1, d2 = S.net_taps()
        n = min(128, q.shape[0])
        W1, W2, W3 = h16(sd["body.0.weight"]), h16(sd["body.4.weight"]), h16(sd["output.weight"])
        e1 = q[:n] @ W1.T + h16(sd["body.0.bias"])          # bias 1 rides on the tensor cores (constant-1 query column)
        print(f"== {D}x{F}f rows={q.shape[0]} Q={Q}")
        print("  layer1 acc: max|d1-exp| =", np.abs(d1[:n] - e1).max(), " scale", np.abs(e1).max())
        a2 = h16(gelu(ln(d1[:n], sd["body.1.weight"], sd["body.1.bias"])))
        e2 = a2 @ W2.T + h16(sd["body.4.bias"])
        print("  layer2 acc (teacher-forced from gpu d1): max|d2-exp| =", np.abs(d2[:n] - e2).max(), " scale", np.abs(e2).max())
        a3 = h16(gelu(ln(d2[:n], sd["body.5.weight"], sd["body.5.bias"])))
        e3 = a3 @ W3.T + sd["output.bias"]
        print("  layer3 out (teacher-forced from gpu d2): max|out-exp| =", np.abs(o[:n] - e3).max(), " scale", np.abs(e3).max())
        with torch.no_grad():
            ref = net(torch.from_numpy(q)).numpy()
        print("  end-to-end vs torch fp32 on the same (fp16-rounded) queries: max abs", np.abs(ref - o).max(), "rel", np.abs(ref - o).max() / np.abs(ref).max())
        np.savez(os.path.join(OUT, f"tc_probe_{D}x{F}.npz"), q=q, o=o, d1=d1, d2=d2, e1=e1, e2=e2, e3=e3, ref=ref)
        S.close()
        # compare a short trajectory with the fp32 path
        Z = rb.WaveSolver(D, F, K, net_mode=rb.NET_FP32); Z.set_weights(w)
        T = rb.WaveSolver(D, F, K, net_mode=rb.NET_TC_F16); T.set_weights(w)
        for s_ in (Z, T):
            s_.begin(np.full(K, -1, np.int32), np.zeros(K, np.int32), b); s_.run(8)
        a, c = Z.fetch(("root_means", "avg")), T.fetch(("root_means", "avg"))
        print("  8 iters tc vs fp32: mu maxabs", np.abs(a["root_means"] - c["root_means"]).max(), "avg maxabs", np.abs(a["avg"] - c["avg"]).max())
        Z.close(); T.close()
    # timing
    D, F, K = 1, 6, 8192
    H = 6
    w = flatten_state_dict(make_selfplay_net(D, F, seed=0).state_dict())
    b = np.random.RandomState(1).rand(K, 2, H); b /= b.sum(-1, keepdims=True)
    for mode in (rb.NET_TC_F16,):
        S = rb.WaveSolver(D, F, K, net_mode=mode); S.set_weights(w)
        S.begin(np.full(K, -1, np.int32), np.zeros(K, np.int32), b)
        S.run(4); S.sync()
        S.set_profiling(True)
        S.run(64); S.sync()
        tot, net = S.last_run_ms()
        print(f"timing tc K={K}: 64 iters {tot:.2f} ms total, net {net:.2f} ms -> {K*64/(tot*1e-3)/1e6:.2f} M subgame-iters/s; net per launch {net/64*1e3:.1f} us, cfr per launch {(tot-net)/65*1e3:.1f} us")
        S.close()


if __name__ == "__main__":
    main()
