"""Where the time of a self-play wave goes: device-side loop through the ctypes C ABI (cfrb_selfplay_*), per wave the total
device time of the 1024 iterations and the value-net share (CUDA-event pairs around every 16th launch), the value-net rows."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rebel_b200 as rb
from rebel_b200.models import flatten_state_dict, make_selfplay_net


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dice", type=int, default=1); ap.add_argument("--faces", type=int, default=6)
    ap.add_argument("--games", type=int, default=8192); ap.add_argument("--iters", type=int, default=1024)
    ap.add_argument("--waves", type=int, default=8); ap.add_argument("--warm", type=int, default=6)
    ap.add_argument("--net", default="tcx2")
    a = ap.parse_args()
    mode = {"tcx2": rb.NET_TC_F16X2, "tc": rb.NET_TC_F16, "fp32": rb.NET_FP32, "zero": rb.NET_ZERO}[a.net]
    S = rb.WaveSolver(a.dice, a.faces, a.games, num_iters=a.iters, net_mode=mode)
    if mode != rb.NET_ZERO:
        S.set_weights(flatten_state_dict(make_selfplay_net(a.dice, a.faces, seed=0).state_dict()))
    S.selfplay_create(np.arange(a.games, dtype=np.uint32) * 1000000 + 7)
    for _ in range(a.warm):
        S.selfplay_wave()
    S.sync()
    S.set_profiling(16)
    out = []
    for _ in range(a.waves):
        S.selfplay_wave()
        S.sync()
        tot, net = S.last_run_ms()
        rows = S.leaf_rows
        out.append((tot, net, rows))
    S.set_profiling(0)
    # pipelined throughput without profiling
    for _ in range(3):
        S.selfplay_wave()
    S.sync()
    S.mark(0)
    for _ in range(a.waves):
        S.selfplay_wave()
    S.mark(1)
    ms = S.elapsed_ms(0, 1)
    tot = np.array([o[0] for o in out]); net = np.array([o[1] for o in out]); rows = np.array([o[2] for o in out])
    print(json.dumps({"game": f"{a.dice}x{a.faces}f", "games": a.games, "net": a.net, "wave_ms": float(tot.mean()), "net_ms": float(net.mean()),
                      "cfr_ms": float((tot - net).mean()), "rows": float(rows.mean()), "net_us_per_launch": float(net.mean() / a.iters * 1e3),
                      "cfr_us_per_launch": float((tot - net).mean() / a.iters * 1e3),
                      "pipelined_ms_per_wave": ms / a.waves, "subgame_iters_per_s": a.games * a.iters * a.waves / (ms * 1e-3)}))


if __name__ == "__main__":
    main()
