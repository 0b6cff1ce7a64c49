"""Rows and device time per self-play wave from the start of a generator: how long the all-games-start-together transient lasts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rebel_b200 as rb
from rebel_b200.models import flatten_state_dict, make_selfplay_net
K, iters, waves = 8192, 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 40
S = rb.WaveSolver(1, 6, K, num_iters=iters, net_mode=rb.NET_TC_F16X2)
S.set_weights(flatten_state_dict(make_selfplay_net(1, 6, seed=0).state_dict()))
S.selfplay_create(np.arange(K, dtype=np.uint32) * 1000000)
out = []
S.mark(0)
for w in range(waves):
    S.selfplay_wave()
    S.mark(1 + (w & 1))
    S.sync()
    ms = S.elapsed_ms(2 - (w & 1) if w else 0, 1 + (w & 1))
    lb, _ = S.wave_roots()
    out.append((w, ms, S.leaf_rows, int((lb == -1).sum())))
for w, ms, rows, roots in out:
    print(f"wave {w:3d}: {ms:7.2f} ms, value-net rows {rows}, games at the initial state {roots}")
