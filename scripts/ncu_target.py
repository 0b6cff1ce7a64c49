"""Tiny driver for ncu captures: one 1x6f K=8192 wave, a few iterations (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rebel_b200 as rb
from rebel_b200.models import make_selfplay_net, flatten_state_dict
D, F, K = 1, 6, 8192
state = rb.STATE_F32 if len(sys.argv) > 1 and sys.argv[1] == "f32" else rb.STATE_F64
w = flatten_state_dict(make_selfplay_net(D, F, seed=0).state_dict())
b = np.random.RandomState(1).rand(K, 2, 6); b /= b.sum(-1, keepdims=True)
S = rb.WaveSolver(D, F, K, net_mode=rb.NET_TC_F16X2, state_dtype=state)
S.set_weights(w)
S.begin(np.full(K, -1, np.int32), np.zeros(K, np.int32), b)
S.run(8); S.sync()
print("done", S.kernel_launches)
