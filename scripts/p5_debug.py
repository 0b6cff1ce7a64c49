"""P5 debugging aid: last-action distribution of the GPU data-generation loop for every value-net mode next to the reference's
(fixture datagen_stats.npz)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from rebel_b200 import rela
from rebel_b200.models import flatten_state_dict, make_selfplay_net
from test_rela_module import make_cfg, _last_action_stats, _chi2, game_dims

D, F = int(sys.argv[1]), int(sys.argv[2])
K, waves = 1024, int(sys.argv[3]) if len(sys.argv) > 3 else 48
A, H, Q = game_dims(D, F)
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "datagen_stats.npz"))
ca, cb = g[f"count_a_{D}x{F}"].astype(np.float64), g[f"count_b_{D}x{F}"].astype(np.float64)
np.set_printoptions(linewidth=250, precision=4, suppress=True)
print("ref a", ca / ca.sum()); print("ref b", cb / cb.sum(), "chi2 a-b", _chi2(ca, cb))
print("ref mean target a", g[f"val_sum_a_{D}x{F}"] / np.maximum(ca, 1) / H)
net = make_selfplay_net(D, F, seed=0)
w = torch.from_numpy(flatten_state_dict(net.state_dict()))
ALL = (("fp32", 1, 0), ("tc", 2, 0), ("tcx2", 3, 0), ("fp32-hostwalk", 1, 1))
want = os.environ.get("MODES", "fp32,tc,tcx2").split(",")
for name, mode, hw in [m for m in ALL if m[0] in want]:
    cfg = make_cfg(rela, D, F, concurrent_games=K, net_mode=mode, state_dtype=0, host_walk=hw)
    q, v = rela.run_selfplay_waves(cfg, 0, 123, waves, w)
    q = q.numpy().reshape(waves, K, 2, Q); v = v.numpy().reshape(waves, K, 2, H)
    starts = q[:, :, 0, 2:2 + A].sum(-1) == 0
    last_start = waves - 1 - np.argmax(starts[::-1], axis=0)
    keep = np.arange(waves)[:, None] < last_start[None, :]
    qk, vk = q[keep].reshape(-1, Q), v[keep].reshape(-1, H)
    cnt, vsum, lsum = _last_action_stats(qk, vk, A, net)
    print(name, len(qk), cnt / cnt.sum(), "chi2 vs a", round(_chi2(cnt, ca), 1), "vs b", round(_chi2(cnt, cb), 1))
    print("   mean target", vsum / np.maximum(cnt, 1) / H)
    # all examples (no game-completion filter)
    cnt2, _, _ = _last_action_stats(q.reshape(-1, Q), v.reshape(-1, H), A, net)
    print("   unfiltered  ", cnt2 / cnt2.sum())
