"""Timing of the rela evaluation entry points (development aid)."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rebel_b200.rela as rela
from rebel_b200.models import make_selfplay_net

D, F, iters = (int(x) for x in (sys.argv[1:4] + ["1", "6", "16"][len(sys.argv) - 1:]))
path = os.path.join(tempfile.mkdtemp(), "net.pt")
torch.jit.script(make_selfplay_net(D, F, seed=0)).save(path)
cfg = rela.RecursiveSolvingParams()
cfg.num_dice, cfg.num_faces, cfg.net_mode = D, F, 1
sp = cfg.subgame_params
sp.num_iters, sp.max_depth, sp.linear_update, sp.use_cfr = iters, 2, True, True
for name, fn in (("compute_exploitability_with_net", lambda: rela.compute_exploitability_with_net(cfg, path)),
                 ("compute_stats_with_net", lambda: rela.compute_stats_with_net(cfg, path)),
                 ("compute_exploitability_with_net (2nd)", lambda: rela.compute_exploitability_with_net(cfg, path))):
    t = time.time(); r = fn(); print(f"{name}: {time.time() - t:.2f} s -> {r}", flush=True)
