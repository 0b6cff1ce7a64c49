#!/bin/bash
# GPU box: recursive-eval parity tests + config-5 timing (2x3f) next to the compiled reference on the host cores.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rela_module.py -x -q -m gpu -k "recursive_eval" 2>&1 | tail -5
for net in "" "--random_net_seed 0"; do
  timeout 600 python -m rebel_b200.recursive_eval --num_dice 2 --num_faces 3 --subgame_iters 1024 --num_repeats 64 --cfr --no_full_tree $net 2>&1 | tail -9
done
timeout 600 python - <<'PY'
import sys, time
sys.path.insert(0, ".")
import numpy as np
from oracle.oracle import Oracle, available
from rebel_b200.models import flatten_state_dict, make_selfplay_net
if available("ref_fast"):
    R = Oracle("ref_fast")
    w = flatten_state_dict(make_selfplay_net(2, 3, seed=0).state_dict())
    for name, nw in (("zero", None), ("net2", w)):
        t = time.time(); R.sampled_strategy(2, 3, seed=0, num_iters=1024, net_w=nw); print("cpu reference 2x3f one repeat,", name, "net:", time.time() - t, "s on one core", flush=True)
PY
