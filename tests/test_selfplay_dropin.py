"""The reference's UNMODIFIED cfvpy/selfplay.py against the rebel_b200 `rela` module (selfplay.py drops in unchanged).

selfplay.py imports hydra-era packages that are not in this image (omegaconf, pytorch_lightning, heyhi's launcher); they are
stubbed with empty modules — none of them is on the data-generation path — and `cfvpy.rela` resolves to rebel_b200.rela.  Runs only
where the reference checkout exists (the build container); nothing of it is copied into the repository."""
import importlib.util
import os
import sys
import types

import pytest
import torch

REF = "/root/reference/cfvpy"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")


@pytest.fixture(scope="module")
def selfplay():
    import rebel_b200.rela as rela
    saved = {k: sys.modules.get(k) for k in ("cfvpy", "cfvpy.rela", "cfvpy.models", "cfvpy.utils", "cfvpy.selfplay", "heyhi", "omegaconf",
                                             "omegaconf.dictconfig", "pytorch_lightning", "pytorch_lightning.logging")}
    pkg = types.ModuleType("cfvpy"); pkg.__path__ = [REF]
    heyhi = types.ModuleType("heyhi"); heyhi.is_on_slurm = lambda: False
    oc = types.ModuleType("omegaconf"); ocd = types.ModuleType("omegaconf.dictconfig")
    ocd.DictConfig = type("DictConfig", (dict,), {}); oc.dictconfig = ocd
    pl = types.ModuleType("pytorch_lightning"); pll = types.ModuleType("pytorch_lightning.logging"); pl.logging = pll
    sys.modules.update({"cfvpy": pkg, "cfvpy.rela": rela, "heyhi": heyhi, "omegaconf": oc, "omegaconf.dictconfig": ocd,
                        "pytorch_lightning": pl, "pytorch_lightning.logging": pll})
    pkg.rela = rela
    spec = importlib.util.spec_from_file_location("cfvpy.selfplay", os.path.join(REF, "selfplay.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["cfvpy.selfplay"] = mod
    spec.loader.exec_module(mod)
    yield mod
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


def test_create_mdp_config_fills_our_params(selfplay):
    """create_mdp_config (selfplay.py:587-610): hasattr / setattr over cfg.env, recursing into subgame_params; B200 knobs are
    ordinary extra keys; unknown keys raise like with the reference module."""
    import rebel_b200.rela as rela
    env = {"num_dice": 1, "num_faces": 6, "random_action_prob": 0.25, "sample_leaf": True,
           "subgame_params": {"num_iters": 1024, "max_depth": 2, "linear_update": True, "use_cfr": True},
           "concurrent_games": 4096, "net_mode": 3}
    cfg = selfplay.create_mdp_config(env)
    assert isinstance(cfg, rela.RecursiveSolvingParams)
    assert (cfg.num_dice, cfg.num_faces, cfg.sample_leaf, cfg.concurrent_games) == (1, 6, True, 4096)
    assert abs(cfg.random_action_prob - 0.25) < 1e-7
    sp = cfg.subgame_params
    assert (sp.num_iters, sp.max_depth, sp.linear_update, sp.use_cfr) == (1024, 2, True, True)
    with pytest.raises(RuntimeError, match="Cannot find key"):
        selfplay.create_mdp_config({"no_such_knob": 1})
    assert isinstance(selfplay.create_mdp_config(None), rela.RecursiveSolvingParams)


def test_reference_model_builder_feeds_our_model_locker(selfplay):
    """_build_model (selfplay.py:31-50) with the YAML's model block (liars_sp.yaml:28-33) produces the TorchScript Net2 our
    ModelLocker accepts; the reference's default Net2 (n_layers=3) is refused instead of being truncated."""
    import rebel_b200.rela as rela
    ns = types.SimpleNamespace
    env = ns(num_faces=6, num_dice=1)
    good = selfplay._build_model("cpu", env, ns(name="Net2", kwargs=dict(n_hidden=256, use_layer_norm=True, n_layers=2)), jit=True)
    locker = rela.ModelLocker([good], "cuda:0")
    assert locker.version == 1
    q = torch.zeros(4, 2 + 13 + 12)
    assert good(q).shape == (4, 6)
    assert torch.equal(selfplay.get_last_action_index(q, 13), torch.full((4,), 13))      # all-zero one-hot = "initial"
    deep = selfplay._build_model("cpu", env, ns(name="Net2", kwargs=dict(n_hidden=256, use_layer_norm=True)), jit=True)
    with pytest.raises(RuntimeError, match="unexpected parameter"):
        rela.ModelLocker([deep], "cuda:0")
