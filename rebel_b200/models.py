"""Value net for Liar's Dice public-belief states — the Python-side PyTorch module of the drop-in.

Architecture, parameter names and state_dict order follow the reference's ``Net2``
(cfvpy/models.py:64-94 with ``build_mlp`` :20-53): ``n_layers`` x [Linear -> LayerNorm -> GELU(erf)] and an output
Linear whose initial weights are scaled by 0.01.  Training of this module stays in PyTorch; inference during
data generation is done by the CUDA kernels behind ``cfrb_set_weights`` (include/cfrb200.h), which take the
parameters as one flat fp32 buffer in state_dict order (``flatten_state_dict``).
"""
import numpy as np
import torch
from torch import nn


def output_size(num_faces, num_dice):
    return num_faces ** num_dice


def input_size(num_faces, num_dice):
    # player, traverser, one-hot last bid (incl. liar), beliefs of both players (models.py:56-61)
    return 2 + (2 * num_faces * num_dice + 1) + 2 * output_size(num_faces, num_dice)


class GELU(nn.Module):
    def forward(self, x):
        return nn.functional.gelu(x)


class Net2(nn.Module):
    def __init__(self, *, num_faces, num_dice, n_hidden=256, use_layer_norm=False, dropout=0, n_layers=3):
        super().__init__()
        width = input_size(num_faces, num_dice)
        blocks = []
        for _ in range(n_layers):
            blocks += [
                nn.Linear(width, n_hidden),
                nn.LayerNorm(n_hidden) if use_layer_norm else nn.Sequential(),
                GELU(),
                nn.Dropout(dropout) if dropout > 0 else nn.Sequential(),
            ]
            width = n_hidden
        self.body = nn.Sequential(*blocks)
        self.output = nn.Linear(width, output_size(num_faces, num_dice))
        with torch.no_grad():
            self.output.weight.data *= 0.01
            self.output.bias *= 0.01

    def forward(self, packed_input: torch.Tensor):
        return self.output(self.body(packed_input))


def make_selfplay_net(num_dice, num_faces, seed=0):
    """The data-generation net of conf/c02_selfplay/liars_sp.yaml:28-33 (256 x 2, LayerNorm), seeded init."""
    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    net = Net2(num_faces=num_faces, num_dice=num_dice, n_hidden=256, n_layers=2, use_layer_norm=True)
    torch.random.set_rng_state(gen_state)
    return net.eval()


FLAT_ORDER = ("body.0.weight", "body.0.bias", "body.1.weight", "body.1.bias", "body.4.weight", "body.4.bias",
              "body.5.weight", "body.5.bias", "output.weight", "output.bias")


def flatten_state_dict(state_dict):
    """Flat fp32 buffer in the order cfrb_set_weights expects (2 hidden layers + LayerNorm)."""
    missing = [k for k in FLAT_ORDER if k not in state_dict]
    if missing:
        raise ValueError(f"state_dict is not a 2-layer LayerNorm Net2 (missing {missing})")
    extra = [k for k in state_dict if k not in FLAT_ORDER]
    if extra:   # e.g. Net2's default n_layers=3 adds body.8 / body.9: never evaluate a truncated network
        raise ValueError(f"state_dict has parameters the accelerated Net2(n_hidden=256, n_layers=2, use_layer_norm=True) does not: {extra}")
    if tuple(state_dict["body.4.weight"].shape) != (256, 256):
        raise ValueError(f"n_hidden must be 256, got body.4.weight {tuple(state_dict['body.4.weight'].shape)}")
    return np.concatenate([state_dict[k].detach().to(torch.float32).cpu().numpy().ravel() for k in FLAT_ORDER])
