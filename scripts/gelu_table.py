"""Dumps the hardware tables of the packed-half GELU (cfrb_debug_gelu_table) to gpurun_out/gelu_table.npz and prints their error
statistics against exact arithmetic."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rebel_b200 as rb
from scipy.special import erf
S = rb.WaveSolver(1, 4, 1, net_mode=rb.NET_ZERO)
x, t = S.gelu_table(0)
_, g = S.gelu_table(1)
_, g32 = S.gelu_table(2)
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/gelu_table.npz", x=x, tanh=t, gelu=g, gelu_t32=g32)
xf = x.astype(np.float64); ok = np.isfinite(xf)
for lo, hi in ((-8, -4), (-4, -2), (-2, -1), (-1, -0.25), (-0.25, 0.25), (0.25, 1), (1, 2), (2, 4), (4, 8)):
    m = ok & (xf >= lo) & (xf < hi)
    e = t.astype(np.float64)[m] - np.tanh(xf[m])
    ideal = np.tanh(xf[m]).astype(np.float16).astype(np.float64) - np.tanh(xf[m])
    y = 2 * xf[m]
    exact = 0.5 * y * (1 + erf(y / np.sqrt(2)))
    eg = g.astype(np.float64)[m] - exact
    e32 = g32.astype(np.float64)[m] - exact
    er = exact.astype(np.float16).astype(np.float64) - exact
    print(f"hy/u in [{lo},{hi}): n {m.sum():5d} tanh err mean {e.mean():+.2e} rms {np.sqrt((e**2).mean()):.2e} max {np.abs(e).max():.2e} (ideal rounding rms {np.sqrt((ideal**2).mean()):.2e}) | "
          f"gelu(y=2hy) packed half: err mean {eg.mean():+.2e} rms {np.sqrt((eg**2).mean()):.2e} max {np.abs(eg).max():.2e} | "
          f"fp32 tanh: mean {e32.mean():+.2e} rms {np.sqrt((e32**2).mean()):.2e} max {np.abs(e32).max():.2e} | exact rounded to fp16: rms {np.sqrt((er**2).mean()):.2e}")
