"""Debug helper: worst relative error of the GPU SC/TC against the float64 spec on a full-size
synthetic clip with long temporal chunks (error accumulation of the running coefficient sum)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from elvis_b200 import ops
from elvis_b200.synth import synth_yuv420
from oracle import spec_scoring
T, H, W = int(sys.argv[1]), 1088, 1920
for chunk in sys.argv[2:]:
    os.environ["ELVIS_SCORE_CHUNK"] = chunk
    y = synth_yuv420(T, H, W, seed=5, device="cuda").y
    sc, tc, _ = ops.score_sc_tc(y, 16)
    rsc, rtc = spec_scoring.sc_tc(y.cpu().numpy(), 16)
    sc, tc = sc.cpu().numpy().astype(np.float64), tc.cpu().numpy().astype(np.float64)
    rel = np.abs(sc - rsc) / np.maximum(rsc, 1e-30)
    i = np.unravel_index(rel.argmax(), rel.shape)
    nz = rtc > 0
    print(f"chunk {chunk}: SC max rel {rel.max():.3e} at {i} (spec {rsc[i]:.5f}), p99.9 {np.quantile(rel, 0.999):.2e}, "
          f"max abs {np.abs(sc - rsc).max():.2e}; TC max rel {(np.abs(tc - rtc)[nz] / rtc[nz]).max():.2e}; SC range {rsc.min():.4f}..{rsc.max():.2f}")
