"""
ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates the attention post-processing of `perform_word_alignment`
(/root/reference/whisper_timestamped/transcribe.py:1540-1568): slice frames, select the
alignment heads, scipy median filter (1,1,9), softmax over frames, mean over heads, L2 norm
over tokens, negate → float64, padding mask, `weights[0,0] = weights.min()`.

`scipy.ndimage.median_filter` and the torch CPU ops are the very calls the reference makes,
so this file is "the reference's arithmetic, same libraries" rather than a re-derivation.
"""
import numpy as np
import torch
from scipy.ndimage import median_filter


def attn_cost(qk_heads, start_token: int, end_token: int, max_duration=None,
              medfilt_width: int = 9, qk_scale: float = 1.0) -> np.ndarray:
    """qk_heads: float32 array/tensor [N, T, >=end_token] holding the *selected* heads'
    pre-softmax cross-attention rows (what transcribe.py:1545 stacks).

    Returns the float64 [T, F] local-cost matrix handed to dtw.dtw (transcribe.py:1581)."""
    w = torch.as_tensor(np.asarray(qk_heads), dtype=torch.float32)
    w = w[..., start_token:end_token]                                  # T.py:1540
    w = median_filter(w.numpy(), (1, 1, medfilt_width))                # T.py:1546
    w = torch.tensor(w * qk_scale).softmax(dim=-1)                     # T.py:1547
    w = w.mean(axis=(0))                                               # T.py:1548
    w = w / w.norm(dim=-2, keepdim=True)                               # T.py:1549
    w = -w.double().numpy()                                            # T.py:1550
    worse_weight = 0
    if max_duration:                                                   # T.py:1561-1565
        if start_token >= max_duration:
            pass  # reference only logs a warning
        else:
            w[:-1, max_duration:] = worse_weight
    w[0, 0] = w.min()                                                  # T.py:1568
    return w


def find_start_padding(mfcc: torch.Tensor):
    """transcribe.py:1795-1805 — first all-equal-to-last (zero) mel column, or None."""
    last = mfcc[0, :, -1]
    if torch.min(last) == torch.max(last) == 0:
        idx = mfcc.shape[-1] - 2
        while idx > 0:
            if not torch.equal(mfcc[0, :, idx], last):
                return idx + 1
            idx -= 1
        return 0
    return None
