"""Deterministic synthetic 16 kHz audio (no dataset, no ffmpeg in this environment): bursts of
harmonic, syllable-modulated "speech" separated by low-level noise gaps.  Used by bench.py and
the tests; real audio goes through the same code paths."""
import numpy as np

SAMPLE_RATE = 16000


def synthetic_speech(duration_s: float, seed: int = 1234, speech_schedule=None) -> np.ndarray:
    """float32 mono waveform in [-1, 1].  `speech_schedule`: optional list of (start, end) seconds
    that contain speech (everything else is the -60 dBFS noise floor)."""
    rng = np.random.default_rng(seed)
    n = int(round(duration_s * SAMPLE_RATE))
    out = (rng.standard_normal(n) * 1e-3).astype(np.float32)
    if speech_schedule is None:
        speech_schedule = []
        t = 0.0
        while t < duration_s:
            gap = rng.uniform(0.2, 2.0)
            seg = rng.uniform(1.0, 6.0)
            s, e = t + gap, min(t + gap + seg, duration_s)
            if e > s:
                speech_schedule.append((s, e))
            t = e
    for (s, e) in speech_schedule:
        i0, i1 = int(s * SAMPLE_RATE), min(int(e * SAMPLE_RATE), n)
        if i1 <= i0:
            continue
        m = i1 - i0
        tt = np.arange(m) / SAMPLE_RATE
        f0 = rng.uniform(90, 220) * (1.0 + 0.05 * np.sin(2 * np.pi * rng.uniform(0.5, 2.0) * tt))
        phase = 2 * np.pi * np.cumsum(f0) / SAMPLE_RATE
        sig = np.zeros(m)
        for h in range(1, 12):
            sig += (1.0 / h) * np.sin(h * phase + rng.uniform(0, 2 * np.pi)) * rng.uniform(0.3, 1.0)
        syll = 0.5 * (1 + np.sin(2 * np.pi * rng.uniform(3.0, 5.0) * tt + rng.uniform(0, 6.28)))
        env = np.minimum(1.0, np.minimum(tt, tt[::-1]) / 0.05)
        gain = rng.uniform(0.3, 1.0) * 0.25
        out[i0:i1] += (gain * env * syll ** 2 * sig / 3.0).astype(np.float32)
    return np.clip(out, -1.0, 1.0).astype(np.float32)
