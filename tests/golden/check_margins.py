"""Conditioning of the greedy goldens: the smallest gap between the best and second-best filtered log-probability over
EVERY decoded row of every window (including rows upstream later discards: they still feed avg_logprob and the seek).
A float32 GPU forward differs from the float32 CPU reference by ~1e-5 on these synthetic models, so a golden whose
gap is below ~1e-4 is a coin flip, not a parity test (round 2: tiny_60s_nocond had a 8.6e-6 gap at window 1, row 68).
Writes `min_top2_gap` into each fixture; tests/test_host_e2e.py requires >= 1e-4.

    python tests/golden/check_margins.py [fixture.json ...]          (CPU, through the oracle engine)
"""
import glob
import json
import logging
import os
import sys
from types import SimpleNamespace

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "whisper-timestamped_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402


def min_gap(g):
    import oracle.engine as OE
    from whisper_timestamped import model_zoo as zoo
    from whisper_timestamped.synthetic_audio import synthetic_speech
    from whisper_timestamped.transcribe import transcribe_timestamped
    dims = zoo.DIMS[g["model"]]
    sd = zoo.synthetic_state_dict(dims, seed=g["model_seed"], **g["model_kwargs"])
    heads = zoo.ALIGNMENT_HEADS[g["model"]]
    om = OE.build_oracle_model(dims, sd, heads)
    eng = OE.OracleEngine(om, heads, keep_logprobs=True)
    shim = SimpleNamespace(dims=dims, is_multilingual=om.is_multilingual, num_languages=om.num_languages)
    kw = dict(g["transcribe_kwargs"])
    if "chunks" in g:
        kw["chunks"] = g["chunks"]
    transcribe_timestamped(shim, synthetic_speech(*g["audio"]), engine=eng, **kw)
    worst = (float("inf"), -1, -1)
    for w, full in enumerate(eng.full_logprobs):
        top2 = torch.topk(full, 2, dim=-1).values
        gap = top2[:, 0] - top2[:, 1]
        gap = torch.where(torch.isfinite(gap), gap, torch.full_like(gap, float("inf")))
        r = int(gap.argmin())
        if float(gap[r]) < worst[0]:
            worst = (float(gap[r]), w, r)
    return worst


def main():
    logging.getLogger("whisper_timestamped").setLevel(logging.ERROR)
    paths = sys.argv[1:] or sorted(glob.glob(os.path.join(HERE, "e2e_*.json")) + glob.glob(os.path.join(HERE, "chunks_*.json")))
    for path in paths:
        g = json.load(open(path))
        kw = g["transcribe_kwargs"]
        if kw.get("beam_size") or isinstance(kw.get("temperature"), (list, tuple)) or (kw.get("temperature") or 0) > 0:
            continue                                    # beam search / sampling: no single argmax per row
        if any(b in g["model"] for b in ("medium", "large")) and os.environ.get("WTS_SLOW") != "1":
            continue
        gap, w, r = min_gap(g)
        g["min_top2_gap"] = {"gap": gap, "window": w, "row": r}
        with open(path, "w") as f:
            json.dump(g, f, indent=1, ensure_ascii=False)
        print(f"{os.path.basename(path)}: {gap:.2e} (window {w}, row {r})", flush=True)


if __name__ == "__main__":
    main()
