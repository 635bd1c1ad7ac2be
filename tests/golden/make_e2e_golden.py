"""Generates tests/golden/e2e_*.json by running the UNMODIFIED reference
(/root/reference/whisper_timestamped) on CPU over the oracle's stand-ins for its two missing
third-party dependencies (oracle/upstream/{whisper,dtw}).  Runs only in the build container
(the reference does not exist on the GPU box); the JSON fixtures are committed.

    python tests/golden/make_e2e_golden.py [case ...]

Inputs are fully synthetic and reproducible from the recipe stored in each fixture:
model = synthetic_state_dict(DIMS[name], seed), audio = synthetic_speech(duration, audio_seed).
"""
import importlib.util
import json
import contextlib
import io
import logging
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "upstream"))
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402
import torch  # noqa: E402
import whisper  # noqa: E402  (oracle stand-in)
import whisper_timestamped as ref  # noqa: E402  (the real reference)

assert ref.__file__.startswith("/root/reference"), ref.__file__


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


PKG = os.path.join(ROOT, "whisper-timestamped_b200", "whisper_timestamped")
zoo = _load(os.path.join(PKG, "model_zoo.py"), "wts_model_zoo")
sa = _load(os.path.join(PKG, "synthetic_audio.py"), "wts_synthetic_audio")

BENCH_KW = {"ts_offset": 4.5, "eot_logit": 14.5}        # == bench.py SYNTH_KW

CASES = {
    # name: (model, model kwargs, audio (duration, seed), transcribe kwargs)
    "tiny_en_30s": ("tiny.en", {}, (30.0, 7), {}),
    "tiny_75s_cond": ("tiny", {}, (75.0, 11), {"language": "en"}),
    "tiny_60s_nocond": ("tiny", {}, (60.0, 113), {"language": "en", "condition_on_previous_text": False}),
    "tiny_detect_lang": ("tiny", {}, (35.0, 13), {}),
    # < 30 s of audio with language detection: the first window is aligned against the detection mel, so the
    # reference applies NO padding mask (T.py:795-799, 708) although the window's own mel is zero-padded
    "tiny_detect_lang_short": ("tiny", {}, (20.0, 24), {}),
    "tiny_stuck": ("tiny", {"eot_logit": 4.0, "ts_offset": 0.5}, (45.0, 14), {"language": "en"}),
    "tiny_opts": ("tiny", {}, (50.0, 15), {"language": "fr", "remove_punctuation_from_words": True,
                                            "include_punctuation_in_confidence": True,
                                            "refine_whisper_precision": 0.2, "min_word_duration": 0.1}),
    "tiny_norefine": ("tiny", {}, (40.0, 16), {"language": "en", "refine_whisper_precision": 0.0}),
    "tiny_short": ("tiny", {}, (3.3, 17), {"language": "en"}),
    "tiny_ja_unspaced": ("tiny", {}, (40.0, 18), {"language": "ja"}),
    # two-pass ("naive") strategy, greedy, with and without trust in Whisper's timestamps
    "tiny_naive": ("tiny", {}, (75.0, 21), {"language": "en", "naive_approach": True, "temperature": 0.0}),
    "tiny_naive_notrust": ("tiny", {}, (65.0, 22), {"language": "en", "naive_approach": True, "temperature": 0.0,
                                                    "trust_whisper_timestamps": False}),
    "tiny_naive_opts": ("tiny", {}, (45.0, 23), {"language": "fr", "naive_approach": True, "temperature": 0.0,
                                                 "include_punctuation_in_confidence": True,
                                                 "remove_punctuation_from_words": True, "refine_whisper_precision": 0.2}),
    # upstream decoding strategies in front of the two-pass alignment (BASELINE config 4): beam search,
    # the README's "accurate" setting (beam + best-of sampling under temperature fallback; the log-prob threshold is moved
    # so that some windows pass at temperature 0 and others fall back), and plain best-of-n sampling
    "tiny_beam5": ("tiny", {}, (65.0, 25), {"language": "en", "beam_size": 5}),
    "tiny_accurate": ("tiny", {}, (75.0, 26), {"language": "en", "beam_size": 5, "best_of": 5,
                                               "temperature": (0.0, 0.2, 0.4, 0.6, 0.8, 1.0), "logprob_threshold": -1.2}),
    "tiny_bestof3": ("tiny", {}, (50.0, 27), {"language": "en", "temperature": 0.3, "best_of": 3}),
    # detect_disfluencies: "[*]" pseudo-words from the peak analysis of the attention rows
    "tiny_disfluencies": ("tiny", {}, (60.0, 12), {"language": "en", "detect_disfluencies": True}),
    "tiny_disfluencies_naive": ("tiny", {}, (70.0, 19), {"language": "en", "detect_disfluencies": True,
                                                         "naive_approach": True, "temperature": 0.0}),
    # explicit-list VAD: speech spans glued, times mapped back, `speech_activity` reported
    "tiny_vad_list": ("tiny", {}, (70.0, 19), {"language": "en", "vad": [(2.0, 21.5), (30.25, 52.0), (58.0, 66.4)]}),
    # stdout of `verbose=True` (T.py:323, 346, 817-820, 844-846, 1304): detection lines + one line per word, for the one-pass
    # strategy, the two-pass strategy and a VAD list (times printed after the mapping back to the original audio)
    "tiny_verbose_detect": ("tiny", {}, (35.0, 13), {"verbose": True}),
    "tiny_verbose_naive": ("tiny", {}, (45.0, 28), {"language": "en", "naive_approach": True, "temperature": 0.0, "verbose": True}),
    "tiny_verbose_vad": ("tiny", {}, (70.0, 19), {"verbose": True, "vad": [(2.0, 21.5), (30.25, 52.0), (58.0, 66.4)]}),
    # ---- the configurations bench.py measures (BASELINE.json configs 2-4), at their real dimensions
    # large-v3 with the bench recipe, sequential (two windows: the second-round / prompt carry-over path)
    "large_v3_45s": ("large-v3", BENCH_KW, (45.0, 31), {"language": "en"}),
    "base_60s": ("base", {}, (60.0, 32), {"language": "en"}),
    "medium_30s": ("medium", {}, (30.0, 33), {"language": "en"}),
}

# `chunks=` mode (the data-parallel unit of work): the reference run INDEPENDENTLY on every fixed cut
# with condition_on_previous_text=False; the product's transcribe(..., chunks=) must equal the stitched per-cut results.
# name: (model, model kwargs, audio (duration, seed), chunk seconds, transcribe kwargs)
CHUNK_CASES = {
    "chunks_tiny_100s": ("tiny", {}, (100.0, 41), 30.0, {"language": "en"}),
    # the first 5 minutes of bench.py's 1-h workload (make_audio: 300-s pieces, seed 1234 + k), large-v3, bench recipe
    "chunks_large_v3_bench300": ("large-v3", BENCH_KW, (300.0, 1234), 30.0, {"language": "en"}),
}


def build_model(name, seed=1234, **kw):
    dims = zoo.DIMS[name]
    sd = zoo.synthetic_state_dict(dims, seed=seed, **kw)
    model = whisper.Whisper(whisper.ModelDimensions(**dims.asdict()))
    model.load_state_dict(sd)
    mask = torch.zeros(dims.n_text_layer, dims.n_text_head, dtype=torch.bool)
    for l, h in zoo.ALIGNMENT_HEADS[name]:
        mask[l, h] = True
    model.register_buffer("alignment_heads", mask.to_sparse(), persistent=False)
    return model.eval()


def to_py(o):
    if isinstance(o, dict):
        return {k: to_py(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [to_py(v) for v in o]
    if isinstance(o, (np.floating,)):
        return float(o)
    if isinstance(o, (np.integer,)):
        return int(o)
    return o


class _Capture(logging.Handler):
    """Warnings the reference logs while it runs: part of the pinned behaviour (e.g. "Got inconsistent length for
    segment ... Some words have been ignored" on the too-much-text truncation path, T.py:1516-1535 / 993-994)."""

    def __init__(self):
        super().__init__(level=logging.WARNING)
        self.messages = []

    def emit(self, record):
        self.messages.append(record.getMessage())


def run_reference(model, audio, **kw):
    cap = _Capture()
    lg = logging.getLogger("whisper_timestamped")
    lg.addHandler(cap)
    out = io.StringIO()
    try:
        with contextlib.redirect_stdout(out):
            res = ref.transcribe(model, audio, **kw)
    finally:
        lg.removeHandler(cap)
    run_reference.stdout = out.getvalue()       # what the reference (and upstream under it) printed: pinned for `verbose`
    return to_py(res), cap.messages


def main():
    names = sys.argv[1:] or (list(CASES) + list(CHUNK_CASES))
    models = {}

    def get_model(mname, mkw):
        key = (mname, json.dumps(mkw, sort_keys=True))
        if key not in models:
            models.clear()                       # one big model in memory at a time
            models[key] = build_model(mname, **mkw)
        return models[key]

    for case in names:
        if case in CHUNK_CASES:
            mname, mkw, (dur, aseed), chunk_s, tkw = CHUNK_CASES[case]
            model = get_model(mname, mkw)
            audio = sa.synthetic_speech(dur, seed=aseed)
            step = int(round(chunk_s * 16000))
            t0 = time.time()
            cuts = []
            for s in range(0, len(audio), step):
                res, warns = run_reference(model, audio[s:s + step], condition_on_previous_text=False, **tkw)
                cuts.append({"offset": s / 16000.0, "result": res, "warnings": warns})
                print(f"  {case} cut @{s / 16000.0:.0f}s: {len(res['segments'])} segments, {len(warns)} warnings, "
                      f"{time.time() - t0:.0f}s", flush=True)
            dt = time.time() - t0
            out = {"case": case, "model": mname, "model_seed": 1234, "model_kwargs": mkw, "audio": [dur, aseed],
                   "chunks": chunk_s, "transcribe_kwargs": tkw, "reference_version": ref.__version__,
                   "cpu_seconds": round(dt, 2), "cuts": cuts}
            with open(os.path.join(HERE, f"{case}.json"), "w") as f:
                json.dump(out, f, indent=1, ensure_ascii=False)
            continue
        mname, mkw, (dur, aseed), tkw = CASES[case]
        model = get_model(mname, mkw)
        audio = sa.synthetic_speech(dur, seed=aseed)
        t0 = time.time()
        res, warns = run_reference(model, audio, **tkw)
        dt = time.time() - t0
        out = {"case": case, "model": mname, "model_seed": 1234, "model_kwargs": mkw, "audio": [dur, aseed],
               "transcribe_kwargs": tkw, "reference_version": ref.__version__, "cpu_seconds": round(dt, 2),
               "warnings": warns, "result": res}
        if tkw.get("verbose") is not None or case == "tiny_detect_lang":
            out["stdout"] = run_reference.stdout
        with open(os.path.join(HERE, f"e2e_{case}.json"), "w") as f:
            json.dump(out, f, indent=1, ensure_ascii=False)
        nseg = len(res["segments"])
        nw = sum(len(s.get("words", [])) for s in res["segments"])
        ntok = sum(len(s["tokens"]) for s in res["segments"])
        print(f"{case}: {nseg} segments, {nw} words, {ntok} tokens, {len(warns)} warnings, {dt:.1f}s", flush=True)


if __name__ == "__main__":
    main()
