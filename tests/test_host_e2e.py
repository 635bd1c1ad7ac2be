"""Product HOST logic (windows.py / words.py / transcribe.py) vs the golden outputs of the
UNMODIFIED reference (tests/golden/e2e_*.json, produced by tests/golden/make_e2e_golden.py).
The device work is done by the test-only OracleEngine (CPU stand-ins), so this runs without a GPU
and checks that the offline replay of the reference's hook state machine gives the same tokens,
segments, words, timestamps and confidences."""
import glob
import contextlib
import io
import json
import logging
import os
from types import SimpleNamespace

import numpy as np
import pytest

from whisper_timestamped import model_zoo as zoo
from whisper_timestamped.synthetic_audio import synthetic_speech
from whisper_timestamped.transcribe import transcribe_timestamped

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = sorted(glob.glob(os.path.join(HERE, "golden", "e2e_*.json")))
CHUNK_CASES = sorted(glob.glob(os.path.join(HERE, "golden", "chunks_*.json")))
# goldens at medium / large-v3 dimensions take minutes per window through the fp32 CPU stand-in: they are GPU tests
# (tests/test_gpu_model.py); on the CPU box they only run with WTS_SLOW=1
BIG = ("medium", "large")


def _is_big(path):
    return any(b in os.path.basename(path) for b in BIG) and os.environ.get("WTS_SLOW") != "1"


class CaptureWarnings(logging.Handler):
    """Collects what the product logs on logger "whisper_timestamped" (the goldens hold the reference's)."""

    def __init__(self):
        super().__init__(level=logging.WARNING)
        self.messages = []

    def emit(self, record):
        self.messages.append(record.getMessage())

    def __enter__(self):
        logging.getLogger("whisper_timestamped").addHandler(self)
        return self

    def __exit__(self, *exc):
        logging.getLogger("whisper_timestamped").removeHandler(self)


def norm_warnings(msgs):
    """First line only (the reference appends decoded text that depends on nothing else), as a sorted multiset:
    the product replays the state machine offline, so the ORDER of messages differs, not their content."""
    return sorted(m.split("\n")[0].strip() for m in msgs)


def run_case(path, **extra):
    from oracle_engine import OracleEngine, build_oracle_model
    g = json.load(open(path))
    dims = zoo.DIMS[g["model"]]
    sd = zoo.synthetic_state_dict(dims, seed=g["model_seed"], **g["model_kwargs"])
    heads = zoo.ALIGNMENT_HEADS[g["model"]]
    om = build_oracle_model(dims, sd, heads)
    eng = OracleEngine(om, heads)
    shim = SimpleNamespace(dims=dims, is_multilingual=om.is_multilingual, num_languages=om.num_languages)
    audio = synthetic_speech(*g["audio"])
    if "chunks" in g:
        extra = dict(extra, chunks=g["chunks"])
    out = io.StringIO()
    with CaptureWarnings() as cap, contextlib.redirect_stdout(out):
        res = transcribe_timestamped(shim, audio, engine=eng, **g["transcribe_kwargs"], **extra)
    res["_warnings"] = cap.messages
    res["_stdout"] = out.getvalue()
    return g, res


def stitch_cuts(g):
    """What `chunks=` must reproduce: the reference's result on every cut, shifted by the cut's offset (the same
    arithmetic as `offset = seek * HOP / SR`, T.py:959-962) and concatenated; ids renumbered over the recording."""
    segs, text, warns = [], [], []
    for cut in g["cuts"]:
        off = cut["offset"]
        for s in cut["result"]["segments"]:
            s = json.loads(json.dumps(s))
            s["start"] = round(s["start"] + off, 2)
            s["end"] = round(s["end"] + off, 2)
            s["seek"] = s["seek"] + int(round(off * 100))
            for w in s.get("words", []):
                w["start"] = round(w["start"] + off, 2)
                w["end"] = round(w["end"] + off, 2)
            s["id"] = len(segs)
            segs.append(s)
        text.append(cut["result"]["text"])
        warns.extend(cut["warnings"])
    return {"text": "".join(text), "segments": segs, "language": g["cuts"][0]["result"]["language"]}, warns


def compare(res, ref, conf_tol=0.0015, time_tol=0.0, prob_tol=1e-5):
    assert res["language"] == ref["language"]
    assert res["text"] == ref["text"]
    assert len(res["segments"]) == len(ref["segments"])
    for a, b in zip(res["segments"], ref["segments"]):
        assert a["tokens"] == b["tokens"], (a["id"], a["tokens"], b["tokens"])
        assert a["text"] == b["text"] and a["seek"] == b["seek"] and a["id"] == b["id"]
        for k in ("start", "end"):
            assert abs(a[k] - b[k]) <= time_tol + 1e-9, (a["id"], k, a[k], b[k])
        for k in ("avg_logprob", "no_speech_prob", "compression_ratio", "temperature"):
            assert abs(a[k] - b[k]) <= 1e-4 * max(1.0, abs(b[k])), (a["id"], k, a[k], b[k])
        assert ("confidence" in a) == ("confidence" in b)
        if "confidence" in b:
            assert abs(a["confidence"] - b["confidence"]) <= conf_tol
        wa, wb = a.get("words", []), b.get("words", [])
        assert [w["text"] for w in wa] == [w["text"] for w in wb], a["id"]
        for x, y in zip(wa, wb):
            assert abs(x["start"] - y["start"]) <= time_tol + 1e-9 and abs(x["end"] - y["end"]) <= time_tol + 1e-9, \
                (a["id"], x, y)
            assert abs(x["confidence"] - y["confidence"]) <= conf_tol, (a["id"], x, y)
    assert ("speech_activity" in res) == ("speech_activity" in ref)
    if "speech_activity" in ref:
        assert res["speech_activity"] == ref["speech_activity"]
    if "language_probs" in ref:
        for k, v in ref["language_probs"].items():
            assert abs(res["language_probs"][k] - v) < prob_tol


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[4:-5] for p in CASES])
def test_host_logic_matches_reference_golden(path):
    if _is_big(path):
        pytest.skip("large model: GPU test (WTS_SLOW=1 runs it through the CPU stand-in)")
    g, res = run_case(path)
    if "min_top2_gap" in g:
        # a greedy golden is only a parity test when no decoded row is a near-tie (tests/golden/check_margins.py)
        assert g["min_top2_gap"]["gap"] >= 1e-4, g["min_top2_gap"]
    compare(res, g["result"])
    if "warnings" in g:
        assert norm_warnings(res["_warnings"]) == norm_warnings(g["warnings"])
    if "stdout" in g:
        # what the reference (and upstream under it) prints: `verbose=True` word lines, language-detection lines
        assert res["_stdout"] == g["stdout"]


@pytest.mark.parametrize("path", CHUNK_CASES, ids=[os.path.basename(p)[:-5] for p in CHUNK_CASES])
def test_chunks_mode_equals_reference_on_every_cut(path):
    """transcribe(..., chunks=30) == the unmodified reference run independently on every 30-s cut
    (condition_on_previous_text=False), after the offset shift — SURVEY.md §8(e)'s definition of the sharded unit."""
    if _is_big(path):
        pytest.skip("large model: GPU test (WTS_SLOW=1 runs it through the CPU stand-in)")
    g, res = run_case(path)
    ref, warns = stitch_cuts(g)
    compare(res, ref, time_tol=1e-6)
    assert norm_warnings(res["_warnings"]) == norm_warnings(warns)


def test_vad_convert_timestamps_hand_cases():
    """Explicit-list VAD remap (restating T.py:2158-2200): glued-axis times back to the original axis."""
    from whisper_timestamped.vad import check_vad_method, convert_timestamps
    spans = [(2.0, 21.5), (30.25, 52.0), (58.0, 66.4)]           # glued lengths 19.5, 21.75, 8.4
    assert convert_timestamps(spans, 0.0) == 2.0
    assert convert_timestamps(spans, 19.5) == 21.5                # last instant of the first span
    assert convert_timestamps(spans, 19.6) == 30.35               # 0.1 s into the second span
    assert convert_timestamps(spans, 41.25 + 1.0) == 59.0
    assert convert_timestamps(spans, 100.0) == 116.75             # beyond the end: shifted by the removed 16.75 s, not clamped
    # a pair straddling the first cut: candidates are [21.0, 21.5] (first span, end clamped) and [30.25, 30.75] (second
    # span, start clamped); both keep 0.5 s of the 1.0 s, the stable sort keeps the first
    assert convert_timestamps(spans, 19.0, 20.0) == [21.0, 21.5]
    assert convert_timestamps(spans, 19.4, 21.0) == [30.25, 31.75]   # here the second span preserves more of the duration
    assert check_vad_method(False) is None and check_vad_method(None) is None
    assert check_vad_method([[0, 1.5], (2, 3)]) == [(0, 1.5), (2, 3)]
    with pytest.raises(NotImplementedError):
        check_vad_method("silero")
    with pytest.raises(NotImplementedError):
        check_vad_method(True)
    with pytest.raises(AssertionError):
        check_vad_method([(0, 1, 2)])
