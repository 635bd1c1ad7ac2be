"""Host-only text writers against the reference's OWN fixtures:
`make_subtitles.split_long_segments` / `write_srt` / `write_vtt` replay TestMakeSubtitles
(/root/reference/tests/test_transcribe.py:619-650, expected tests/expected/split_subtitles/*), and `write_csv` /
`write_tsv` reproduce the csv / tsv files the reference's command line wrote next to its `.words.json` results
(tests/expected/punctuations_{yes,no}).  Fixtures copied by tests/golden/make_subtitles_vectors.py."""
import io
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(HERE, "golden", "subtitles")
PKG = os.path.join(os.path.dirname(HERE), "whisper-timestamped_b200", "whisper_timestamped")


def _load(name):
    # loaded by file path: these modules are host-only and must work without importing the CUDA package
    import importlib.util
    spec = importlib.util.spec_from_file_location("wts_" + name, os.path.join(PKG, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("length", [6, 20, 50])
@pytest.mark.parametrize("name", ["smartphone.mp3", "no_punctuations.mp3", "yes_punctuations.mp3"])
def test_split_long_segments_matches_reference_fixtures(name, length):
    ms = _load("make_subtitles")
    transcript = json.load(open(os.path.join(G, f"in_{name}.words.json"), encoding="utf-8"))
    segments = ms.split_long_segments(transcript["segments"], length,
                                      use_space=transcript["language"] not in ms.LANGUAGES_WITHOUT_SPACES)
    expected_stem = name.split("_")[-1]                      # the reference maps no_/yes_punctuations to one fixture
    for ext, writer in (("srt", ms.write_srt), ("vtt", ms.write_vtt)):
        buf = io.StringIO()
        writer(segments, file=buf)
        expected = open(os.path.join(G, f"split_{expected_stem}_{length}.{ext}"), encoding="utf-8").read()
        assert buf.getvalue() == expected, (name, length, ext)


def test_make_subtitles_cli_file_and_folder(tmp_path):
    script = os.path.join(PKG, "make_subtitles.py")
    src = os.path.join(G, "in_yes_punctuations.mp3.words.json")
    inp = tmp_path / "yes_punctuations.mp3.words.json"
    inp.write_bytes(open(src, "rb").read())
    out_dir = tmp_path / "out"
    subprocess.run([sys.executable, script, str(inp), str(out_dir), "--max_length", "20"], check=True)
    for ext in ("srt", "vtt"):
        got = (out_dir / f"yes_punctuations.mp3.{ext}").read_text(encoding="utf-8")
        assert got == open(os.path.join(G, f"split_punctuations.mp3_20.{ext}"), encoding="utf-8").read()
    one = tmp_path / "single" / "x.srt"
    subprocess.run([sys.executable, script, str(inp), str(one), "--max_length", "6"], check=True)
    assert one.read_text(encoding="utf-8") == open(os.path.join(G, "split_punctuations.mp3_6.srt"), encoding="utf-8").read()
    subprocess.run([sys.executable, script, str(tmp_path), str(out_dir), "--max_length", "50", "--format", "vtt"], check=True)
    assert (out_dir / "yes_punctuations.mp3.vtt").read_text(encoding="utf-8") == \
        open(os.path.join(G, "split_punctuations.mp3_50.vtt"), encoding="utf-8").read()


def test_format_timestamp():
    ms = _load("make_subtitles")
    assert ms.format_timestamp(0) == "00:00.000"
    assert ms.format_timestamp(61.0049) == "01:01.005"
    assert ms.format_timestamp(3599.9996) == "01:00:00.000"
    assert ms.format_timestamp(2.76, always_include_hours=True, decimal_marker=",") == "00:00:02,760"
    with pytest.raises(AssertionError):
        ms.format_timestamp(-0.1)


@pytest.mark.parametrize("stem", ["punctuations.mp3", "bonjour.wav"])
@pytest.mark.parametrize("folder", ["punctuations_yes", "punctuations_no"])
def test_csv_tsv_writers_match_reference_outputs(folder, stem):
    wr = _load("writers")
    result = json.load(open(os.path.join(G, f"{folder}_{stem}.words.json"), encoding="utf-8"))

    def written(fn, items):
        buf = io.StringIO(newline="")
        fn(items, file=buf)
        return buf.getvalue()

    def expected(ext):
        with open(os.path.join(G, f"{folder}_{stem}.{ext}"), encoding="utf-8", newline="") as f:
            return f.read()

    assert written(wr.write_csv, result["segments"]) == expected("csv")
    assert written(wr.write_csv, wr.flatten(result["segments"], "words")) == expected("words.csv")
    assert written(wr.write_tsv, result["segments"]).replace("\r\n", "\n") == expected("tsv")
    assert written(wr.write_tsv, wr.flatten(result["segments"], "words")).replace("\r\n", "\n") == expected("words.tsv")


def test_filtered_keys_and_remove_keys():
    wr = _load("writers")
    res = {"text": " a", "language": "fr", "language_probs": {"fr": 0.123456}, "junk": 1,
           "segments": [{"id": 0, "start": 0.123456, "end": 1.0, "text": " a", "tokens": [1], "confidence": 0.98765,
                         "words": [{"text": "a", "start": 0.123456, "end": 1.0, "confidence": 0.5}]}]}
    out = wr.filtered_keys(res)
    assert out == {"text": " a", "language": "fr", "language_probs": {"fr": 0.123456},
                   "segments": [{"start": 0.12, "end": 1.0, "text": " a", "confidence": 0.99,
                                 "words": [{"text": "a", "start": 0.12, "end": 1.0, "confidence": 0.5}]}]}
    assert list(wr.remove_keys(res["segments"], "words"))[0].keys() == {"id", "start", "end", "text", "tokens", "confidence"}
    assert list(wr.flatten([[1, 2], [3]])) == [1, 2, 3]
