"""Tokenizer facts the reference relies on (upstream `whisper.tokenizer`, reached at
/root/reference/whisper_timestamped/transcribe.py:1417-1421) that can be pinned offline.

* special-token ids: the reference's own tests name them (tests/test_transcribe.py:733-734, 869, 893 and
  SURVEY.md Appendix A);
* the `suppress_tokens="-1"` set on the REAL-vocabulary (tiktoken) path: upstream's symbol list includes the CJK
  corner brackets, which are single tokens in multilingual.tiktoken.  The real rank file is not in this image, so the
  rule (a symbol is suppressed iff it encodes to ONE token, with and without a leading space) is exercised on a
  miniature rank file in which exactly those symbols are merged tokens.
"""
import base64
import os

import pytest

from whisper_timestamped import tokenizer as T


def test_special_token_ids_match_the_reference_tests():
    ml = T.get_tokenizer(True, num_languages=99, language="en", task="transcribe")
    assert (ml.eot, ml.sot, ml.timestamp_begin) == (50257, 50258, 50364)       # <|0.00|> multilingual
    assert ml.timestamp_begin + 350 == 50714 and ml.timestamp_begin + 1500 == 51864   # <|7.00|>, <|30.00|>
    assert ml.sot_sequence == (50258, 50259, 50359)
    en = T.get_tokenizer(False, num_languages=99)
    assert (en.eot, en.sot, en.timestamp_begin) == (50256, 50257, 50363)       # <|0.00|> English-only
    assert en.timestamp_begin + 1450 == 51813                                   # <|29.00|>
    assert en.sot_sequence == (50257,)
    v3 = T.get_tokenizer(True, num_languages=100, language="en", task="transcribe")
    assert v3.timestamp_begin == 50365 and v3.n_vocab == 51866


def _write_ranks(path, n_text, merged):
    """A syntactically valid .tiktoken file: 256 byte tokens, the given merged strings (with every prefix needed to
    reach them by pair merges), unreachable filler up to n_text."""
    toks = [bytes([i]) for i in range(256)]
    seen = set(toks)
    for m in merged:
        b = m.encode("utf-8")
        for k in range(2, len(b) + 1):
            if b[:k] not in seen:
                seen.add(b[:k])
                toks.append(b[:k])
    i = 0
    while len(toks) < n_text:
        f = b"\xf5\xf6" + i.to_bytes(4, "big")          # bytes that never occur in UTF-8 text
        toks.append(f)
        i += 1
    with open(path, "w") as f:
        for rank, t in enumerate(toks):
            f.write(base64.b64encode(t).decode() + f" {rank}\n")
    return {t: r for r, t in enumerate(toks)}


def test_non_speech_tokens_on_the_tiktoken_path(tmp_path, monkeypatch):
    pytest.importorskip("tiktoken")
    single = ["「", "」", "『", "』", " 「", "<<", " >>", "♪"]
    ranks = _write_ranks(os.path.join(tmp_path, "multilingual.tiktoken"), 50257, single)
    monkeypatch.setenv("WTS_WHISPER_ASSETS", str(tmp_path))
    tok = T.Tokenizer(True, 99, "ja", "transcribe")
    assert tok.vocab.kind == "tiktoken"
    ns = set(tok.non_speech_tokens)
    for s in single:
        assert ranks[s.encode("utf-8")] in ns, s
    for c in '"#()*+/:;<=>@[\\]^_`{|}~':                  # single bytes are single tokens
        assert ranks[c.encode()] in ns
    assert ranks[b" "] in ns                               # first token of " -" and " '" (no merge in this miniature)
    # symbols that stay multi-token are NOT suppressed through their first byte (only the musical notes are)
    assert ranks[b"\xe3"] not in ns                        # first byte of the CJK brackets
    assert ranks["♫".encode()[:2]] in ns              # ♫ is two tokens here (the ♪ prefix merge + a byte): upstream adds
                                                       # the FIRST token of a musical note whatever the length
    # the byte-level synthetic vocabulary keeps the ASCII set
    monkeypatch.delenv("WTS_WHISPER_ASSETS")
    syn = T.Tokenizer(True, 99, "ja", "transcribe")
    assert syn.vocab.kind == "synthetic-v1" and len(syn.non_speech_tokens) == 23
