"""Host-side (string / integer) half of word alignment.

The reference does all of this inside `perform_word_alignment`
(/root/reference/whisper_timestamped/transcribe.py:1428-1793), interleaved with the numerics.
Here the numerics (attention post-processing + DTW, T.py:1510-1581) run batched on the GPU
(alignment.py / libwts), and this module does what surrounds them:

  prepare_alignment()  T.py:1466-1508, 1514-1535   frame window, token->word grouping, truncation
  words_from_jumps()   T.py:1654, 1711-1717, 1739-1793   word begin/end from the DTW jumps
  split_on_unicode() / split_on_spaces()   T.py:1815-1868
"""
import string
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

N_FRAMES = 3000
HOP_LENGTH = 160
SAMPLE_RATE = 16000
FRAMES_PER_WINDOW = N_FRAMES // 2            # 1500 encoder positions
TIME_PER_FRAME = HOP_LENGTH * 2 / SAMPLE_RATE   # 0.02 s

# T.py:1813 — note that membership tests below are SUBSTRING tests on this string, as in the reference
PUNCTUATION = "".join(c for c in string.punctuation if c not in ["-", "'"]) + "。，！？：”、…"


def round_timestamp(x):
    return round(x, 2)


def round_confidence(x):
    return round(x, 3)


def split_on_unicode(tokens, tokenizer, remove_punctuation_from_words=False, isolate_punctuations=False):
    """Group tokens into the smallest units that decode to valid UTF-8; punctuation marks are glued
    to the preceding unit unless that unit is a timestamp (T.py:1815-1842)."""
    units, unit_pieces, unit_ids = [], [], []
    pending = []
    for t in tokens:
        pending.append(t)
        text = tokenizer.decode_with_timestamps(
            [p for p in pending if p < tokenizer.eot or p >= tokenizer.timestamp_begin])
        if "�" in text:
            continue                         # incomplete UTF-8 sequence: wait for the next token
        pieces = [""] * (len(pending) - 1) + [text]
        stripped = text.strip()
        is_punct = (not isolate_punctuations) and bool(stripped) and stripped in PUNCTUATION
        after_special = bool(unit_ids) and unit_ids[-1][-1] >= tokenizer.timestamp_begin
        if is_punct and not after_special:
            if not units:
                units, unit_pieces = [""], [[]]
            if not remove_punctuation_from_words:
                units[-1] += text
            unit_pieces[-1].extend(pieces)
            unit_ids[-1].extend(pending)
        else:
            units.append(text)
            unit_pieces.append(pieces)
            unit_ids.append(pending)
        pending = []
    return units, unit_pieces, unit_ids


def split_on_spaces(tokens, tokenizer, remove_punctuation_from_words=False):
    """Merge unicode units into space-delimited words (T.py:1845-1868)."""
    units, unit_pieces, unit_ids = split_on_unicode(tokens, tokenizer, remove_punctuation_from_words)
    words, word_pieces, word_ids = [], [], []
    ts0 = tokenizer.timestamp_begin
    for i, (unit, pieces, ids) in enumerate(zip(units, unit_pieces, unit_ids)):
        special = ids[0] >= ts0
        prev_special = i > 0 and unit_ids[i - 1][0] >= ts0
        next_special = i + 1 < len(unit_ids) and unit_ids[i + 1][0] >= ts0
        prev_blank = i > 0 and not units[i - 1].strip()
        blank = not unit.strip()
        leading_space = unit.startswith(" ") and not blank
        punct = (not blank) and unit.strip() in PUNCTUATION
        starts_word = special or (not prev_blank and (prev_special or (leading_space and not punct)
                                                     or (blank and not next_special)))
        if starts_word:
            words.append(unit.strip())
            word_pieces.append(pieces)
            word_ids.append(ids)
        else:
            words[-1] = words[-1] + unit.strip()
            word_pieces[-1].extend(pieces)
            word_ids[-1].extend(ids)
    return words, word_pieces, word_ids


@dataclass
class AlignRequest:
    """One alignment problem, host view.  `row_sel` says which qk rows of the window feed the cost
    matrix: rows row0 .. row0+T-2 and then `last_row` (differs from row0+T-1 after truncation)."""
    tokens: List[int]
    T: int
    f0: int                      # start_token (frames)
    F: int                       # end_token - start_token
    row_offset_last: int         # index (relative to the segment's first row) of the row used last
    words: List[str]
    word_pieces: List[List[str]]
    word_ids: List[List[int]]
    punct_at_end: List[int]
    unfinished: bool
    refine_nframes: int
    truncated: bool = False
    warnings: List[str] = field(default_factory=list)

    @property
    def start_time(self):
        return self.f0 * TIME_PER_FRAME


def prepare_alignment(tokens, n_rows, tokenizer, use_space=True, refine_nframes=0,
                      remove_punctuation_from_words=False, include_punctuation_in_timing=False,
                      unfinished_decoding=False) -> Optional[AlignRequest]:
    """tokens: [start timestamp, text..., end timestamp | eot | fallback]; n_rows == len(tokens).

    Returns None for an empty segment (T.py:1478-1481).  Raises RuntimeError like the reference for
    a missing start timestamp or a null/negative duration (T.py:1471-1472, 1491-1492)."""
    tokens = list(tokens)
    assert len(tokens) > 1, f"Got unexpected sequence of tokens of length {len(tokens)}"
    assert n_rows == len(tokens), f"Attention weights have wrong shape: {n_rows} (expected {len(tokens)})."
    ts0 = tokenizer.timestamp_begin
    f0 = tokens[0] - ts0
    f1 = tokens[-1] - ts0
    if f0 < 0:
        raise RuntimeError(f"Missing start token in: {tokenizer.decode_with_timestamps(tokens)}")
    if f1 < 0:                                     # no end timestamp (model stuck / early <|endoftext|>)
        f1 = FRAMES_PER_WINDOW
    if f1 == f0 and refine_nframes == 0:
        return None
    f1 = min(FRAMES_PER_WINDOW, max(f1, f0 + len(tokens)))       # minimal duration (reference issue #67)
    if refine_nframes > 0:
        f0 = max(f0 - refine_nframes, 0)
        f1 = min(f1 + refine_nframes, FRAMES_PER_WINDOW)
    if f1 <= f0:
        raise RuntimeError("Got segment with null or negative duration "
                           f"{tokenizer.decode_with_timestamps(tokens)}: {f0} {f1}")
    F = f1 - f0
    T = len(tokens)
    truncated = False
    warnings = []
    last_rel = T - 1
    if T > F:
        # T.py:1516-1535: keep the first F-1 tokens and the final one; decoding counts as unfinished
        warnings.append(f"Too much text ({T} tokens) for the given number of frames ({F}) in: "
                        f"{tokenizer.decode_with_timestamps(tokens)}\nThe end of the text will be removed.")
        tokens = tokens[:F - 1] + [tokens[-1]]
        T = len(tokens)
        truncated = True
        unfinished_decoding = True
        # the recursive call of the reference recomputes the window from the truncated tokens
        inner = prepare_alignment(tokens, T, tokenizer, use_space=use_space, refine_nframes=refine_nframes,
                                  remove_punctuation_from_words=remove_punctuation_from_words,
                                  unfinished_decoding=True)
        if inner is None:
            return None
        inner.truncated = True
        inner.row_offset_last = last_rel
        inner.warnings = warnings + inner.warnings
        return inner

    split = split_on_spaces if use_space else split_on_unicode
    words, word_pieces, word_ids = split(tokens, tokenizer, remove_punctuation_from_words=remove_punctuation_from_words)
    # a trailing punctuation token does not extend its word's end time (T.py:1503-1508)
    punct_at_end = [0 if len(w) == 1 or w[-1] not in PUNCTUATION else 1 for w in word_pieces]
    if include_punctuation_in_timing:
        punct_at_end[:-2] = [0] * (len(punct_at_end) - 2)
    return AlignRequest(tokens=tokens, T=T, f0=f0, F=F, row_offset_last=last_rel, words=words,
                        word_pieces=word_pieces, word_ids=word_ids, punct_at_end=punct_at_end,
                        unfinished=unfinished_decoding, refine_nframes=refine_nframes, truncated=truncated,
                        warnings=warnings)


DISFLUENCY_MARK = "[*]"


def words_from_jumps(req: AlignRequest, jumps, lefts=None, tokenizer=None) -> List[dict]:
    """jumps: T+1 frame indices (first frame of every token row on the DTW path, then the last frame).
    Word begin = jump of its first token, end = jump after its last non-punctuation token
    (T.py:1711-1717); the leading/trailing timestamp pseudo-words are dropped (T.py:1739-1754).

    lefts (detect_disfluencies, T.py:1654-1736): per token -1, or the offset from its jump at which the peak analysis
    of its attention row says the token really starts; what lies before becomes a "[*]" pseudo-word (no tokens)."""
    jumps = np.asarray(jumps)
    assert len(jumps) == req.T + 1
    jumps_start = jumps
    disfluencies = {}
    if lefts is not None:
        jumps_start = jumps.copy()
        for i_token, (tok, begin, end) in enumerate(zip(req.tokens, jumps[:-1], jumps[1:])):
            if lefts[i_token] < 0:
                continue
            new_begin = int(lefts[i_token]) + int(begin)
            jumps_start[i_token] = new_begin
            if new_begin != begin:
                if tokenizer.decode_with_timestamps([tok]) not in PUNCTUATION:
                    disfluencies[i_token] = (int(begin), new_begin)
                else:
                    disfluencies[i_token + 1] = (int(begin), int(end))
    bounds = np.concatenate([[0], np.cumsum([len(p) for p in req.word_pieces])])
    begin = jumps_start[bounds[:-1]] * TIME_PER_FRAME
    end = jumps[bounds[1:] - np.asarray(req.punct_at_end, dtype=np.int64)] * TIME_PER_FRAME
    words, pieces, ids = list(req.words), list(req.word_pieces), list(req.word_ids)
    if lefts is not None:
        inserts = []
        first = 0
        for i_word, toks in enumerate(pieces[:-1]):
            if first in disfluencies and i_word > 0:
                b, e = disfluencies[first]
                inserts.append((i_word, b * TIME_PER_FRAME, e * TIME_PER_FRAME))
            first += len(toks)
        for (i_word, b, e) in reversed(inserts):        # from the end, so the indices stay valid
            words.insert(i_word, DISFLUENCY_MARK)
            pieces.insert(i_word, [])
            ids.insert(i_word, [])
            begin = np.insert(begin, i_word, b)
            end = np.insert(end, i_word, e)
    if not req.refine_nframes:
        begin[1] = begin[0]
        end[-2] = end[-1]
    if req.unfinished:
        sl = slice(1, None)
    else:
        sl = slice(1, -1)
    out = []
    for w, b, e, p, i in zip(words[sl], begin[sl], end[sl], pieces[sl], ids[sl]):
        if w.startswith("<|"):
            continue
        out.append(dict(text=w, start=round_timestamp(b + req.start_time), end=round_timestamp(e + req.start_time),
                        tokens=p, tokens_indices=i))
    return out


def ensure_increasing_positions(items, min_duration=0):
    """Force word (or segment) times to be monotone, splitting overlaps at their midpoint
    (T.py:2265-2295)."""
    while True:
        moved_back = False
        prev_end = 0
        for k, it in enumerate(items):
            if it["start"] < prev_end:
                assert k > 0
                mid = round_timestamp((prev_end + it["start"]) / 2)
                if mid < items[k - 1]["start"] + min_duration:
                    mid = prev_end
                else:
                    items[k - 1]["end"] = mid
                    moved_back = True
                it["start"] = mid
            if it["end"] <= it["start"] + min_duration:
                it["end"] = it["start"] + min_duration
            prev_end = it["end"]
        if not moved_back:
            break
    prev_end = 0
    for it in items:
        it["start"] = round_timestamp(it["start"])
        it["end"] = round_timestamp(it["end"])
        assert it["start"] >= prev_end, f"Got segment {it} coming before the previous finishes ({prev_end} > {it['start']})"
        assert it["end"] >= it["start"], f"Got segment {it} with end < start"
        prev_end = it["end"]
    return items


def remove_last_null_duration_words(transcription, words, recompute_text=False):
    """Drop zero-length words at the end of each 30-s chunk (T.py:2202-2262)."""
    chunk_of_segment = {}
    seek, chunk = None, -1
    for i, seg in enumerate(transcription["segments"]):
        if seg["seek"] != seek:
            chunk += 1
            seek = seg["seek"]
        chunk_of_segment[i] = chunk
    current, tail_empty = -1, False
    doomed = []
    for i in range(len(words) - 1, -1, -1):
        word = words[i]
        empty = word["start"] == word["end"]
        seg_idx = word["idx_segment"]
        grp = chunk_of_segment[seg_idx]
        if grp != current:
            tail_empty, current = empty, grp
        elif not empty:
            tail_empty = False
        if not tail_empty:
            continue
        doomed.append(i)
        full = "".join(word["tokens"])
        seg = transcription["segments"][seg_idx]
        text = seg["text"]
        if not text.endswith(full):
            if text.endswith(full[:-1]):
                full = full[:-1]
            elif text[:-1].endswith(full):
                text = text[:-1]
            else:
                raise RuntimeError(f"\"{text}\" not ending with \"{full}\"")
        text = text[:-len(full)]
        if i > 0 and words[i - 1]["idx_segment"] == seg_idx:
            seg["text"] = text
        else:
            transcription["segments"].pop(seg_idx)
            for j in range(i + 1, len(words)):
                words[j]["idx_segment"] -= 1
        recompute_text = True
    for i in doomed:
        words.pop(i)
    if recompute_text:
        transcription["text"] = "".join(s["text"] for s in transcription["segments"])
    return transcription, words
