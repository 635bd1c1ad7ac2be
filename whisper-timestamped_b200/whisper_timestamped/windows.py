"""Host-side decode bookkeeping: what upstream `whisper.transcribe` does around each 30-s window
(seek loop, prompt carry-over, timestamp-token segment slicing — SURVEY.md Appendix A; driven by the
reference at /root/reference/whisper_timestamped/transcribe.py:904) and what the reference's
forward-hook state machine derives from the same token stream (T.py:419-781, 801-881), restated
OFFLINE: the GPU engine decodes a whole batch of windows first and hands back one `WindowRecord`
per window; no hook runs inside the decode loop.

Row convention: qk row r and log-prob row r of a window belong to the decoder position that
PREDICTED sampled token r (row 0 = last prompt token as input).  A segment made of sampled tokens
i..j therefore owns rows i..j (T.py:488, 525: tokens = segment_tokens[1:], rows = weights[:-1]).
"""
import zlib
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

N_FRAMES = 3000
HOP_LENGTH = 160
SAMPLE_RATE = 16000
INPUT_STRIDE = 2
TIME_PRECISION = 0.02


def compression_ratio(text: str) -> float:
    b = text.encode("utf-8")
    return len(b) / len(zlib.compress(b))


@dataclass
class DecodeSetup:
    """Static decoding configuration (upstream `DecodingTask.__init__`, reached by the reference through
    get_logit_filters, T.py:1371-1393)."""
    tokenizer: object
    n_ctx: int
    sample_len: int
    suppress_tokens: tuple
    blank_tokens: tuple            # encode(" ") + [eot], suppressed at the first sampled position
    max_initial_timestamp_index: Optional[int]
    temperature: float = 0.0
    # upstream decoding strategies (DecodingOptions): beam search at temperature 0, best-of-n sampling above
    beam_size: Optional[int] = None
    patience: Optional[float] = None
    best_of: Optional[int] = None
    length_penalty: Optional[float] = None

    @property
    def n_group(self):
        """Hypotheses decoded per window (upstream DecodingTask.n_group)."""
        return self.beam_size or self.best_of or 1

    def at_temperature(self, t):
        """Upstream decode_with_fallback: beam search only at temperature 0, best_of only above."""
        from dataclasses import replace
        if t > 0:
            return replace(self, temperature=float(t), beam_size=None, patience=None)
        return replace(self, temperature=float(t), best_of=None)

    def initial_tokens(self, prompt_tokens):
        tok = self.tokenizer
        seq = list(tok.sot_sequence)
        if prompt_tokens:
            seq = [tok.sot_prev] + list(prompt_tokens)[-(self.n_ctx // 2 - 1):] + seq
        return seq


def make_decode_setup(tokenizer, n_text_ctx, sample_len=None, suppress_tokens="-1", temperature=0.0,
                      max_initial_timestamp=1.0, suppress_blank=True, beam_size=None, patience=None, best_of=None,
                      length_penalty=None) -> DecodeSetup:
    # upstream DecodingTask._verify_options
    if patience is not None and beam_size is None:
        raise ValueError("patience requires beam_size to be given")
    if length_penalty is not None and not (0 <= length_penalty <= 1):
        raise ValueError("length_penalty (alpha) should be a value between 0 and 1")
    if isinstance(suppress_tokens, str):
        suppress = [int(t) for t in suppress_tokens.split(",")]
    elif suppress_tokens is None:
        suppress = []
    else:
        suppress = list(suppress_tokens)
    if -1 in suppress:
        suppress = [t for t in suppress if t >= 0]
        suppress.extend(tokenizer.non_speech_tokens)
    suppress.extend([tokenizer.transcribe, tokenizer.translate, tokenizer.sot, tokenizer.sot_prev, tokenizer.sot_lm])
    if tokenizer.no_speech is not None:
        suppress.append(tokenizer.no_speech)
    blank = tuple(tokenizer.encode(" ") + [tokenizer.eot]) if suppress_blank else ()
    mit = round(max_initial_timestamp / TIME_PRECISION) if max_initial_timestamp else None
    return DecodeSetup(tokenizer=tokenizer, n_ctx=n_text_ctx, sample_len=sample_len or n_text_ctx // 2,
                       suppress_tokens=tuple(sorted(set(suppress))), blank_tokens=blank,
                       max_initial_timestamp_index=mit, temperature=temperature, beam_size=beam_size, patience=patience,
                       best_of=best_of, length_penalty=length_penalty)


@dataclass
class WindowRecord:
    """Everything the host needs about one decoded 30-s window."""
    seek: int                       # mel frame where the window starts
    segment_size: int               # content frames in the window (< 3000 => zero-padded mel)
    prompt: List[int]               # initial tokens fed (prompt + sot sequence)
    tokens: List[int]               # sampled tokens, <|endoftext|> excluded
    logprobs: np.ndarray            # row r: filtered log-softmax of the token chosen at row r
    ended_by_eot: bool
    no_speech_prob: float
    qk_window: int                  # index of this window in the engine's qk buffer
    temperature: float = 0.0
    language: Optional[str] = None
    last_row_logprobs: object = None   # callable(token) -> logprob at the last row (rare fallback path)
    sum_logprob: Optional[float] = None   # beam search / sampling: cumulative log-prob of the selected hypothesis
    mel_from_language_detection: bool = False   # see max_duration

    @property
    def n_rows(self):
        return len(self.logprobs) if self.logprobs is not None else 0

    @property
    def avg_logprob(self):
        # upstream: sum of the log-probs of every sampled token (EOT included) / (len(tokens) + 1)
        if self.sum_logprob is not None:
            return float(self.sum_logprob) / (len(self.tokens) + 1)
        return float(np.sum(self.logprobs.astype(np.float32), dtype=np.float32)) / (len(self.tokens) + 1)

    @property
    def max_duration(self):
        """find_start_padding(mfcc) // 2 (T.py:1556-1558): None unless the window's mel is zero-padded.

        One more None: when the language is auto-detected, the reference's conv1 hook first fires on upstream's
        detect_language pass (T.py:795-799: `mfcc` is only set while it is None) and `mfcc` is not replaced before
        the NEXT window starts (T.py:708).  The first window of the file is therefore aligned against the
        detection mel — the first 30 s of the log-mel of the zero-PADDED AUDIO, whose tail columns are a negative
        constant, not zero — so find_start_padding() finds no padding and no mask is applied."""
        if self.mel_from_language_detection:
            return None
        return self.segment_size // 2 if self.segment_size < N_FRAMES else None


def slice_window_segments(rec: WindowRecord, tokenizer, text_of=None):
    """Upstream's timestamp-token slicing of one window.  Returns (segments, seek_advance, skipped).
    `segments` are upstream-style dicts (no `id` yet)."""
    tok = tokenizer
    tokens = list(rec.tokens)
    time_offset = float(rec.seek * HOP_LENGTH / SAMPLE_RATE)
    segment_duration = rec.segment_size * HOP_LENGTH / SAMPLE_RATE
    is_ts = [t >= tok.timestamp_begin for t in tokens]
    text = tok.decode(tokens).strip()
    meta = dict(temperature=rec.temperature, avg_logprob=rec.avg_logprob,
                compression_ratio=compression_ratio(text), no_speech_prob=rec.no_speech_prob)

    def new_segment(start, end, toks):
        text_tokens = [t for t in toks if t < tok.eot]
        return {"seek": rec.seek, "start": start, "end": end, "text": tok.decode(text_tokens), "tokens": list(toks), **meta}

    segments = []
    single_timestamp_ending = is_ts[-2:] == [False, True]
    consecutive = [i + 1 for i in range(len(tokens) - 1) if is_ts[i] and is_ts[i + 1]]
    if consecutive:
        slices = list(consecutive)
        if single_timestamp_ending:
            slices.append(len(tokens))
        last = 0
        for cur in slices:
            sl = tokens[last:cur]
            s_pos = sl[0] - tok.timestamp_begin
            e_pos = sl[-1] - tok.timestamp_begin
            segments.append(new_segment(time_offset + s_pos * TIME_PRECISION, time_offset + e_pos * TIME_PRECISION, sl))
            last = cur
        if single_timestamp_ending:
            advance = rec.segment_size
        else:
            advance = (tokens[last - 1] - tok.timestamp_begin) * INPUT_STRIDE
    else:
        duration = segment_duration
        ts_tokens = [t for t in tokens if t >= tok.timestamp_begin]
        if ts_tokens and ts_tokens[-1] != tok.timestamp_begin:
            duration = (ts_tokens[-1] - tok.timestamp_begin) * TIME_PRECISION
        segments.append(new_segment(time_offset, time_offset + duration, tokens))
        advance = rec.segment_size
    for seg in segments:          # instantaneous or text-less segments are cleared
        if seg["start"] == seg["end"] or seg["text"].strip() == "":
            seg["text"] = ""
            seg["tokens"] = []
            seg["words"] = []
    return segments, advance


@dataclass
class AlignedSegmentPlan:
    """A segment the reference would hand to perform_word_alignment (T.py:482-566)."""
    tokens: List[int]          # [start ts, text..., end ts | eot | fallback]
    row0: int                  # first qk / log-prob row
    n_rows: int
    unfinished: bool
    last_token_reliable: bool = True
    appended_token: Optional[int] = None     # eot / fallback token appended by the flush logic


def plan_window_alignment(rec: WindowRecord, setup: DecodeSetup, next_prompt: Optional[List[int]],
                          yields_words=None):
    """Offline equivalent of must_flush_segment / align_last_segment / reset (T.py:427-566) with
    trust_whisper_timestamps=True.  Returns (plans, chunk_info).

    `yields_words(plan) -> bool` tells whether perform_word_alignment would return at least one word for
    the flushed tokens (it depends on the tokens only).  When it would not, the reference does not add the
    segment and resets its token list to EMPTY instead of keeping the closing timestamp (T.py:559-564,
    427-442), which is replayed here — including the RuntimeError("Missing start token") the reference then
    raises at the next flush of the same window."""
    tok = setup.tokenizer
    ts0 = tok.timestamp_begin
    S = list(rec.tokens)
    n_rows = rec.n_rows
    n_fed = n_rows - 1                    # sampled tokens that were fed back as decoder input
    P = len(rec.prompt)

    def limit_reached(n_nosot):
        n = n_nosot + 1
        m = n + P
        return n + 1 >= setup.sample_len or m > setup.n_ctx

    plans: List[AlignedSegmentPlan] = []
    cur = [rec.prompt[-1]]                # T.py:836: the last prompt token opens the first list
    rows = [0]
    saw_consecutive = False
    for k in range(n_fed):
        t = S[k]
        if t >= ts0 and cur and cur[-1] >= ts0:
            saw_consecutive = True
            seg_tokens = cur[1:]
            unfinished = limit_reached(k)
            if unfinished:
                # the model is stuck right at a segment boundary: recover like the final flush does
                plan = _flush(cur, rows, unfinished, rec, setup, None, mid_chunk=True)
            else:
                plan = AlignedSegmentPlan(tokens=seg_tokens, row0=rows[0], n_rows=len(rows) - 1, unfinished=False)
            if yields_words is None or yields_words(plan):
                plans.append(plan)
                cur, rows = [cur[-1]], [rows[-1]]
            else:
                cur, rows = [], []
        cur.append(t)
        rows.append(k + 1)

    reached = limit_reached(n_fed)
    # T.py:878-881: greedy guess of the token that would follow, only tracked once the limit is reached
    last_chunk_token = (S[n_fed] if n_fed < len(S) else tok.eot) if reached else None
    must_flush = len(cur) > 1 and not saw_consecutive
    if not must_flush:
        if last_chunk_token is None:
            must_flush = len(cur) > 2 and cur[-1] >= ts0
        else:
            must_flush = last_chunk_token >= ts0
    final_plan = None
    if must_flush:
        final_plan = _flush(cur, rows, reached, rec, setup, next_prompt, last_chunk_token=last_chunk_token)
        if yields_words is None or yields_words(final_plan):
            plans.append(final_plan)
        else:
            final_plan = None
    for plan in plans:
        # T.py:529-538: the reference re-estimates an end timestamp that is not after the start one from the full last
        # log-prob row.  With greedy decoding under ApplyTimestampRules a closing timestamp is strictly greater than the
        # opening one, so this cannot happen on the built paths; if it ever does (new sampling rules, a fallback token
        # taken from the next prompt), fail loudly instead of silently diverging from the reference.
        if len(plan.tokens) >= 2 and plan.tokens[0] >= ts0 and plan.tokens[-1] >= ts0 and plan.tokens[-1] <= plan.tokens[0]:
            raise NotImplementedError(
                f"segment ends with timestamp {plan.tokens[-1]} <= its start {plan.tokens[0]}: the reference re-estimates "
                "the end from the last log-prob row (T.py:529-538), which this drop-in does not keep per row")
    info = dict(n_fed=n_fed, reached=reached, last_chunk_token=last_chunk_token,
                final_unfinished=bool(final_plan and final_plan.unfinished))
    return plans, info


def _flush(cur, rows, unfinished, rec, setup, next_prompt, last_chunk_token=None, mid_chunk=False):
    tok = setup.tokenizer
    ts0 = tok.timestamp_begin
    tokens = cur[1:]
    reliable = True
    appended = None
    if unfinished:
        if next_prompt is not None and next_prompt[0] == tok.sot_prev:
            idx = next_prompt.index(tok.sot)
            assert idx > 0
            appended = next_prompt[idx - 1]
        else:
            if last_chunk_token is None:
                # argmax of the last log-prob row == the token greedy decoding sampled there
                last_chunk_token = rec.tokens[rec.n_rows - 1] if rec.n_rows - 1 < len(rec.tokens) else tok.eot
            appended = last_chunk_token
            reliable = setup.temperature == 0
        tokens = tokens + [appended]
        use_rows = len(rows)
    elif tokens and tokens[-1] < ts0:
        appended = tok.eot                          # <|endoftext|> came without a closing timestamp
        tokens = tokens + [appended]
        use_rows = len(rows)
    else:
        use_rows = len(rows) - 1
    return AlignedSegmentPlan(tokens=tokens, row0=rows[0], n_rows=use_rows, unfinished=unfinished,
                              last_token_reliable=reliable, appended_token=appended)


def needs_fallback(rec: WindowRecord, tokenizer, compression_ratio_threshold, logprob_threshold, no_speech_threshold) -> bool:
    """Upstream decode_with_fallback's verdict on one decoded window (driven by the reference through T.py:111-113)."""
    need = False
    if compression_ratio_threshold is not None:
        text = tokenizer.decode(list(rec.tokens)).strip()
        if compression_ratio(text) > compression_ratio_threshold:
            need = True                                   # too repetitive
    if logprob_threshold is not None and rec.avg_logprob < logprob_threshold:
        need = True                                       # average log probability is too low
    if (no_speech_threshold is not None and rec.no_speech_prob > no_speech_threshold
            and logprob_threshold is not None and rec.avg_logprob < logprob_threshold):
        need = False                                      # silence
    return need
