"""Second pass of the two-pass ("naive") strategy — SURVEY.md §8 row A15, restating
/root/reference/whisper_timestamped/transcribe.py:1131-1327.

Pass 1 (plain decoding) has produced upstream-style segments.  For every segment (or, without trust in Whisper's
timestamps, for every 30-s window of segments) this pass
  1. picks the audio span to look at (T.py:1137-1174: the previous word end when it is close enough, margins around
     the segment otherwise) — which makes the pass SEQUENTIAL: span i+1 depends on the words of segment i;
  2. recomputes the log-mel of that span alone (T.py:1213-1215) and runs the decoder teacher-forced on
     `sot sequence + <|0.00|> + text tokens` (T.py:1244), capturing every cross-attention row of the alignment heads;
  3. aligns all tokens at once with the same attention post-processing + DTW as the efficient path (T.py:1251-1262);
  4. turns the teacher-forced token log-probabilities into word / segment confidences (T.py:1285-1320).

The device work (mel, encoder, teacher-forced decoder, prep, DTW, log-prob gather) is the engine's
(`log_mel(audio, pad_30s=False)`, `teacher_forced`, `align`); this file is host logic only.
"""
import logging

import numpy as np

from . import words as W
from .windows import HOP_LENGTH, N_FRAMES, SAMPLE_RATE

logger = logging.getLogger("whisper_timestamped")

AUDIO_SAMPLES_PER_TOKEN = HOP_LENGTH * 2
AUDIO_TIME_PER_TOKEN = AUDIO_SAMPLES_PER_TOKEN / SAMPLE_RATE
SEGMENT_DURATION = N_FRAMES * HOP_LENGTH / SAMPLE_RATE


def _span_trusting_timestamps(segment, next_segment, previous_end, audio_duration, refine_sec, min_word_duration):
    """Audio span (start, end) in seconds for one segment, or None to skip it (T.py:1139-1186)."""
    start, end = segment["start"], segment["end"]
    if end < start:                                   # Whisper mispredicted the segment end
        end = min(audio_duration, start + SEGMENT_DURATION)
    lo, hi = start - refine_sec, start + refine_sec
    if start >= audio_duration - min_word_duration or lo <= previous_end <= hi:
        start = previous_end                          # decoding restarts at <|0.00|>: begin where the last word ended
    else:
        start = lo
    if start > audio_duration - min_word_duration:
        logger.warning(f"Skipping segment outside of audio duration {audio_duration} (original: {segment['start']}-"
                       f"{segment['end']}, new: {start}-XXX)")
        return None
    end_lo, end_hi = end - refine_sec, end + refine_sec
    if next_segment is not None:
        # try to keep  end + min_word_duration <= next start + margin
        cap = next_segment["start"] + refine_sec - min_word_duration
        if cap >= end_lo:
            end_hi = min(cap, end_hi)
    end = min(audio_duration, end_hi)
    if end < start + min_word_duration:
        logger.warning(f"Got super short segment (original from whisper: {segment['start']}-{segment['end']}, "
                       f"new: {start, end})")
        end = min(audio_duration, start + min_word_duration)
        if end <= start:
            logger.warning("Skipping this short segment occuring too close to the end of the audio")
            return None
    return start, end


def _listed_words(req):
    """(pieces, token ids) of the words `W.words_from_jumps` will emit for this request, in order."""
    sl = slice(1, None) if req.unfinished else slice(1, -1)
    return [(p, i) for (w, p, i) in zip(req.words[sl], req.word_pieces[sl], req.word_ids[sl]) if not w.startswith("<|")]


def second_pass(eng, audio, whisper_segments, tokenizer, language, *, use_space, refine_nframes, trust_whisper_timestamps,
                remove_punctuation_from_words, compute_word_confidence, include_punctuation_in_confidence,
                min_word_duration=0.0, detect_disfluencies=False, verbose=False):
    """Adds `confidence` (and possibly corrected `tokens` / `text`) to the segments in place; returns the word list
    (each word carries `idx_segment`)."""
    tok = tokenizer
    ts0 = tok.timestamp_begin
    n_samples = int(audio.shape[-1])
    audio_duration = n_samples / SAMPLE_RATE
    refine_sec = refine_nframes * AUDIO_TIME_PER_TOKEN

    sot_sequence = tuple(tok.sot_sequence)
    if language and len(sot_sequence) == 3:
        sot_sequence = (sot_sequence[0], tok.to_language_token(language), sot_sequence[2])
    i_start0 = len(sot_sequence)

    words = []
    previous_end = 0
    window_tokens, token_to_segment = [], []          # only without trust in Whisper's timestamps
    for i_segment, segment in enumerate(whisper_segments):
        nxt = whisper_segments[i_segment + 1] if i_segment + 1 < len(whisper_segments) else None
        start = end = tokens = None
        if trust_whisper_timestamps:
            span = _span_trusting_timestamps(segment, nxt, previous_end, audio_duration, refine_sec, min_word_duration)
            if span is None:
                continue
            start, end = span
            tokens = list(segment["tokens"])
        else:
            # all segments of one 30-s window are aligned together (T.py:1188-1211)
            seek = segment["seek"]
            new_tokens = list(segment["tokens"])
            if not new_tokens:
                continue
            window_start = seek * HOP_LENGTH / SAMPLE_RATE
            if new_tokens[0] < ts0:
                new_tokens = [round((segment["start"] - window_start) * SAMPLE_RATE / AUDIO_SAMPLES_PER_TOKEN) + ts0] + new_tokens
            if new_tokens[-1] < ts0:
                new_tokens = new_tokens + [round((segment["end"] - window_start) * SAMPLE_RATE / AUDIO_SAMPLES_PER_TOKEN) + ts0]
            window_tokens.extend(new_tokens)
            token_to_segment.extend([i_segment] * len(new_tokens))
            if nxt is None or seek != nxt["seek"]:
                start = float(window_start)
                assert start < audio_duration, f"Got start {start} which is outside of audio duration {audio_duration}"
                end = min(start + SEGMENT_DURATION, audio_duration)
                tokens = window_tokens
        if tokens is None or not len(tokens):
            continue

        start_sample = min(round(start * SAMPLE_RATE), n_samples)
        end_sample = min(round(end * SAMPLE_RATE), n_samples)
        sub = audio[start_sample:end_sample]
        if sub.shape[-1] <= 200:                       # audio_minimum_padding (T.py:1349-1352)
            sub = _pad_to(sub, 201)
        mel = eng.log_mel(sub, pad_30s=False)
        content_frames = min(N_FRAMES, eng.mel_frames(mel))
        max_duration = content_frames // 2 if content_frames < N_FRAMES else None

        check = []                                    # what the segment's tokens should have been
        if tokens[0] >= ts0:
            check.append(tokens[0])
        while tokens[0] >= ts0:
            tokens = tokens[1:]
            assert len(tokens), "Got transcription with only timestamps!"
        last_token_check = None
        while tokens[-1] >= ts0:
            last_token_check = tokens[-1]
            tokens = tokens[:-1]

        tokens_in = [*sot_sequence, ts0] + list(tokens)
        i_start = i_start0
        end_token = ts0 + round(min(N_FRAMES * HOP_LENGTH, end_sample - start_sample) // AUDIO_SAMPLES_PER_TOKEN)
        tokens_out = tokens_in[i_start:] + [end_token]

        req = W.prepare_alignment(tokens_out, len(tokens_out), tok, use_space=use_space, refine_nframes=refine_nframes,
                                  remove_punctuation_from_words=remove_punctuation_from_words)
        ws = []
        lp = None
        if req is not None:
            for msg in req.warnings:
                logger.warning(msg)
            if max_duration and req.f0 >= max_duration:
                logger.warning("Got start time outside of audio boundary")
            listed = _listed_words(req)
            # (decoder position, token) pairs whose teacher-forced log-probability feeds the confidences
            pairs, per_word = [], []
            step = i_start
            for pieces, ids in listed:
                ids_kept, pieces_kept = list(ids), list(pieces)
                if include_punctuation_in_confidence:     # sic: the flag name is inverted in the reference (T.py:1290-1293)
                    while len(pieces_kept) > 1 and len(pieces_kept[-1]) and pieces_kept[-1][-1] in W.PUNCTUATION:
                        pieces_kept, ids_kept = pieces_kept[:-1], ids_kept[:-1]
                sel = list(zip(range(step, step + len(ids_kept)), ids_kept))
                per_word.append((len(pairs), len(sel)))
                pairs.extend(sel)
                step += len(pieces)
            window, lp = eng.teacher_forced(mel, tokens_in, i_start, pairs if compute_word_confidence else [])
            item = dict(window=window, row0=0, last_row=req.row_offset_last, T=req.T, f0=req.f0, F=req.F,
                        max_dur=max_duration or 0)
            if detect_disfluencies:
                jl, ll = eng.align([item], disfluencies=True)
                ws = W.words_from_jumps(req, jl[0], ll[0], tokenizer=tok)
            else:
                ws = W.words_from_jumps(req, eng.align([item])[0])
            assert sum(1 for w_ in ws if w_["text"] != W.DISFLUENCY_MARK or w_["tokens"]) == len(listed)

        segment_logprobs = []
        i_token = 1
        k = -1                                        # index among the words that carry tokens ("[*]" marks have none)
        for word in ws:
            is_mark = word["text"] == W.DISFLUENCY_MARK and not word["tokens"]
            if not is_mark:
                k += 1
            word["start"] = round(word["start"] + start, 2)
            word["end"] = round(word["end"] + start, 2)
            if trust_whisper_timestamps:
                word["idx_segment"] = i_segment
            else:
                assert i_token < len(tokens_out)
                assert not len(word["tokens_indices"]) or word["tokens_indices"][0] == tokens_out[i_token]
                word["idx_segment"] = token_to_segment[i_token]
                i_token += len(word["tokens"])
                while i_token < len(tokens_out) and tokens_out[i_token] >= ts0:
                    i_token += 1
            check.extend(word["tokens_indices"])
            if compute_word_confidence:
                a, n = (0, 0) if is_mark else per_word[k]
                wl = np.asarray(lp[a:a + n], dtype=np.float32)
                if n:
                    segment_logprobs.append(wl)
                    word["confidence"] = W.round_confidence(float(np.exp(wl.mean(dtype=np.float32))))
                else:
                    word["confidence"] = W.round_confidence(0)
            words.append(word)
            if verbose:                                   # T.py:1304-1305
                from .transcribe import print_timestamped
                print_timestamped(word)

        if last_token_check is not None:
            check.append(last_token_check)
        if trust_whisper_timestamps and check != segment["tokens"]:
            assert len(check) < len(segment["tokens"]), \
                f"First should be longer by one token: '{tok.decode_with_timestamps(check)}' should include " \
                f"'{tok.decode_with_timestamps(segment['tokens'])}'"
            assert check[:-1] == segment["tokens"][:len(check) - 1], \
                f"Got inconsistent tokens: {tok.decode_with_timestamps(check)} != {tok.decode_with_timestamps(segment['tokens'])}"
            segment["tokens"] = check
            segment["text"] = tok.decode(segment["tokens"])
        if segment_logprobs:
            cat = np.concatenate(segment_logprobs)
            segment["confidence"] = W.round_confidence(float(np.exp(cat.mean(dtype=np.float32))))
        if ws:
            previous_end = ws[-1]["end"]
        if not trust_whisper_timestamps:
            window_tokens, token_to_segment = [], []
    return words


def _pad_to(x, n):
    import torch
    out = torch.zeros(n, dtype=x.dtype, device=x.device)
    out[: x.shape[-1]] = x
    return out
