"""Public API of the drop-in: transcribe_timestamped() / transcribe().

Same signature, option checks and return value as the reference
(/root/reference/whisper_timestamped/transcribe.py:79-357), but organised for a batched GPU engine:

  1. every audio "stream" (the whole file, or independent fixed-length cuts when `chunks=` is given)
     advances through upstream's seek loop; each round the next window of EVERY active stream is
     decoded in ONE engine batch (encoder + greedy decoder on the GPU, cross-attention rows of the
     alignment heads written straight into the alignment buffer — no forward hooks, no per-token
     device->host copies: replaces T.py:783-793, 849-881);
  2. when all windows are decoded, the hook state machine of the reference is replayed offline
     (windows.py) to find the segments, and ALL segments are aligned in one batch on the GPU
     (alignment.py: attention post-processing + DTW);
  3. words, confidences and post-processing follow T.py:912-1002 and 313-357.
"""
import logging
import sys
from typing import List, Optional

import numpy as np

from . import vad as V
from . import words as W
from .tokenizer import LANGUAGES, TO_LANGUAGE_CODE, get_tokenizer
from .writers import filtered_keys, flatten, remove_keys, write_csv, write_tsv  # noqa: F401  (T.py:2298-2323, 3183-3199)
from .windows import (HOP_LENGTH, N_FRAMES, SAMPLE_RATE, WindowRecord, make_decode_setup, plan_window_alignment,
                      slice_window_segments)

logger = logging.getLogger("whisper_timestamped")

AUDIO_TIME_PER_TOKEN = 0.02
USE_EFFICIENT_BY_DEFAULT = True
TRUST_WHISPER_TIMESTAMP_BY_DEFAULT = True
DISFLUENCY_MARK = "[*]"


def should_use_space(language):
    return norm_language(language) not in ["zh", "ja", "th", "lo", "my", "yue"]


def norm_language(language):
    if language is None:
        return "en"
    return TO_LANGUAGE_CODE.get(language.lower(), language)


class _Stream:
    """One independently transcribed piece of audio: upstream's seek loop state."""

    def __init__(self, index, mel_handle, content_frames, time_shift, initial_prompt_tokens):
        self.index = index
        self.mel = mel_handle
        self.content_frames = content_frames
        self.time_shift = time_shift               # seconds added to every time of this stream
        self.seek = 0
        self.all_tokens = list(initial_prompt_tokens)
        self.n_initial_prompt = len(initial_prompt_tokens)
        self.prompt_reset_since = 0
        self.segments = []                          # upstream-style segment dicts
        self.records: List[WindowRecord] = []       # every decoded window, in order
        self.kept = []                              # per record: was it kept (not skipped as silence)

    @property
    def active(self):
        return self.seek < self.content_frames

    def next_job(self, setup):
        size = min(N_FRAMES, self.content_frames - self.seek)
        prompt = setup.initial_tokens(self.all_tokens[self.prompt_reset_since:])
        return dict(stream=self.index, mel=self.mel, seek=self.seek, segment_size=size, prompt=prompt)

    def consume(self, rec: WindowRecord, tokenizer, no_speech_threshold, logprob_threshold,
                condition_on_previous_text):
        self.records.append(rec)
        if no_speech_threshold is not None:
            skip = rec.no_speech_prob > no_speech_threshold
            if logprob_threshold is not None and rec.avg_logprob > logprob_threshold:
                skip = False
            if skip:
                self.kept.append(False)
                self.seek += rec.segment_size
                return
        self.kept.append(True)
        segs, advance = slice_window_segments(rec, tokenizer)
        self.seek += advance
        for s in segs:
            s["id"] = len(self.segments)
            self.segments.append(s)
            self.all_tokens.extend(s["tokens"])
        if not condition_on_previous_text or rec.temperature > 0.5:
            self.prompt_reset_since = len(self.all_tokens)


def decode_with_fallback(eng, jobs, setup, temperatures, tokenizer, compression_ratio_threshold, logprob_threshold,
                         no_speech_threshold):
    """Upstream `decode_with_fallback` (reached by the reference through model.transcribe, T.py:904 / 1068, options
    T.py:111-113) for a BATCH of windows: every window is decoded at the first temperature (beam search or greedy at
    0, best-of-n sampling above); the windows whose result is too repetitive or too unlikely — and not silence — are
    decoded again, together, at the next temperature; the last attempt stands."""
    from .windows import needs_fallback
    final = [None] * len(jobs)
    pending = list(range(len(jobs)))
    for k, t in enumerate(temperatures):
        recs = eng.decode_windows([jobs[i] for i in pending], setup.at_temperature(t))
        again = []
        for i, rec in zip(pending, recs):
            final[i] = rec
            if k + 1 < len(temperatures) and needs_fallback(rec, tokenizer, compression_ratio_threshold, logprob_threshold,
                                                            no_speech_threshold):
                again.append(i)
        pending = again
        if not pending:
            break
    return final


def transcribe_timestamped(
    model,
    audio,
    language=None,
    task="transcribe",
    remove_punctuation_from_words=False,
    compute_word_confidence=True,
    include_punctuation_in_confidence=False,
    refine_whisper_precision=0.5,
    min_word_duration=0.02,
    plot_word_alignment=False,
    word_alignment_most_top_layers=None,
    remove_empty_words=False,
    use_backend_timestamps=False,
    seed=1234,
    vad=False,
    detect_disfluencies=False,
    trust_whisper_timestamps=TRUST_WHISPER_TIMESTAMP_BY_DEFAULT,
    naive_approach=False,
    temperature=0.0 if USE_EFFICIENT_BY_DEFAULT else (0.0, 0.2, 0.4, 0.6, 0.8, 1.0),
    best_of=None,
    beam_size=None,
    patience=None,
    length_penalty=None,
    compression_ratio_threshold=2.4,
    logprob_threshold=-1.0,
    no_speech_threshold=0.6,
    fp16=None,
    condition_on_previous_text=True,
    initial_prompt=None,
    suppress_tokens="-1",
    sample_len=None,
    verbose=False,
    *,
    chunks=None,
    engine=None,
    continuous_batching=False,
):
    """Drop-in for whisper_timestamped.transcribe (reference T.py:79-357).

    Extra keyword-only arguments (not in the reference):
      chunks: None = upstream semantics (one sequential stream over the whole file).  A number of
              seconds = cut the audio at fixed boundaries and transcribe every cut as an independent
              file (condition_on_previous_text is forced off across cuts); all cuts are decoded in the
              same GPU batches.  This is the data-parallel mode BASELINE.json's north_star names.
      engine: inject a decode/alignment engine (tests); default = the model's CUDA engine.
      continuous_batching: let the engine admit a stream's next window into the running decode batch as soon as its
              previous window finishes, instead of decoding in rounds that wait for their slowest window.  Same
              results.  Off by default: on the bench workload (20 % of the windows run to the 224-token limit, so do
              follow-up windows) it measured 1.6 % slower than rounds; it pays when windows end early and unevenly.
    """
    # ---- option checks, as T.py:223-261
    assert refine_whisper_precision >= 0 and refine_whisper_precision / AUDIO_TIME_PER_TOKEN == round(
        refine_whisper_precision / AUDIO_TIME_PER_TOKEN), \
        f"refine_whisper_precision must be a positive multiple of {AUDIO_TIME_PER_TOKEN}"
    refine_nframes = round(refine_whisper_precision / AUDIO_TIME_PER_TOKEN)
    assert min_word_duration >= 0, "min_word_duration must be a positive number"
    assert word_alignment_most_top_layers is None or word_alignment_most_top_layers > 0, \
        "word_alignment_most_top_layers must be a strictly positive number"
    if isinstance(temperature, (list, tuple)) and len(temperature) == 1:
        temperature = temperature[0]
    if isinstance(temperature, (list, tuple)):
        naive_approach = True
    elif temperature > 0 and best_of is not None and best_of > 1:
        naive_approach = True
    if beam_size is not None:
        naive_approach = True
    if use_backend_timestamps:
        naive_approach = True
    if isinstance(model, str):
        from .model import load_model
        model = load_model(model)
    if use_backend_timestamps:
        raise NotImplementedError("use_backend_timestamps (upstream whisper.timing / HF token timestamps) is not built in "
                                  "this B200 drop-in")
    if not naive_approach and temperature != 0:
        raise NotImplementedError(
            "a scalar temperature > 0 inside the one-pass strategy (sampling under the attention hooks) is not built; "
            "pass naive_approach=True, best_of > 1 or a temperature tuple (two-pass strategy, SURVEY.md §8 rows A14/A15)")
    if not trust_whisper_timestamps and not naive_approach:
        raise NotImplementedError("trust_whisper_timestamps=False is only built for the two-pass strategy (naive_approach=True)")
    if plot_word_alignment:
        raise NotImplementedError("plot_word_alignment is out of scope of the hot path")
    vad = V.check_vad_method(vad)          # explicit (start, end) lists only; detector names raise NotImplementedError
    if seed is not None:                   # T.py:223-225: sampling (temperature > 0) draws from torch's global generator
        import torch
        torch.manual_seed(seed)
    if word_alignment_most_top_layers is not None:
        raise NotImplementedError("word_alignment_most_top_layers: only the alignment-head tables are built")

    eng = engine if engine is not None else model.engine()
    if hasattr(eng, "release"):
        eng.release()                      # alignment buffers of a previous call (they are per call, 1.7 GB per 128 windows)
    dims = model.dims
    is_multilingual = model.is_multilingual
    num_languages = model.num_languages

    # ---- audio -> log-mel on the device, per stream (upstream pads 30 s and floors at the stream max)
    audio = eng.load_audio(audio)
    vad_spans = convert_timestamps = None
    if vad is not None:                    # T.py:294-296: the model only sees the glued speech
        audio, vad_spans, convert_timestamps = V.remove_non_speech(audio, vad)
    n_samples = int(audio.shape[-1])
    if chunks is None:
        cuts = [(0, n_samples)]
    else:
        step = int(round(float(chunks) * SAMPLE_RATE))
        assert step > 0
        cuts = [(s, min(s + step, n_samples)) for s in range(0, max(n_samples, 1), step)]
        condition_on_previous_text = False

    # ---- language (T.py:811-820 + upstream detection on the first window of the file)
    language_probs = None
    mels = [eng.log_mel(audio[s:e]) for (s, e) in cuts]
    language_detected = False
    if language is None:
        if not is_multilingual:
            language = "en"
        else:
            language_detected = True
            tok0 = get_tokenizer(True, num_languages=num_languages)
            # stdout as the reference + upstream produce it (T.py:817-820, 844-846, 1030-1032, 1073-1075; upstream
            # prints the result whenever verbose is not None).  With a VAD the inner call runs with verbose=False (T.py:286).
            inner_verbose = verbose if (vad is None or verbose is not True) else False
            if inner_verbose:
                print("Detecting language using up to the first 30 seconds. Use `--language` to specify the language")
            language, language_probs = eng.detect_language(mels[0], tok0)
            if inner_verbose is not None:
                print(f"Detected language: {LANGUAGES[language].title()}")
                sys.stdout.flush()
    language = language.lower() if language else language
    if language not in LANGUAGES and language in TO_LANGUAGE_CODE:
        language = TO_LANGUAGE_CODE[language]
    tokenizer = get_tokenizer(is_multilingual, num_languages=num_languages, language=language, task=task)
    temperatures = [float(t) for t in temperature] if isinstance(temperature, (list, tuple)) else [float(temperature)]
    setup = make_decode_setup(tokenizer, dims.n_text_ctx, sample_len=sample_len, suppress_tokens=suppress_tokens,
                              temperature=temperatures[0], beam_size=beam_size, patience=patience, best_of=best_of,
                              length_penalty=length_penalty)
    initial_prompt_tokens = tokenizer.encode(" " + initial_prompt.strip()) if initial_prompt is not None else []

    streams = []
    for i, ((s, e), mel) in enumerate(zip(cuts, mels)):
        content_frames = eng.mel_frames(mel) - N_FRAMES
        streams.append(_Stream(i, mel, content_frames, s / SAMPLE_RATE, initial_prompt_tokens))

    # ---- decode: the next window of every active stream in one GPU batch
    def consume(job, rec):
        st = streams[job["stream"]]
        if language_detected and job["stream"] == 0 and not st.records:
            rec.mel_from_language_detection = True        # first window of the file, see WindowRecord.max_duration
        st.consume(rec, tokenizer, no_speech_threshold, logprob_threshold, condition_on_previous_text)
        return st.next_job(setup) if st.active else None

    plain_greedy = len(temperatures) == 1 and temperatures[0] == 0 and setup.beam_size is None
    if plain_greedy and hasattr(eng, "decode_stream") and continuous_batching:
        # continuous batching: a stream's next window is admitted into the decode batch as soon as its previous one
        # finishes (no round barrier); per stream the windows are still decoded strictly in upstream's order
        eng.decode_stream([st.next_job(setup) for st in streams if st.active], setup, consume)
    else:
        while True:
            jobs = [st.next_job(setup) for st in streams if st.active]
            if not jobs:
                break
            records = decode_with_fallback(eng, jobs, setup, temperatures, tokenizer, compression_ratio_threshold,
                                           logprob_threshold, no_speech_threshold)
            for job, rec in zip(jobs, records):
                consume(job, rec)

    use_space = should_use_space(language)
    if naive_approach:
        # ---- two-pass strategy (T.py:1004-1338): pass 1 above was plain decoding; pass 2 re-runs the decoder teacher-forced
        # on every segment's own audio window and aligns all of its tokens at once
        assert chunks is None, "the two-pass strategy works on one sequential stream (chunks=None)"
        from . import naive as NV
        st = streams[0]
        all_segments = list(st.segments)
        all_words = NV.second_pass(eng, audio, all_segments, tokenizer, language, use_space=use_space,
                                   refine_nframes=refine_nframes, trust_whisper_timestamps=trust_whisper_timestamps,
                                   remove_punctuation_from_words=remove_punctuation_from_words,
                                   compute_word_confidence=compute_word_confidence,
                                   include_punctuation_in_confidence=include_punctuation_in_confidence,
                                   min_word_duration=0.0, detect_disfluencies=detect_disfluencies,
                                   verbose=bool(verbose) and vad is None)
        for w in all_words:
            w["_stream"] = 0
        text_parts = [tokenizer.decode(st.all_tokens[st.n_initial_prompt:])]
    else:
        # ---- replay the reference's segment/flush logic offline and align everything in one batch
        use_space = should_use_space(language)
        pending = []          # (stream, window idx, plan, AlignRequest)
        per_window = {}
        for st in streams:
            for wi, rec in enumerate(st.records):
                nxt = st.records[wi + 1].prompt if wi + 1 < len(st.records) else None
                reqs = {}

                def yields_words(plan, reqs=reqs):
                    # T.py:540-559: `ws` is empty when there is nothing between the timestamps or every word is a
                    # special token; this only depends on the tokens, so it is known before the DTW runs
                    req = None
                    if len(plan.tokens) > 1:
                        req = W.prepare_alignment(plan.tokens, plan.n_rows, tokenizer, use_space=use_space,
                                                  refine_nframes=refine_nframes,
                                                  remove_punctuation_from_words=remove_punctuation_from_words,
                                                  unfinished_decoding=plan.unfinished)
                    reqs[id(plan)] = req
                    if req is None:
                        return False
                    kept = req.words[1:] if req.unfinished else req.words[1:-1]
                    return any(not w.startswith("<|") for w in kept)

                plans, info = plan_window_alignment(rec, setup, nxt, yields_words)
                per_window[(st.index, wi)] = (plans, info)
                for plan in plans:
                    req = reqs[id(plan)]
                    for msg in req.warnings:
                        logger.warning(msg)
                    if rec.max_duration and req.f0 >= rec.max_duration:
                        logger.warning("Got start time outside of audio boundary")
                    pending.append((st, wi, plan, req))
        items = []
        for (st, wi, plan, req) in pending:
            if req is None:
                continue
            rec = st.records[wi]
            items.append(dict(window=rec.qk_window, row0=plan.row0, last_row=plan.row0 + req.row_offset_last,
                              T=req.T, f0=req.f0, F=req.F, max_dur=rec.max_duration or 0))
        lefts_list = None
        if items and detect_disfluencies:
            jumps_list, lefts_list = eng.align(items, disfluencies=True)
        else:
            jumps_list = eng.align(items) if items else []
        jit = iter(zip(jumps_list, lefts_list if lefts_list is not None else [None] * len(jumps_list)))

        # ---- per stream: words, confidences, compile (T.py:712-771, 912-1002)
        all_segments, all_words = [], []
        text_parts = []
        for st in streams:
            seg_words = []        # words of every flushed segment of this stream, in order
            seg_logprobs = []     # log-probs of the text tokens of the same segments
            seg_avglogprob = []
            seg_tokens = []       # the token list each flushed segment ended up with
            pend_iter = [p for p in pending if p[0] is st]
            by_window = {}
            for p in pend_iter:
                by_window.setdefault(p[1], []).append(p)
            for wi, rec in enumerate(st.records):
                plans, info = per_window[(st.index, wi)]
                ws_of_window, kept_plans = [], []
                for (_, _, plan, req) in by_window.get(wi, []):
                    ws = W.words_from_jumps(req, *next(jit), tokenizer=tokenizer) if req is not None else []
                    assert ws, "plan_window_alignment only keeps segments that yield words"
                    ws_of_window.append(ws)
                    kept_plans.append(plan)
                # chunk-level log-probs and the silence rule (T.py:712-748)
                should_skip = False
                if compute_word_confidence or no_speech_threshold is not None:
                    should_skip = (rec.no_speech_prob > no_speech_threshold) if no_speech_threshold is not None else False
                    lp = np.array(rec.logprobs, dtype=np.float32)
                    n = len(lp)
                    last_unfinished = bool(kept_plans) and kept_plans[-1].unfinished and plans and plans[-1] is kept_plans[-1] \
                        and info["final_unfinished"]
                    if last_unfinished:
                        fallback = kept_plans[-1].appended_token
                        chosen_last = rec.tokens[n - 1] if n - 1 < len(rec.tokens) else tokenizer.eot
                        if fallback != chosen_last:
                            lp[-1] = rec.last_row_logprobs(fallback)
                        ws_of_window[-1][-1]["avg_logprob_reliable"] = kept_plans[-1].last_token_reliable
                        n += 1
                    elif info["reached"] and ws_of_window:
                        ws_of_window[-1][-1]["avg_logprob_reliable"] = (setup.temperature == 0)
                    assert np.all(np.isfinite(lp)), "Got infinite logprob"
                    avg_logprob = float(lp.sum(dtype=np.float32)) / n if n else 0.0
                    if logprob_threshold is not None and avg_logprob > logprob_threshold:
                        should_skip = False
                if should_skip:
                    continue                       # upstream skipped this window too (no segments)
                for plan, ws in zip(kept_plans, ws_of_window):
                    seg_words.append(ws)
                    seg_tokens.append(list(plan.tokens))
                    if compute_word_confidence:
                        a = plan.row0 + 1          # skip the start timestamp
                        b = plan.row0 + len(plan.tokens) - (0 if (plan.unfinished and plan is kept_plans[-1] and info["final_unfinished"]) else 1)
                        seg_logprobs.append(lp[a:b])
                        seg_avglogprob.append(avg_logprob)
                    else:
                        seg_logprobs.append(None)
                        seg_avglogprob.append(None)

            whisper_segments = [s for s in st.segments if s["text"]] if any(not s["text"] for s in st.segments) \
                else list(st.segments)
            l1, l2 = len(whisper_segments), len(seg_words)
            assert l1 == l2 or l1 == 0, \
                f"Inconsistent number of segments: whisper_segments ({l1}) != timestamped_word_segments ({l2})"
            special0 = min(tokenizer.sot, tokenizer.eot)

            def strip_special(toks):
                toks = list(toks)
                while toks and toks[0] >= special0:
                    toks = toks[1:]
                while toks and toks[-1] >= special0:
                    toks = toks[:-1]
                return toks

            for i, (segment, ws, lps, avglp, flushed) in enumerate(zip(whisper_segments, seg_words, seg_logprobs,
                                                                        seg_avglogprob, seg_tokens)):
                # T.py:941-957: the tokens the state machine flushed vs the tokens upstream kept
                ours, theirs = strip_special(flushed), strip_special(segment["tokens"])
                if ours != theirs:
                    if len(ours) == len(theirs) + 1:
                        logger.warning(f"An additional token was added on segment {i}")
                    elif len(theirs) == 0:
                        logger.warning(f"Whisper has empty segment {i}")
                        assert segment["end"] == segment["start"], f"Fatal Error: Got empty segment {i} with non-zero duration"
                        segment["tokens"] = ours
                        segment["text"] = tokenizer.decode(ours)
                    else:
                        assert len(ours) < len(theirs) and ours == theirs[:len(ours)], \
                            f"Fatal Error: Got inconsistent text for segment {i}:\n{ours}\n!=\n{theirs}"
                        segment["tokens"] = list(flushed)
                        segment["text"] = tokenizer.decode(segment["tokens"])
                        logger.warning(f"Text had to be shortned on segment {i}")
                    ws[-1]["avg_logprob_reliable"] = False
                offset = segment["seek"] * HOP_LENGTH / SAMPLE_RATE
                for w in ws:
                    w["start"] += offset
                    w["end"] += offset
                    w["idx_segment"] = len(all_segments) + i      # index in the FILTERED list, used on the full list (as T.py:963 / 329-331)
                if compute_word_confidence:
                    if ws[-1].get("avg_logprob_reliable", True):
                        if abs(segment["avg_logprob"] - avglp) >= 1e-2:
                            logger.warning(f"Recomputed different logprob for segment {i}: {avglp} != {segment['avg_logprob']}")
                    if include_punctuation_in_confidence:
                        segment["confidence"] = W.round_confidence(float(np.exp(lps.mean(dtype=np.float32))))
                    nopunc = []
                    i_end = 0
                    for w in ws:
                        i_start = i_end
                        pieces = w["tokens"]
                        i_end += len(pieces)
                        assert i_end <= len(lps), f"Fatal Error: Got out-of-bound index for segment {i}: {i_end} > {len(lps)}"
                        if include_punctuation_in_confidence:
                            wl = lps[i_start:i_end]
                        else:
                            while len(pieces) > 1 and len(pieces[-1]) and pieces[-1][-1] in W.PUNCTUATION:
                                pieces = pieces[:-1]
                            wl = lps[i_start:i_start + len(pieces)]
                            nopunc.append(wl)
                        w["confidence"] = W.round_confidence(float(np.exp(wl.mean(dtype=np.float32))) if len(wl) else 0.0)
                    if i_end not in (len(lps), len(lps) - 1):
                        logger.warning(f"Got inconsistent length for segment {i} ({len(lps)} != {i_end}). Some words have been ignored.")
                    if not include_punctuation_in_confidence:
                        cat = np.concatenate(nopunc) if nopunc else np.zeros(0, np.float32)
                        segment["confidence"] = W.round_confidence(float(np.exp(cat.mean(dtype=np.float32))))
                for w in ws:
                    w["_stream"] = st.index
                all_words.extend(ws)
            # stream time shift (independent cuts) is applied after the per-window offsets
            if st.time_shift:
                for w in (w for ws in seg_words for w in ws):
                    w["start"] = W.round_timestamp(w["start"] + st.time_shift)
                    w["end"] = W.round_timestamp(w["end"] + st.time_shift)
                for s in st.segments:
                    s["start"] += st.time_shift
                    s["end"] += st.time_shift
                    s["seek"] += int(round(st.time_shift * SAMPLE_RATE / HOP_LENGTH))
            all_segments.extend(st.segments)              # empty-text segments stay in the output, like the reference
            text_parts.append(tokenizer.decode(st.all_tokens[st.n_initial_prompt:]))

    transcription = dict(text="".join(text_parts), segments=all_segments, language=language)
    if language_probs:
        transcription["language_probs"] = language_probs
    words = all_words

    # ---- post-processing, as T.py:313-357
    if remove_empty_words:
        transcription, words = W.remove_last_null_duration_words(transcription, words, recompute_text=True)
    # independent cuts are post-processed independently (each is "the reference run on that cut alone")
    for idx in sorted({w["_stream"] for w in words}):
        W.ensure_increasing_positions([w for w in words if w["_stream"] == idx],
                                      min_duration=min_word_duration if trust_whisper_timestamps else 0)
    segs = transcription["segments"]
    for word in words:
        if verbose and not naive_approach and vad is None:        # T.py:323-324
            print_timestamped(word)
        word.pop("tokens", None)
        word.pop("tokens_indices", None)
        word.pop("avg_logprob_reliable", None)
        word.pop("_stream", None)
        idx = word.pop("idx_segment")
        assert idx < len(segs), f"Fatal error: Got unexpected segment index {idx} >= {len(segs)}"
        seg = segs[idx]
        if "words" in seg:
            seg["words"].append(word)
        else:
            seg["words"] = [word]
            if refine_whisper_precision:
                seg["start"] = word["start"]
        if refine_whisper_precision:
            seg["end"] = word["end"]
    if chunks is not None:
        for i, seg in enumerate(segs):        # independent cuts: ids / seeks are those of the whole recording
            seg["id"] = i
    if vad is not None:
        # back to the time axis of the original audio (T.py:341-355)
        for seg in segs:
            for word in seg.get("words", []):
                word["start"], word["end"] = convert_timestamps(word["start"], word["end"])
                if verbose:                                        # T.py:346-347
                    print_timestamped(word)
            if refine_whisper_precision and len(seg.get("words", [])):
                seg["start"] = seg["words"][0]["start"]
                seg["end"] = seg["words"][-1]["end"]
            else:
                seg["start"], seg["end"] = convert_timestamps(seg["start"], seg["end"])
        transcription["speech_activity"] = [{"start": s, "end": e} for (s, e) in vad_spans]
    if hasattr(eng, "release"):
        eng.release()
    return transcription


def print_timestamped(w):
    """`[mm:ss.mmm --> mm:ss.mmm] text` on stdout (T.py:1363-1368)."""
    from .make_subtitles import format_timestamp
    line = f"[{format_timestamp(w['start'])} --> {format_timestamp(w['end'])}] {w['text']}\n"
    sys.stdout.write(line.encode(sys.getdefaultencoding(), errors="replace").decode())
    sys.stdout.flush()


transcribe = transcribe_timestamped
