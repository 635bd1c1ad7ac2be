"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU engine with the oracle's stand-in for openai-whisper (oracle/upstream/whisper, fp32 torch) and the
oracle alignment numerics (scipy median filter + torch CPU ops + oracle DTW), exposed through the
engine protocol of the drop-in's transcribe().  Used by
  * tests: lets the product's HOST logic run on the CPU-only build box against the golden outputs of the
    unmodified reference, and checks the CUDA engine window by window on the GPU;
  * bench.py --impl reference / cpu_baseline: "the reference's CPU path" timed on the GPU box's host cores
    (the real reference and its dependencies do not exist there).
The product never imports this module."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UP = os.path.join(ROOT, "oracle", "upstream")
for _p in (UP, os.path.join(ROOT, "whisper-timestamped_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import whisper  # noqa: E402  (oracle stand-in)
from whisper.model import disable_sdpa  # noqa: E402

import oracle  # noqa: E402
from oracle.prep import attn_cost  # noqa: E402
from whisper_timestamped.windows import WindowRecord  # noqa: E402


def build_oracle_model(dims, state_dict, alignment_heads):
    m = whisper.Whisper(whisper.ModelDimensions(**dims.asdict()))
    m.load_state_dict(state_dict)
    mask = torch.zeros(dims.n_text_layer, dims.n_text_head, dtype=torch.bool)
    for l, h in alignment_heads:
        mask[l, h] = True
    m.register_buffer("alignment_heads", mask.to_sparse(), persistent=False)
    return m.eval()


class _Recorder:
    """Appended after the real logit filters: sees exactly what hook_output_logits computes
    (T.py:871-875)."""

    def __init__(self):
        self.rows = []

    def apply(self, logits, tokens):
        self.rows.append(F.log_softmax(logits.float(), dim=-1).clone())


class OracleEngine:
    def __init__(self, model, alignment_heads, keep_logprobs=False):
        self.model = model
        self.heads = list(alignment_heads)
        self.qk = []              # per window: float32 [N, rows, 1500]
        self.full_logprobs = []   # per window: [rows, V] (only when keep_logprobs)
        self.keep_logprobs = keep_logprobs

    # ---- audio
    def load_audio(self, audio):
        if isinstance(audio, np.ndarray):
            audio = torch.from_numpy(audio)
        return audio.float()

    def log_mel(self, audio, pad_30s=True):
        # pad_30s=False: the two-pass strategy's per-segment mel (reference T.py:1213: no padding argument)
        return whisper.log_mel_spectrogram(audio, self.model.dims.n_mels, padding=whisper.audio.N_SAMPLES if pad_30s else 0)

    def mel_frames(self, mel):
        return mel.shape[-1]

    def detect_language(self, mel, tokenizer):
        seg = whisper.pad_or_trim(mel, whisper.audio.N_FRAMES)
        _, probs = self.model.detect_language(seg)
        return max(probs, key=probs.get), probs

    # ---- decode
    @torch.no_grad()
    def decode_windows(self, jobs, setup):
        out = []
        tok = setup.tokenizer
        for job in jobs:
            mel = job["mel"][:, job["seek"]: job["seek"] + job["segment_size"]]
            mel = whisper.pad_or_trim(mel, whisper.audio.N_FRAMES)
            prompt = job["prompt"]
            ptoks = prompt[1:-len(tok.sot_sequence)] if prompt[0] == tok.sot_prev else None
            opts = whisper.DecodingOptions(task=tok.task or "transcribe", language=tok.language or "en",
                                           temperature=setup.temperature, beam_size=getattr(setup, "beam_size", None),
                                           patience=getattr(setup, "patience", None), best_of=getattr(setup, "best_of", None),
                                           length_penalty=getattr(setup, "length_penalty", None),
                                           sample_len=setup.sample_len, prompt=ptoks or None, fp16=False,
                                           suppress_tokens=list(setup.suppress_tokens) or None)
            task = whisper.decoding.DecodingTask(self.model, opts)
            if opts.beam_size is not None or opts.temperature > 0:
                # beam search / sampling (two-pass strategy): only the selected hypothesis matters to the caller
                for f in task.logit_filters:
                    if isinstance(f, whisper.decoding.SuppressTokens):
                        f.suppress_tokens = list(setup.suppress_tokens)
                assert list(task.initial_tokens) == list(prompt), (task.initial_tokens, prompt)
                res = task.run(mel.unsqueeze(0))[0]
                tokens = list(res.tokens)
                self.qk.append(None)
                out.append(WindowRecord(seek=job["seek"], segment_size=job["segment_size"], prompt=list(prompt), tokens=tokens,
                                        logprobs=None, ended_by_eot=True, no_speech_prob=float(res.no_speech_prob),
                                        qk_window=len(self.qk) - 1, temperature=float(setup.temperature), language=tok.language,
                                        sum_logprob=float(res.avg_logprob) * (len(tokens) + 1)))
                continue
            # the product computed the suppress list itself; make the stand-in use exactly that list
            for f in task.logit_filters:
                if isinstance(f, whisper.decoding.SuppressTokens):
                    f.suppress_tokens = list(setup.suppress_tokens)
            assert list(task.initial_tokens) == list(prompt), (task.initial_tokens, prompt)
            rec = _Recorder()
            task.logit_filters.append(rec)
            rows = [[] for _ in self.model.decoder.blocks]
            hooks = []
            for i, blk in enumerate(self.model.decoder.blocks):
                hooks.append(blk.cross_attn.register_forward_hook(
                    lambda layer, ins, outs, index=i: rows[index].append(outs[-1][:, :, -1:, :])))
            try:
                with disable_sdpa():
                    res = task.run(mel.unsqueeze(0))[0]
            finally:
                for h in hooks:
                    h.remove()
            layers = [torch.cat(r, dim=-2)[0] for r in rows]                   # per layer [H, rows, 1500]
            qk = torch.stack([layers[l][h] for (l, h) in self.heads]).float()  # [N, rows, 1500]
            lp_rows = torch.cat(rec.rows, dim=0)                               # [rows, V]
            tokens = list(res.tokens)
            n_rows = lp_rows.shape[0]
            ended = n_rows == len(tokens) + 1
            chosen = tokens + ([tok.eot] if ended else [])
            assert len(chosen) == n_rows, (len(chosen), n_rows)
            lps = lp_rows[torch.arange(n_rows), torch.tensor(chosen)].numpy().astype(np.float32)
            last = lp_rows[-1].clone()
            self.qk.append(qk)
            if self.keep_logprobs:
                self.full_logprobs.append(lp_rows)
            out.append(WindowRecord(seek=job["seek"], segment_size=job["segment_size"], prompt=list(prompt),
                                    tokens=tokens, logprobs=lps, ended_by_eot=ended,
                                    no_speech_prob=float(res.no_speech_prob), qk_window=len(self.qk) - 1,
                                    temperature=0.0, language=tok.language,
                                    last_row_logprobs=lambda t, last=last: float(last[t])))
        return out

    # ---- teacher-forced pass of the two-pass strategy (reference T.py:1213-1249)
    @torch.no_grad()
    def teacher_forced(self, mel, tokens_in, i_start, pairs):
        """mel: un-padded [n_mels, frames]; tokens_in: sot sequence + <|0.00|> + text tokens.  Registers the alignment
        heads' cross-attention rows from position i_start-1 on as a new window and returns (window id, float32
        log_softmax(logits[step])[token] for every (step, token) in `pairs`)."""
        mfcc = whisper.pad_or_trim(mel, whisper.audio.N_FRAMES).unsqueeze(0)
        captured = [None] * len(self.model.decoder.blocks)
        hooks = []
        for i, blk in enumerate(self.model.decoder.blocks):
            hooks.append(blk.cross_attn.register_forward_hook(
                lambda layer, ins, outs, index=i: captured.__setitem__(index, outs[-1])))
        try:
            with disable_sdpa():
                logits = self.model(mfcc, torch.tensor(list(tokens_in), dtype=torch.int32).unsqueeze(0))
        finally:
            for h in hooks:
                h.remove()
        lp = torch.nn.functional.log_softmax(logits, dim=-1)
        qk = torch.stack([captured[l][0, h, i_start - 1:, :] for (l, h) in self.heads]).float()   # [N, rows, 1500]
        self.qk.append(qk)
        vals = np.array([float(lp[0, s_, t_]) for (s_, t_) in pairs], dtype=np.float32)
        return len(self.qk) - 1, vals

    # ---- alignment numerics (oracle)
    def align(self, items, disfluencies=False):
        out, lefts = [], []
        for it in items:
            qk = self.qk[it["window"]]
            T, row0, last_row = it["T"], it["row0"], it["last_row"]
            rows = list(range(row0, row0 + T - 1)) + [last_row]
            sel = qk[:, rows, :].numpy()
            cost = attn_cost(sel, it["f0"], it["f0"] + it["F"], max_duration=it["max_dur"] or None)
            _, _, jumps, _ = oracle.dtw_symmetric1(cost)
            out.append(jumps)
            if disfluencies:
                # reference T.py:1656-1670, with scipy itself
                from scipy.signal import find_peaks
                w = -np.asarray(cost, dtype=np.float64)
                left = []
                for i_token, (begin, end) in enumerate(zip(jumps[:-1], jumps[1:])):
                    peaks, props = find_peaks(w[i_token, begin:end], width=3, prominence=0.02)
                    left.append(int(round(props["left_ips"][-1])) if len(peaks) > 1 else -1)
                lefts.append(np.asarray(left, dtype=np.int32))
        return (out, lefts) if disfluencies else out
