"""Multi-GPU sharding of one long recording: one process per GPU, each transcribes a contiguous run of
independent fixed-length cuts (the `chunks=` streams of transcribe.py) and rank 0 stitches the results.

There is NO data-path collective: the cuts are independent (each restarts the decoder prompt, exactly
like running the reference on the cut alone), so the only exchange is the final gather of the (small)
result dictionaries.  The reference itself has no multi-device path; this is the
data-parallel extension named by BASELINE.json's north_star.
"""
import json

import numpy as np

SAMPLE_RATE = 16000


def chunk_range(n_chunks: int, rank: int, world: int):
    """Contiguous, disjoint, covering split of `n_chunks` cuts over `world` ranks (sizes differ by at most 1)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return rank * n_chunks // world, (rank + 1) * n_chunks // world


def shard_audio(audio, chunk_seconds: float, rank: int, world: int):
    """Returns (samples of this rank, time offset in seconds, (lo, hi) cut range)."""
    step = int(round(float(chunk_seconds) * SAMPLE_RATE))
    n = int(audio.shape[-1])
    n_chunks = (n + step - 1) // step
    lo, hi = chunk_range(n_chunks, rank, world)
    return audio[lo * step: min(hi * step, n)], lo * step / SAMPLE_RATE, (lo, hi)


def shift_segments(segments, offset: float):
    """Moves every time stamp of a shard's result by the shard's start time (in place)."""
    if offset == 0:
        return segments
    for s in segments:
        s["start"] = round(s["start"] + offset, 2)
        s["end"] = round(s["end"] + offset, 2)
        if "seek" in s:
            s["seek"] = int(s["seek"] + round(offset * 100))
        for w in s.get("words", []):
            w["start"] = round(w["start"] + offset, 2)
            w["end"] = round(w["end"] + offset, 2)
    return segments


def merge_results(shard_results):
    """Stitches per-shard result dicts (already time-shifted, in rank order) into one reference-shaped result."""
    segs = []
    for r in shard_results:
        for s in r["segments"]:
            s = dict(s)
            s["id"] = len(segs)
            segs.append(s)
    out = dict(shard_results[0])
    out["segments"] = segs
    out["text"] = "".join(r["text"] for r in shard_results)
    return out


def gather_results(result, rank: int, world: int, group=None):
    """all_gather of the JSON-encoded shard results (works on NCCL and gloo); rank 0 gets the merged dict,
    the other ranks get their own shard back."""
    if world == 1:
        return result
    import torch.distributed as dist
    gathered = [None] * world
    dist.all_gather_object(gathered, json.dumps(result, default=_np_default), group=group)
    if rank != 0:
        return result
    return merge_results([json.loads(g) for g in gathered])


def _np_default(o):
    if isinstance(o, np.floating):
        return float(o)
    if isinstance(o, np.integer):
        return int(o)
    if isinstance(o, np.ndarray):
        return o.tolist()
    raise TypeError(type(o))


def transcribe_sharded(model, audio, chunk_seconds: float, rank: int, world: int, group=None, **kwargs):
    """transcribe() of this rank's cuts + gather.  `audio` is the WHOLE recording (numpy or tensor) on every rank."""
    from .tokenizer import get_tokenizer
    from .transcribe import transcribe_timestamped
    language_probs = None
    if kwargs.get("language") is None and model.is_multilingual:
        # every rank detects the language on the FIRST 30 s of the whole recording (what the reference does for a file),
        # not on its own shard: identical on all ranks without a collective
        eng = kwargs.get("engine") or model.engine()
        head = eng.load_audio(audio[..., : 30 * SAMPLE_RATE])
        tok0 = get_tokenizer(True, num_languages=model.num_languages)
        language, language_probs = eng.detect_language(eng.log_mel(head), tok0)
        kwargs = dict(kwargs, language=language)
    mine, offset, _ = shard_audio(audio, chunk_seconds, rank, world)
    res = transcribe_timestamped(model, mine, chunks=chunk_seconds, **kwargs)
    if language_probs:
        res["language_probs"] = language_probs
    shift_segments(res["segments"], offset)
    return gather_results(res, rank, world, group)
