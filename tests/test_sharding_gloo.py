"""N>1 path on CPU: two `gloo` ranks each transcribe their contiguous run of independent 30-s cuts (device work
done by the test-only OracleEngine) and rank 0 stitches the results; the stitched result must equal the
single-process `chunks=30` result of the whole recording.  Also unit-tests the split itself."""
import json
import os
import socket
import sys
import tempfile
from types import SimpleNamespace

import numpy as np
import pytest
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

AUDIO = (75.0, 11)
CHUNK = 30.0
LANGUAGE = None          # detected on the first 30 s of the whole recording by every rank


def _model_and_engine():
    sys.path.insert(0, HERE)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "whisper-timestamped_b200"))
    from oracle_engine import OracleEngine, build_oracle_model
    from whisper_timestamped import model_zoo as zoo
    dims = zoo.DIMS["tiny"]
    sd = zoo.synthetic_state_dict(dims, seed=1234)
    heads = zoo.ALIGNMENT_HEADS["tiny"]
    om = build_oracle_model(dims, sd, heads)
    shim = SimpleNamespace(dims=dims, is_multilingual=om.is_multilingual, num_languages=om.num_languages)
    return shim, OracleEngine(om, heads)


def _worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist
    torch.set_num_threads(2)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shim, eng = _model_and_engine()
        from whisper_timestamped.sharding import transcribe_sharded
        from whisper_timestamped.synthetic_audio import synthetic_speech
        audio = synthetic_speech(*AUDIO)
        res = transcribe_sharded(shim, audio, CHUNK, rank, world, language=LANGUAGE, engine=eng)
        if rank == 0:
            with open(out_path, "w") as f:
                json.dump(res, f)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_chunk_range_is_a_partition():
    from whisper_timestamped.sharding import chunk_range
    for n in (0, 1, 2, 7, 120, 121):
        for world in (1, 2, 3, 4, 8):
            spans = [chunk_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        chunk_range(4, 2, 2)


def test_shard_audio_offsets():
    from whisper_timestamped.sharding import shard_audio
    audio = np.arange(int(75 * 16000), dtype=np.float32)
    a0, off0, r0 = shard_audio(audio, 30.0, 0, 2)
    a1, off1, r1 = shard_audio(audio, 30.0, 1, 2)
    assert r0 == (0, 1) and r1 == (1, 3)
    assert off0 == 0.0 and off1 == 30.0
    assert len(a0) == 30 * 16000 and len(a1) == 45 * 16000
    assert a1[0] == 30 * 16000


@pytest.mark.timeout(600)
def test_two_gloo_ranks_equal_single_process():
    shim, eng = _model_and_engine()
    from whisper_timestamped.synthetic_audio import synthetic_speech
    from whisper_timestamped.transcribe import transcribe_timestamped
    audio = synthetic_speech(*AUDIO)
    whole = transcribe_timestamped(shim, audio, language=LANGUAGE, engine=eng, chunks=CHUNK)
    assert [s["id"] for s in whole["segments"]] == list(range(len(whole["segments"])))
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "merged.json")
        mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
        merged = json.load(open(out))
    assert merged["text"] == whole["text"] and merged["language"] == whole["language"]
    assert set(merged["language_probs"]) == set(whole["language_probs"])
    for k, v in whole["language_probs"].items():
        assert abs(merged["language_probs"][k] - v) < 1e-6
    assert len(merged["segments"]) == len(whole["segments"]) > 2
    for a, b in zip(merged["segments"], whole["segments"]):
        assert a["id"] == b["id"] and a["seek"] == b["seek"] and a["tokens"] == b["tokens"] and a["text"] == b["text"]
        assert abs(a["start"] - b["start"]) < 1e-6 and abs(a["end"] - b["end"]) < 1e-6
        assert [w["text"] for w in a.get("words", [])] == [w["text"] for w in b.get("words", [])]
        for x, y in zip(a.get("words", []), b.get("words", [])):
            assert abs(x["start"] - y["start"]) < 1e-6 and abs(x["end"] - y["end"]) < 1e-6
            assert abs(x["confidence"] - y["confidence"]) < 1e-6
    # every rank-1 segment lies in the second shard's time range
    assert any(s["start"] >= 30.0 for s in merged["segments"])
