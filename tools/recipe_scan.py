"""Dev tool (GPU): scan the synthetic-weight recipe knobs (timestamp offset / EOT logit) for a model size and print
decode statistics, to pick defaults that make greedy decoding look like speech (windows ending with
<|endoftext|> after ~100 tokens, several closed segments per window)."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "whisper-timestamped_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import whisper_timestamped as wt
from whisper_timestamped.engine import CudaEngine
from whisper_timestamped.model import SB16
from whisper_timestamped import model_zoo as zoo
from bench import make_audio

name = sys.argv[1] if len(sys.argv) > 1 else "large-v3"
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 600.0
settings = [tuple(map(float, s.split(","))) for s in sys.argv[3:]] or [(2.5, 11.0)]
model = wt.load_model(f"synthetic:{name}", device="cuda:0")
eot, sot, nl, tsb = zoo.special_token_layout(model.dims.n_vocab)
audio = make_audio(secs)
for (tso, eol) in settings:
    model.w.emb[tsb:, 0] = tso
    model.w.emb[eot, 0] = eol
    model.w.emb_sb = SB16.from_f32(model.w.emb)
    eng = CudaEngine(model, max_batch=128)
    import whisper_timestamped.transcribe as T
    recs = []
    orig = eng.decode_windows
    def spy(jobs, setup, orig=orig):
        r = orig(jobs, setup); recs.extend(r); return r
    eng.decode_windows = spy
    res = wt.transcribe(model, audio, language="en", chunks=30.0, engine=eng)
    lens = np.array([len(r.tokens) for r in recs]); eots = np.array([r.ended_by_eot for r in recs])
    nts = np.array([sum(t >= tsb for t in r.tokens) for r in recs])
    kept = sum(len(s["tokens"]) for s in res["segments"])
    print(json.dumps(dict(ts_offset=tso, eot_logit=eol, windows=len(recs), mean_len=float(lens.mean()), p90_len=float(np.percentile(lens, 90)),
                          max_len=int(lens.max()), eot_frac=float(eots.mean()), mean_ts=float(nts.mean()), segments=len(res["segments"]),
                          kept_tokens=kept, words=sum(len(s.get("words", [])) for s in res["segments"]))))
