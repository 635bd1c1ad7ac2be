"""Decode-step probe: milliseconds per decode step of a model at a fixed number of active windows, for the two step
implementations that share the session state:
  graph  : the per-operator step (tensor-core skinny GEMMs, one kernel per operator) replayed as ONE CUDA graph
  steps  : the persistent small-batch kernel (wts_decode_steps, <= 32 active windows), `--steps` tokens per launch

  python tools/step_probe.py --active 128,32,16,8,4,1 --cap 128 --steps 24
  ncu ... python tools/step_probe.py --eager --steps 2        (plain launches, for an ncu launch list)
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "whisper-timestamped_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="synthetic:large-v3")
    ap.add_argument("--cap", type=int, default=128)
    ap.add_argument("--active", type=str, default="128,32,16,8,4,1")
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--eager", action="store_true")
    ap.add_argument("--only", default="graph,lean_mma,lean,steps")
    args = ap.parse_args()
    import whisper_timestamped as wt
    from whisper_timestamped.engine import CudaEngine
    from whisper_timestamped.tokenizer import get_tokenizer
    from whisper_timestamped.windows import make_decode_setup

    m = wt.load_model(args.model, device="cuda")
    eng = CudaEngine(m, max_batch=args.cap, small_batch_rows=32)
    tok = get_tokenizer(m.is_multilingual, num_languages=m.num_languages, language="en", task="transcribe")
    setup = make_decode_setup(tok, m.dims.n_text_ctx)
    ses = eng._decoder_session(setup, args.cap)
    cap = ses["cap"]
    g = torch.Generator(device="cuda").manual_seed(1)
    for li in range(m.dims.n_text_layer):
        for name in ("ck", "cv"):
            t = ses["st8"][name][li]
            t.copy_((torch.randn(t.shape, device="cuda", generator=g) * 0.5).to(t.dtype))
        ses["st8"]["ckal"][li].normal_(0, 0.5, generator=g)
    ses["suppress"].zero_()
    ses["suppress"][tok.eot] = 1                 # keep every window alive for the whole probe
    ses["blank"].zero_()
    prompt = list(tok.sot_sequence)
    P = len(prompt)

    def reset(n_active):
        tokens = np.zeros((cap, m.dims.n_text_ctx + 1), dtype=np.int32)
        tokens[:, :P] = prompt
        tokens[:, P:P + 8] = 1000
        ses["tokens"].copy_(torch.from_numpy(tokens))
        ses["n_tokens"].copy_(torch.from_numpy(np.full(cap, P + 8, dtype=np.int32)))
        ses["n_prompt"].copy_(torch.from_numpy(np.full(cap, P, dtype=np.int32)))
        dn = np.ones(cap, dtype=np.int32)
        dn[:n_active] = 0
        ses["done"].copy_(torch.from_numpy(dn))

    out = {}
    for n_active in [int(a) for a in args.active.split(",")]:
        res = {}
        reset(n_active)
        if args.eager:
            for _ in range(args.steps):
                eng._step(ses)
            torch.cuda.synchronize()
            continue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if "graph" in args.only:
            graph = eng._step_graph(ses)
            for _ in range(3):
                graph.replay()
            e0.record()
            for _ in range(args.steps):
                graph.replay()
            e1.record()
            torch.cuda.synchronize()
            res["graph_ms_per_step"] = round(e0.elapsed_time(e1) / args.steps, 3)
        for name, mma in (("lean_mma", True), ("lean", False)):
            if name in args.only.split(",") and ses["steps"] is not None and n_active <= eng.small_batch_rows:
                reset(n_active)
                eng.small_batch_mma = mma
                graph = eng._lean_graph(ses, n_active)
                for _ in range(3):
                    graph.replay()
                e0.record()
                for _ in range(args.steps):
                    graph.replay()
                e1.record()
                torch.cuda.synchronize()
                res[name + "_ms_per_step"] = round(e0.elapsed_time(e1) / args.steps, 3)
        if "steps" in args.only and ses["steps"] is not None and n_active <= eng.small_batch_rows:
            reset(n_active)
            eng._run_steps(ses, 3, n_active)
            torch.cuda.synchronize()
            e0.record()
            eng._run_steps(ses, args.steps, n_active)
            e1.record()
            torch.cuda.synchronize()
            flags = ses["steps"]["keep"]["sync"].cpu().numpy()
            res["steps_ms_per_step"] = round(e0.elapsed_time(e1) / max(1, int(flags[2])), 3)
            res["steps_completed"] = int(flags[2])
            res["barrier_timeout"] = int(flags[1])
            # phase timeline of ONE step (CTA 0's %globaltimer after every grid barrier)
            L = m.dims.n_text_layer
            nb = 8 * L + 3
            prof = torch.zeros(nb + 2, dtype=torch.int64, device="cuda")
            p = ses["steps"]["args"]
            p.prof, p.prof_cap = prof.data_ptr(), nb + 2
            reset(n_active)
            eng._run_steps(ses, 1, n_active)
            torch.cuda.synchronize()
            p.prof, p.prof_cap = None, 0
            t = prof.cpu().numpy().astype(np.float64)
            d = np.diff(t[: nb + 1]) / 1e3                      # microseconds per phase (incl. its closing barrier)
            names = ["qkv", "self_attn", "out", "cross_q", "cross_attn", "cross_out", "fc1", "fc2"]
            per = {n: round(float(d[1 + i: 1 + 8 * L: 8].mean()), 2) for i, n in enumerate(names)}
            res["phase_us"] = dict(embed=round(float(d[0]), 2), **per, logits=round(float(d[1 + 8 * L]), 2),
                                   select=round(float(d[2 + 8 * L]), 2), total=round(float(d.sum()), 1))
        out[n_active] = res
        print(f"active {n_active:4d} / cap {cap}: {res}", flush=True)
    print(json.dumps({"model": args.model, "cap": cap, "steps": args.steps, "ms": out}))


if __name__ == "__main__":
    main()
