"""Decode-step probe: times one decode step of the large model at a fixed number of active windows,
either through the captured CUDA graph (default) or as plain launches (--eager, for an ncu launch list).

  python tools/step_probe.py --active 128 --cap 128 --steps 24
  ncu --cache-control none --metrics gpu__time_duration.sum ... python tools/step_probe.py --eager --steps 2
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "whisper-timestamped_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="synthetic:large-v3")
    ap.add_argument("--cap", type=int, default=128)
    ap.add_argument("--active", type=str, default="128,32,1")
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--eager", action="store_true")
    args = ap.parse_args()
    import whisper_timestamped as wt
    from whisper_timestamped.engine import CudaEngine
    from whisper_timestamped.tokenizer import get_tokenizer
    from whisper_timestamped.windows import make_decode_setup

    m = wt.load_model(args.model, device="cuda")
    eng = CudaEngine(m, max_batch=args.cap)
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
    dev = torch.device("cuda")
    for n_active in [int(a) for a in args.active.split(",")]:
        tokens = np.zeros((cap, m.dims.n_text_ctx + 1), dtype=np.int32)
        tokens[:, :P] = prompt
        tokens[:, P:P + 8] = 1000
        ses["tokens"].copy_(torch.from_numpy(tokens))
        nt = np.full(cap, P + 8, dtype=np.int32)
        ses["n_tokens"].copy_(torch.from_numpy(nt))
        ses["n_prompt"].copy_(torch.from_numpy(np.full(cap, P, dtype=np.int32)))
        dn = np.ones(cap, dtype=np.int32)
        dn[:n_active] = 0
        ses["done"].copy_(torch.from_numpy(dn))
        if args.eager:
            for _ in range(args.steps):
                eng._step(ses)
            torch.cuda.synchronize()
            continue
        if ses["graph"] is None:
            eng._step(ses)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream(device=dev)
            s.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(s):
                with torch.cuda.graph(graph, stream=s):
                    eng._step(ses)
            torch.cuda.current_stream(dev).wait_stream(s)
            ses["graph"] = graph
        for _ in range(3):
            ses["graph"].replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            ses["graph"].replay()
        e1.record()
        torch.cuda.synchronize()
        print(f"cap={cap} active={n_active}: {e0.elapsed_time(e1) / args.steps:.3f} ms/step "
              f"(n_tokens now {int(ses['n_tokens'][0])})", flush=True)


if __name__ == "__main__":
    main()
