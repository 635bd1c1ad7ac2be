"""Model dimension tables, alignment-head tables and the synthetic-weight recipe.

* `DIMS`: the official openai-whisper model dimensions (upstream `ModelDimensions`; mirrored by
  `states_to_dim`, /root/reference/whisper_timestamped/transcribe.py:2909-2923).
* `ALIGNMENT_HEADS`: decoded form of the base85 masks at transcribe.py:2343-2357 — the
  (decoder layer, head) pairs whose cross-attention carries word timing.
* `synthetic_state_dict`: there are no checkpoints in this environment (no network), so benchmarks
  and parity tests run on seeded synthetic weights of the exact architecture.  The recipe makes
  greedy decoding behave like speech: logits have a spread of ~4 nats, timestamp tokens share a
  fixed offset so that timestamp pairs appear every ~15 tokens, and <|endoftext|> carries a
  deterministic logit that only wins once the timestamps approach 30 s.
"""
import math
from dataclasses import dataclass, asdict

import torch


@dataclass
class ModelDimensions:
    n_mels: int
    n_audio_ctx: int
    n_audio_state: int
    n_audio_head: int
    n_audio_layer: int
    n_vocab: int
    n_text_ctx: int
    n_text_state: int
    n_text_head: int
    n_text_layer: int

    def asdict(self):
        return asdict(self)


def _dims(n_mels, d, h, enc_l, dec_l, vocab):
    return ModelDimensions(n_mels, 1500, d, h, enc_l, vocab, 448, d, h, dec_l)


DIMS = {
    "tiny.en": _dims(80, 384, 6, 4, 4, 51864), "tiny": _dims(80, 384, 6, 4, 4, 51865),
    "base.en": _dims(80, 512, 8, 6, 6, 51864), "base": _dims(80, 512, 8, 6, 6, 51865),
    "small.en": _dims(80, 768, 12, 12, 12, 51864), "small": _dims(80, 768, 12, 12, 12, 51865),
    "medium.en": _dims(80, 1024, 16, 24, 24, 51864), "medium": _dims(80, 1024, 16, 24, 24, 51865),
    "large-v1": _dims(80, 1280, 20, 32, 32, 51865), "large-v2": _dims(80, 1280, 20, 32, 32, 51865),
    "large-v3": _dims(128, 1280, 20, 32, 32, 51866), "large": _dims(128, 1280, 20, 32, 32, 51866),
    "large-v3-turbo": _dims(128, 1280, 20, 32, 4, 51866), "turbo": _dims(128, 1280, 20, 32, 4, 51866),
}

ALIGNMENT_HEADS = {
    "tiny.en": [(1, 0), (2, 0), (2, 5), (3, 0), (3, 1), (3, 2), (3, 3), (3, 4)],
    "tiny": [(2, 2), (3, 0), (3, 2), (3, 3), (3, 4), (3, 5)],
    "base.en": [(3, 3), (4, 7), (5, 1), (5, 5), (5, 7)],
    "base": [(3, 1), (4, 2), (4, 3), (4, 7), (5, 1), (5, 2), (5, 4), (5, 6)],
    "small.en": [(6, 6), (7, 0), (7, 3), (7, 8), (8, 2), (8, 5), (8, 7), (9, 0), (9, 4), (9, 8), (9, 10),
                 (10, 0), (10, 1), (10, 2), (10, 3), (10, 6), (10, 11), (11, 2), (11, 4)],
    "small": [(5, 3), (5, 9), (8, 0), (8, 4), (8, 7), (8, 8), (9, 0), (9, 7), (9, 9), (10, 5)],
    "medium.en": [(11, 4), (14, 1), (14, 12), (14, 14), (15, 4), (16, 0), (16, 4), (16, 9), (17, 12), (17, 14),
                  (18, 7), (18, 10), (18, 15), (20, 0), (20, 3), (20, 9), (20, 14), (21, 12)],
    "medium": [(13, 15), (15, 4), (15, 15), (16, 1), (20, 0), (23, 4)],
    "large-v1": [(9, 19), (11, 2), (11, 4), (11, 17), (22, 7), (22, 11), (22, 17), (23, 2), (23, 15)],
    "large-v2": [(10, 12), (13, 17), (16, 11), (16, 12), (16, 13), (17, 15), (17, 16), (18, 4), (18, 11),
                 (18, 19), (19, 11), (21, 2), (21, 3), (22, 3), (22, 9), (22, 12), (23, 5), (23, 7), (23, 13),
                 (25, 5), (26, 1), (26, 12), (27, 15)],
    "large-v3": [(7, 0), (10, 17), (12, 18), (13, 12), (16, 1), (17, 14), (19, 11), (21, 4), (24, 1), (25, 6)],
    "large-v3-turbo": [(2, 4), (2, 11), (3, 3), (3, 6), (3, 11), (3, 14)],
    "turbo": [(2, 4), (2, 11), (3, 3), (3, 6), (3, 11), (3, 14)],
}
ALIGNMENT_HEADS["large"] = ALIGNMENT_HEADS["large-v3"]


def default_alignment_heads(dims: ModelDimensions):
    """Upstream default for non-official checkpoints: every head of the top half of the layers."""
    return [(l, h) for l in range(dims.n_text_layer // 2, dims.n_text_layer) for h in range(dims.n_text_head)]


def special_token_layout(n_vocab: int):
    """(eot, sot, n_languages, timestamp_begin) implied by the vocabulary size."""
    multilingual = n_vocab >= 51865
    num_languages = n_vocab - 51765 - int(multilingual)
    eot = 50257 if multilingual else 50256
    sot = eot + 1
    timestamp_begin = sot + 1 + num_languages + 6
    return eot, sot, num_languages, timestamp_begin


def sinusoids(length, channels, max_timescale=10000):
    inc = math.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2))
    st = torch.arange(length)[:, None] * inv[None, :]
    return torch.cat([torch.sin(st), torch.cos(st)], dim=1)


def synthetic_state_dict(dims: ModelDimensions, seed: int = 1234, logit_spread: float = 4.0,
                         ts_offset: float = 2.5, eot_logit: float = 11.0, w_std: float = None,
                         res_scale: float = 3.0, cross_scale: float = 0.35):
    """Seeded fp32 state dict with openai-whisper key names.

    Coordinate 0 of the decoder's final LayerNorm is turned into a constant (weight 0, bias 1) and
    coordinate 0 of the tied embedding carries per-token logit offsets: `ts_offset` for every
    timestamp token, `eot_logit` for <|endoftext|> (whose random part is zeroed, so its logit is
    deterministic), 0 for text tokens.  All other embedding coordinates are N(0, s^2) with
    s = logit_spread / sqrt(d-1), so text/timestamp logits are ~N(offset, logit_spread^2).
    `res_scale` sets the size of every sub-layer's contribution to the residual stream; it must dwarf
    the (tied) input embedding, otherwise the model just repeats its previous token.
    """
    g = torch.Generator().manual_seed(seed)
    d_a, d_t = dims.n_audio_state, dims.n_text_state
    w_std = w_std if w_std is not None else 0.02

    def rnd(*shape, std=w_std):
        return torch.empty(*shape).normal_(0.0, std, generator=g)

    sd = {}
    sd["encoder.conv1.weight"] = rnd(d_a, dims.n_mels, 3, std=1.0 / math.sqrt(3 * dims.n_mels))
    sd["encoder.conv1.bias"] = rnd(d_a, std=0.1)
    sd["encoder.conv2.weight"] = rnd(d_a, d_a, 3, std=1.0 / math.sqrt(3 * d_a))
    sd["encoder.conv2.bias"] = rnd(d_a, std=0.1)
    sd["encoder.positional_embedding"] = sinusoids(dims.n_audio_ctx, d_a)

    def block(prefix, d, cross):
        names = ["attn"] + (["cross_attn"] if cross else [])
        for a in names:
            std_qk = 1.5 / math.sqrt(d)            # lively attention logits
            sd[f"{prefix}.{a}.query.weight"] = rnd(d, d, std=std_qk)
            sd[f"{prefix}.{a}.query.bias"] = rnd(d, std=0.1)
            sd[f"{prefix}.{a}.key.weight"] = rnd(d, d, std=std_qk)
            sd[f"{prefix}.{a}.value.weight"] = rnd(d, d, std=1.0 / math.sqrt(d))
            sd[f"{prefix}.{a}.value.bias"] = rnd(d, std=0.02)
            scale = res_scale * (cross_scale if a == "cross_attn" else 1.0)
            sd[f"{prefix}.{a}.out.weight"] = rnd(d, d, std=scale / math.sqrt(d))
            sd[f"{prefix}.{a}.out.bias"] = rnd(d, std=0.02)
            sd[f"{prefix}.{a}_ln.weight"] = 1.0 + rnd(d, std=0.05)
            sd[f"{prefix}.{a}_ln.bias"] = rnd(d, std=0.05)
        sd[f"{prefix}.mlp.0.weight"] = rnd(4 * d, d, std=1.0 / math.sqrt(d))
        sd[f"{prefix}.mlp.0.bias"] = rnd(4 * d, std=0.1)
        w2 = rnd(d, 4 * d, std=2.0 * res_scale / math.sqrt(4 * d))
        sd[f"{prefix}.mlp.2.weight"] = w2 - w2.mean(dim=1, keepdim=True)    # no response to the mean activation
        sd[f"{prefix}.mlp.2.bias"] = rnd(d, std=0.02)
        sd[f"{prefix}.mlp_ln.weight"] = 1.0 + rnd(d, std=0.05)
        sd[f"{prefix}.mlp_ln.bias"] = rnd(d, std=0.05)

    for i in range(dims.n_audio_layer):
        block(f"encoder.blocks.{i}", d_a, cross=False)
    sd["encoder.ln_post.weight"] = 1.0 + rnd(d_a, std=0.05)
    sd["encoder.ln_post.bias"] = rnd(d_a, std=0.05)

    eot, sot, n_lang, ts_begin = special_token_layout(dims.n_vocab)
    emb = rnd(dims.n_vocab, d_t, std=logit_spread / math.sqrt(d_t - 1))
    emb[:, 0] = 0.0
    emb[ts_begin:, 0] = ts_offset
    emb[eot, :] = 0.0
    emb[eot, 0] = eot_logit
    emb[eot + 1:ts_begin, 0] = -30.0        # language / task / control tokens are never produced as text
    emb[128:256, 0] = -30.0                 # lone UTF-8 continuation bytes of the synthetic vocabulary: a real model
                                            # never emits an undecodable tail (the reference raises on it, T.py:1471)
    sd["decoder.token_embedding.weight"] = emb
    sd["decoder.positional_embedding"] = rnd(dims.n_text_ctx, d_t, std=logit_spread / math.sqrt(d_t - 1))
    for i in range(dims.n_text_layer):
        block(f"decoder.blocks.{i}", d_t, cross=True)
    lnw = 1.0 + rnd(d_t, std=0.05)
    lnb = rnd(d_t, std=0.05)
    lnw[0] = 0.0
    lnb[0] = 1.0
    sd["decoder.ln.weight"] = lnw
    sd["decoder.ln.bias"] = lnb
    return sd
