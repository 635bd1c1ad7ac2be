"""Independent pin of the oracle's restatement of openai-whisper (oracle/upstream/whisper, written from the
published algorithm because the real wheel cannot be installed here — SURVEY.md §8c).

`transformers`' Whisper is a third-party implementation of the same architecture that IS installed in this image.
The same synthetic state dict is loaded into both (key names related by the reference's own rename table
`hf_to_whisper_states`, /root/reference/whisper_timestamped/transcribe.py:2876-2907, inverted here) and the
following must agree in float32:
  * encoder output                      <= 1e-4   (the CUDA path is then held within 1e-3 of the oracle)
  * decoder logits (teacher-forced)     <= 1e-4 relative to the logit scale
  * cross-attention: softmax of the oracle's pre-softmax `qk` (what the reference calls attention weights,
    T.py:783-793) vs HF's attention probabilities (the reference itself goes back and forth with `.log()` for HF
    models, T.py:1111-1114)                                   <= 1e-5 absolute, log-values <= 1e-4 where p > 1e-6
  * log-mel: `whisper.log_mel_spectrogram` vs `WhisperFeatureExtractor`  <= 2e-5 for 80 and 128 mel bands
NOT used for DTW parity: HF's `_dynamic_time_warping` breaks ties differently.
"""
import os
import re
import sys

import numpy as np
import pytest
import torch

from whisper_timestamped import model_zoo as zoo
from whisper_timestamped.synthetic_audio import synthetic_speech

transformers = pytest.importorskip("transformers")
import oracle_engine  # noqa: E402,F401  (puts oracle/upstream on sys.path: `import whisper` below is the oracle stand-in)

SMALL128 = zoo.ModelDimensions(128, 1500, 128, 2, 2, 51866, 448, 128, 2, 2)   # large-v3's mel / vocabulary layout, 2 layers
CASES = {"tiny": zoo.DIMS["tiny"], "mel128": SMALL128}


def hf_key_to_whisper(k):
    """The reference's rename table (T.py:2888-2906), applied the way the reference applies it."""
    for a, b in (('.layers.', '.blocks.'), ('.self_attn.', '.attn.'), ('.q_proj.', '.query.'), ('.k_proj.', '.key.'),
                 ('.v_proj.', '.value.'), ('.out_proj.', '.out.'), ('.fc1.', '.mlp.0.'), ('.fc2.', '.mlp.2.'),
                 ('.encoder_attn.', '.cross_attn.'), ('.cross_attn.ln.', '.cross_attn_ln.'),
                 ('.embed_positions.weight', '.positional_embedding'), ('.embed_tokens.', '.token_embedding.'),
                 ('model.', ''), ('attn.layer_norm.', 'attn_ln.'), ('.final_layer_norm.', '.mlp_ln.'),
                 ('encoder.layer_norm.', 'encoder.ln_post.'), ('decoder.layer_norm.', 'decoder.ln.')):
        k = re.sub(a, b, k)
    return k


def build_pair(dims, **kw):
    from oracle_engine import build_oracle_model
    sd = zoo.synthetic_state_dict(dims, seed=1234, **kw)
    heads = [(l, h) for l in range(dims.n_text_layer) for h in range(dims.n_text_head)]
    om = build_oracle_model(dims, sd, heads)
    cfg = transformers.WhisperConfig(
        vocab_size=dims.n_vocab, num_mel_bins=dims.n_mels, d_model=dims.n_audio_state,
        encoder_layers=dims.n_audio_layer, encoder_attention_heads=dims.n_audio_head,
        decoder_layers=dims.n_text_layer, decoder_attention_heads=dims.n_text_head,
        encoder_ffn_dim=4 * dims.n_audio_state, decoder_ffn_dim=4 * dims.n_text_state,
        max_source_positions=dims.n_audio_ctx, max_target_positions=dims.n_text_ctx, activation_function="gelu",
        dropout=0.0, attention_dropout=0.0, activation_dropout=0.0, attn_implementation="eager",
        pad_token_id=0, bos_token_id=1, eos_token_id=2, decoder_start_token_id=1)
    hf = transformers.WhisperForConditionalGeneration(cfg).float().eval()
    new = {}
    for k, v in hf.state_dict().items():
        if k == "proj_out.weight":
            new[k] = sd["decoder.token_embedding.weight"].clone()            # tied output embedding
            continue
        wk = hf_key_to_whisper(k)
        assert wk in sd, (k, wk)
        assert tuple(sd[wk].shape) == tuple(v.shape), (k, sd[wk].shape, v.shape)
        new[k] = sd[wk].clone().float()
    hf.load_state_dict(new)
    return om, hf


@pytest.mark.parametrize("name", list(CASES))
def test_model_forward_matches_transformers(name):
    import whisper
    from whisper.model import disable_sdpa
    dims = CASES[name]
    torch.manual_seed(0)
    om, hf = build_pair(dims)
    audio = torch.from_numpy(synthetic_speech(30.0, seed=5))
    mel = whisper.log_mel_spectrogram(audio, dims.n_mels)
    mel = whisper.pad_or_trim(mel, 3000)[None]
    eot, sot, n_lang, ts0 = zoo.special_token_layout(dims.n_vocab)
    tokens = torch.tensor([[sot, sot + 1, ts0 - 5, ts0 + 10, 1000, 2000, 345, 11, ts0 + 60, ts0 + 60, 777]])
    captured = []
    hooks = [blk.cross_attn.register_forward_hook(lambda m, i, o: captured.append(o[-1])) for blk in om.decoder.blocks]
    with torch.no_grad(), disable_sdpa():
        xa = om.encoder(mel)
        logits = om.decoder(tokens, xa)
    for h in hooks:
        h.remove()
    with torch.no_grad():
        out = hf(input_features=mel, decoder_input_ids=tokens, output_attentions=True, use_cache=False)
    enc_err = (out.encoder_last_hidden_state - xa).abs().max().item()
    assert enc_err <= 1e-4, enc_err
    scale = max(1.0, logits.abs().max().item())
    lg_err = (out.logits - logits).abs().max().item()
    assert lg_err <= 1e-4 * scale, (lg_err, scale)
    assert len(captured) == dims.n_text_layer
    for l, qk in enumerate(captured):
        p_ref = torch.softmax(qk.float(), dim=-1)
        p_hf = out.cross_attentions[l]
        assert (p_ref - p_hf).abs().max().item() <= 1e-5, l
        big = p_ref > 1e-6
        assert (p_ref[big].log() - p_hf[big].log()).abs().max().item() <= 1e-4, l


@pytest.mark.parametrize("n_mels", [80, 128])
def test_log_mel_matches_feature_extractor(n_mels):
    import whisper
    fe = transformers.WhisperFeatureExtractor(feature_size=n_mels)
    for dur, seed in ((30.0, 3), (12.7, 4)):
        audio = synthetic_speech(dur, seed=seed)
        feats = fe(audio, sampling_rate=16000, return_tensors="np")["input_features"][0]     # [n_mels, 3000]
        padded = whisper.pad_or_trim(torch.from_numpy(audio), 480000)
        ref = whisper.log_mel_spectrogram(padded, n_mels).numpy()
        assert feats.shape == ref.shape
        err = np.max(np.abs(feats - ref))
        assert err <= 2e-5, (n_mels, dur, err)
    # the filterbank the product multiplies with == the one transformers derives (upstream ships it as an asset)
    from whisper_timestamped.model import mel_filterbank
    from transformers.audio_utils import mel_filter_bank
    fb = mel_filter_bank(num_frequency_bins=201, num_mel_filters=n_mels, min_frequency=0.0, max_frequency=8000.0,
                         sampling_rate=16000, norm="slaney", mel_scale="slaney").T
    assert np.max(np.abs(fb - mel_filterbank(n_mels))) <= 1e-7
