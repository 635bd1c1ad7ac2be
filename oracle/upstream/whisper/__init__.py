"""ORACLE stand-in for the `whisper` package (openai-whisper) — see oracle/upstream/README.md."""
from . import audio, decoding, model, tokenizer, utils  # noqa: F401
from .audio import load_audio, log_mel_spectrogram, pad_or_trim  # noqa: F401
from .decoding import DecodingOptions, DecodingResult, decode, detect_language  # noqa: F401
from .model import ModelDimensions, Whisper  # noqa: F401
from .transcribe import transcribe  # noqa: F401

__version__ = "20240930"

_MODELS = {}
normalizers = None


def available_models():
    return list(_MODELS.keys())


def _download(url, root, in_memory):
    raise RuntimeError("oracle whisper stand-in: no network, nothing to download")


def load_model(name, device=None, download_root=None, in_memory=False):
    """Upstream: name or path of a checkpoint {"dims":..., "model_state_dict":...}."""
    import os
    import torch
    if not os.path.isfile(name):
        raise RuntimeError(f"Model {name} not found; available models = {available_models()}")
    ckpt = torch.load(name, map_location="cpu")
    m = Whisper(ModelDimensions(**ckpt["dims"]))
    m.load_state_dict(ckpt["model_state_dict"])
    return m.to(device or "cpu")
