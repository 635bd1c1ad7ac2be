"""ORACLE restatement of openai-whisper `whisper.tokenizer` (see oracle/upstream/README.md).

The special-token layout follows upstream exactly; the BPE vocabulary itself ships inside the
openai-whisper wheel (absent here), so text tokens use the deterministic SYNTHETIC vocabulary v1:
  id 0..255      -> the single byte `id`
  id 256..n-1    -> a pseudo word piece derived from a multiplicative hash of the id
                    (2-6 lowercase letters; ~60 % start with a space; ~2 % are punctuation marks)
`decode` concatenates bytes and decodes UTF-8 with errors="replace", like tiktoken.
"""
from dataclasses import dataclass, field
from functools import cached_property, lru_cache
from typing import Dict, List, Optional, Tuple

LANGUAGES = {
    "en": "english", "zh": "chinese", "de": "german", "es": "spanish", "ru": "russian", "ko": "korean",
    "fr": "french", "ja": "japanese", "pt": "portuguese", "tr": "turkish", "pl": "polish", "ca": "catalan",
    "nl": "dutch", "ar": "arabic", "sv": "swedish", "it": "italian", "id": "indonesian", "hi": "hindi",
    "fi": "finnish", "vi": "vietnamese", "he": "hebrew", "uk": "ukrainian", "el": "greek", "ms": "malay",
    "cs": "czech", "ro": "romanian", "da": "danish", "hu": "hungarian", "ta": "tamil", "no": "norwegian",
    "th": "thai", "ur": "urdu", "hr": "croatian", "bg": "bulgarian", "lt": "lithuanian", "la": "latin",
    "mi": "maori", "ml": "malayalam", "cy": "welsh", "sk": "slovak", "te": "telugu", "fa": "persian",
    "lv": "latvian", "bn": "bengali", "sr": "serbian", "az": "azerbaijani", "sl": "slovenian",
    "kn": "kannada", "et": "estonian", "mk": "macedonian", "br": "breton", "eu": "basque",
    "is": "icelandic", "hy": "armenian", "ne": "nepali", "mn": "mongolian", "bs": "bosnian",
    "kk": "kazakh", "sq": "albanian", "sw": "swahili", "gl": "galician", "mr": "marathi",
    "pa": "punjabi", "si": "sinhala", "km": "khmer", "sn": "shona", "yo": "yoruba", "so": "somali",
    "af": "afrikaans", "oc": "occitan", "ka": "georgian", "be": "belarusian", "tg": "tajik",
    "sd": "sindhi", "gu": "gujarati", "am": "amharic", "yi": "yiddish", "lo": "lao", "uz": "uzbek",
    "fo": "faroese", "ht": "haitian creole", "ps": "pashto", "tk": "turkmen", "nn": "nynorsk",
    "mt": "maltese", "sa": "sanskrit", "lb": "luxembourgish", "my": "myanmar", "bo": "tibetan",
    "tl": "tagalog", "mg": "malagasy", "as": "assamese", "tt": "tatar", "haw": "hawaiian",
    "ln": "lingala", "ha": "hausa", "ba": "bashkir", "jw": "javanese", "su": "sundanese",
    "yue": "cantonese",
}

TO_LANGUAGE_CODE = {
    **{language: code for code, language in LANGUAGES.items()},
    "burmese": "my", "valencian": "ca", "flemish": "nl", "haitian": "ht", "letzeburgesch": "lb",
    "pushto": "ps", "panjabi": "pa", "moldavian": "ro", "moldovan": "ro", "sinhalese": "si",
    "castilian": "es", "mandarin": "zh",
}

_PUNCT_PIECES = [".", ",", "?", "!", "...", ":"]


def synthetic_piece(i: int) -> bytes:
    """Bytes of text token `i` in the synthetic vocabulary v1."""
    if i < 256:
        return bytes([i])
    h = (i * 2654435761) & 0xFFFFFFFF
    if h % 53 == 0:
        return _PUNCT_PIECES[(h >> 7) % len(_PUNCT_PIECES)].encode()
    n = 2 + (h >> 3) % 5
    x = h
    letters = []
    for _ in range(n):
        x = (x * 1103515245 + 12345) & 0x7FFFFFFF
        letters.append(chr(ord("a") + (x >> 16) % 26))
    lead = " " if (h >> 11) % 5 < 3 else ""
    return (lead + "".join(letters)).encode()


class SyntheticEncoding:
    """Minimal tiktoken.Encoding look-alike."""

    def __init__(self, name: str, n_text: int, special_tokens: Dict[str, int]):
        self.name = name
        self.n_text = n_text
        self.special_tokens = special_tokens
        self._special_by_id = {v: k for k, v in special_tokens.items()}
        self.n_vocab = n_text + len(special_tokens)
        self.eot_token = special_tokens["<|endoftext|>"]
        self.special_tokens_set = set(special_tokens.keys())

    def encode(self, text: str, **kwargs) -> List[int]:
        return list(text.encode("utf-8"))

    def encode_single_token(self, text: str) -> int:
        if text in self.special_tokens:
            return self.special_tokens[text]
        b = text.encode("utf-8")
        if len(b) != 1:
            raise KeyError(text)
        return b[0]

    def decode_bytes(self, tokens) -> bytes:
        out = []
        for t in tokens:
            t = int(t)
            if t < self.n_text:
                out.append(synthetic_piece(t))
            else:
                out.append(self._special_by_id[t].encode())
        return b"".join(out)

    def decode(self, tokens, errors: str = "replace") -> str:
        return self.decode_bytes(tokens).decode("utf-8", errors=errors)


@dataclass
class Tokenizer:
    encoding: SyntheticEncoding
    num_languages: int
    language: Optional[str] = None
    task: Optional[str] = None
    sot_sequence: Tuple[int] = ()
    special_tokens: Dict[str, int] = field(default_factory=dict)

    def __post_init__(self):
        for special, tid in self.encoding.special_tokens.items():
            self.special_tokens[special] = tid
        sot = self.special_tokens["<|startoftranscript|>"]
        translate = self.special_tokens["<|translate|>"]
        transcribe = self.special_tokens["<|transcribe|>"]
        langs = tuple(LANGUAGES.keys())[: self.num_languages]
        sot_sequence = [sot]
        if self.language is not None:
            sot_sequence.append(sot + 1 + langs.index(self.language))
        if self.task is not None:
            sot_sequence.append(transcribe if self.task == "transcribe" else translate)
        self.sot_sequence = tuple(sot_sequence)

    def encode(self, text, **kwargs):
        return self.encoding.encode(text, **kwargs)

    def decode(self, token_ids, **kwargs) -> str:
        token_ids = [t for t in token_ids if t < self.timestamp_begin]
        return self.encoding.decode(token_ids, **kwargs)

    def decode_with_timestamps(self, token_ids, **kwargs) -> str:
        return self.encoding.decode(token_ids, **kwargs)

    @cached_property
    def eot(self) -> int:
        return self.encoding.eot_token

    @cached_property
    def transcribe(self) -> int:
        return self.special_tokens["<|transcribe|>"]

    @cached_property
    def translate(self) -> int:
        return self.special_tokens["<|translate|>"]

    @cached_property
    def sot(self) -> int:
        return self.special_tokens["<|startoftranscript|>"]

    @cached_property
    def sot_lm(self) -> int:
        return self.special_tokens["<|startoflm|>"]

    @cached_property
    def sot_prev(self) -> int:
        return self.special_tokens["<|startofprev|>"]

    @cached_property
    def no_speech(self) -> int:
        return self.special_tokens["<|nospeech|>"]

    @cached_property
    def no_timestamps(self) -> int:
        return self.special_tokens["<|notimestamps|>"]

    @cached_property
    def timestamp_begin(self) -> int:
        return self.special_tokens["<|0.00|>"]

    @cached_property
    def language_token(self) -> int:
        if self.language is None:
            raise ValueError("This tokenizer does not have language token configured")
        return self.to_language_token(self.language)

    def to_language_token(self, language):
        if token := self.special_tokens.get(f"<|{language}|>", None):
            return token
        raise KeyError(f"Language {language} not found in tokenizer.")

    @cached_property
    def all_language_tokens(self) -> Tuple[int]:
        result = []
        for token, token_id in self.special_tokens.items():
            if token.strip("<|>") in LANGUAGES:
                result.append(token_id)
        return tuple(result)[: self.num_languages]

    @cached_property
    def all_language_codes(self) -> Tuple[str]:
        return tuple(self.decode([_l]).strip("<|>") for _l in self.all_language_tokens)

    @cached_property
    def sot_sequence_including_notimestamps(self) -> Tuple[int]:
        return tuple(list(self.sot_sequence) + [self.no_timestamps])

    @cached_property
    def non_speech_tokens(self) -> Tuple[int]:
        """Upstream: ids of symbol strings that are not speech.  Synthetic vocabulary: the byte
        tokens of the ASCII symbols upstream lists (single characters only)."""
        symbols = list('"#()*+/:;<=>@[\\]^_`{|}~\u300c\u300d\u300e\u300f')
        # upstream keeps a symbol only when it is ONE token: the CJK corner brackets are, in multilingual.tiktoken;
        # in the byte-level synthetic vocabulary they are three bytes each and drop out
        return tuple(sorted({self.encoding.encode(s)[0] for s in symbols if len(self.encoding.encode(s)) == 1}))


@lru_cache(maxsize=None)
def get_encoding(name: str = "gpt2", num_languages: int = 99):
    n_text = 50256 if name == "gpt2" else 50257
    specials = [
        "<|endoftext|>", "<|startoftranscript|>",
        *[f"<|{lang}|>" for lang in list(LANGUAGES.keys())[:num_languages]],
        "<|translate|>", "<|transcribe|>", "<|startoflm|>", "<|startofprev|>", "<|nospeech|>",
        "<|notimestamps|>", *[f"<|{i * 0.02:.2f}|>" for i in range(1501)],
    ]
    special_tokens = {}
    n = n_text
    for token in specials:
        special_tokens[token] = n
        n += 1
    return SyntheticEncoding(name, n_text, special_tokens)


@lru_cache(maxsize=None)
def get_tokenizer(multilingual: bool, *, num_languages: int = 99, language: Optional[str] = None,
                  task: Optional[str] = None) -> Tokenizer:
    if language is not None:
        language = language.lower()
        if language not in LANGUAGES:
            if language in TO_LANGUAGE_CODE:
                language = TO_LANGUAGE_CODE[language]
            else:
                raise ValueError(f"Unsupported language: {language}")
    if multilingual:
        encoding_name = "multilingual"
        language = language or "en"
        task = task or "transcribe"
    else:
        encoding_name = "gpt2"
        language = None
        task = None
    encoding = get_encoding(name=encoding_name, num_languages=num_languages)
    return Tokenizer(encoding=encoding, num_languages=num_languages, language=language, task=task)
