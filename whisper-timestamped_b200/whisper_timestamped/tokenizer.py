"""Whisper tokenizer for the drop-in package.

Special-token layout as in openai-whisper, which is what
/root/reference/whisper_timestamped/transcribe.py:1406-1426 obtains through
`whisper.tokenizer.get_tokenizer(multilingual, num_languages=, task=, language=)`.

Text vocabulary:
  * if the openai-whisper BPE rank files are available (env `WTS_WHISPER_ASSETS` pointing at a
    directory with `multilingual.tiktoken` / `gpt2.tiktoken`), they are loaded through tiktoken;
  * otherwise (this environment: no network, no assets) the deterministic SYNTHETIC vocabulary v1:
      id 0..255   -> the single byte `id`
      id >= 256   -> pseudo word piece derived from a multiplicative hash of the id
    decode() concatenates the byte strings and decodes UTF-8 with errors="replace".
"""
import base64
import os
from functools import lru_cache

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
TO_LANGUAGE_CODE = {name: code for code, name in LANGUAGES.items()}
TO_LANGUAGE_CODE.update({
    "burmese": "my", "valencian": "ca", "flemish": "nl", "haitian": "ht", "letzeburgesch": "lb",
    "pushto": "ps", "panjabi": "pa", "moldavian": "ro", "moldovan": "ro", "sinhalese": "si",
    "castilian": "es", "mandarin": "zh"})

_PUNCT = (".", ",", "?", "!", "...", ":")
_NON_SPEECH_SYMBOLS = '"#()*+/:;<=>@[\\]^_`{|}~'
# upstream's list continues with the CJK corner brackets; they are single tokens in multilingual.tiktoken (never in
# the byte-level synthetic vocabulary, whose suppress set therefore stays the ASCII one)
_NON_SPEECH_SYMBOLS_CJK = "\u300c\u300d\u300e\u300f"


def synthetic_piece(i: int) -> bytes:
    if i < 256:
        return bytes((i,))
    h = (i * 2654435761) & 0xFFFFFFFF
    if h % 53 == 0:
        return _PUNCT[(h >> 7) % len(_PUNCT)].encode()
    x, chars = h, []
    for _ in range(2 + (h >> 3) % 5):
        x = (x * 1103515245 + 12345) & 0x7FFFFFFF
        chars.append(chr(97 + (x >> 16) % 26))
    return ((" " if (h >> 11) % 5 < 3 else "") + "".join(chars)).encode()


class _SyntheticVocab:
    kind = "synthetic-v1"

    def __init__(self, n_text):
        self.n_text = n_text

    def piece(self, t):
        return synthetic_piece(t)

    def encode(self, text):
        return list(text.encode("utf-8"))


class _TiktokenVocab:
    kind = "tiktoken"

    def __init__(self, path, n_text):
        import tiktoken
        with open(path) as f:
            ranks = {base64.b64decode(tok): int(rank) for tok, rank in (line.split() for line in f if line)}
        assert len(ranks) == n_text, f"{path}: {len(ranks)} ranks, expected {n_text}"
        self.n_text = n_text
        self._by_id = {v: k for k, v in ranks.items()}
        self._enc = tiktoken.Encoding(
            name=os.path.basename(path), explicit_n_vocab=n_text,
            pat_str=r"""'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+""",
            mergeable_ranks=ranks, special_tokens={})

    def piece(self, t):
        return self._by_id[t]

    def encode(self, text):
        return self._enc.encode(text)


class Tokenizer:
    """Attributes mirror upstream's `whisper.tokenizer.Tokenizer` as used by the reference:
    sot, eot, sot_prev, sot_lm, no_speech, no_timestamps, timestamp_begin, transcribe, translate,
    sot_sequence, all_language_tokens, all_language_codes, non_speech_tokens, decode,
    decode_with_timestamps, encode, to_language_token."""

    def __init__(self, multilingual: bool, num_languages: int, language, task):
        self.multilingual = multilingual
        self.num_languages = num_languages
        n_text = 50257 if multilingual else 50256
        assets = os.environ.get("WTS_WHISPER_ASSETS", "")
        fname = os.path.join(assets, "multilingual.tiktoken" if multilingual else "gpt2.tiktoken")
        self.vocab = _TiktokenVocab(fname, n_text) if assets and os.path.exists(fname) else _SyntheticVocab(n_text)
        langs = list(LANGUAGES)[:num_languages]
        names = (["<|endoftext|>", "<|startoftranscript|>"] + [f"<|{c}|>" for c in langs]
                 + ["<|translate|>", "<|transcribe|>", "<|startoflm|>", "<|startofprev|>", "<|nospeech|>",
                    "<|notimestamps|>"] + [f"<|{i * 0.02:.2f}|>" for i in range(1501)])
        self.special_tokens = {name: n_text + i for i, name in enumerate(names)}
        self._special_by_id = {v: k for k, v in self.special_tokens.items()}
        self.n_vocab = n_text + len(names)
        sp = self.special_tokens
        self.eot, self.sot = sp["<|endoftext|>"], sp["<|startoftranscript|>"]
        self.translate, self.transcribe = sp["<|translate|>"], sp["<|transcribe|>"]
        self.sot_lm, self.sot_prev = sp["<|startoflm|>"], sp["<|startofprev|>"]
        self.no_speech, self.no_timestamps = sp["<|nospeech|>"], sp["<|notimestamps|>"]
        self.timestamp_begin = sp["<|0.00|>"]
        self.all_language_tokens = tuple(sp[f"<|{c}|>"] for c in langs)
        self.all_language_codes = tuple(langs)
        self.language, self.task = language, task
        seq = [self.sot]
        if language is not None:
            seq.append(self.to_language_token(language))
        if task is not None:
            seq.append(self.transcribe if task == "transcribe" else self.translate)
        self.sot_sequence = tuple(seq)

    def to_language_token(self, language):
        tok = self.special_tokens.get(f"<|{language}|>")
        if tok is None:
            raise KeyError(f"Language {language} not found in tokenizer.")
        return tok

    def encode(self, text):
        return self.vocab.encode(text)

    def _bytes(self, ids):
        nt = self.vocab.n_text
        return b"".join(self.vocab.piece(int(t)) if t < nt else self._special_by_id[int(t)].encode() for t in ids)

    def decode(self, ids, **kw):
        """Timestamp tokens are dropped (upstream behaviour)."""
        return self._bytes([t for t in ids if t < self.timestamp_begin]).decode("utf-8", errors="replace")

    def decode_with_timestamps(self, ids, **kw):
        return self._bytes(ids).decode("utf-8", errors="replace")

    @property
    def non_speech_tokens(self):
        if self.vocab.kind == "synthetic-v1":
            return tuple(sorted({ord(c) for c in _NON_SPEECH_SYMBOLS}))
        # upstream: symbols whose (single-token) encodings are suppressed
        symbols = list(_NON_SPEECH_SYMBOLS + _NON_SPEECH_SYMBOLS_CJK) + "<< >> <<< >>> -- --- -( -[ (' (\" (( )) ((( ))) [[ ]] {{ }} ♪♪ ♪♪♪".split()
        miscellaneous = set("♩♪♫♬♭♮♯")
        result = {self.vocab.encode(" -")[0], self.vocab.encode(" '")[0]}
        for symbol in symbols + list(miscellaneous):
            for tokens in (self.vocab.encode(symbol), self.vocab.encode(" " + symbol)):
                if len(tokens) == 1 or symbol in miscellaneous:
                    result.add(tokens[0])
        return tuple(sorted(result))


@lru_cache(maxsize=None)
def get_tokenizer(multilingual: bool, *, num_languages: int = 99, language=None, task=None) -> Tokenizer:
    if language is not None:
        language = language.lower()
        if language not in LANGUAGES:
            if language in TO_LANGUAGE_CODE:
                language = TO_LANGUAGE_CODE[language]
            else:
                raise ValueError(f"Unsupported language: {language}")
    if multilingual:
        language, task = language or "en", task or "transcribe"
    else:
        language, task = None, None
    return Tokenizer(multilingual, num_languages, language, task)
