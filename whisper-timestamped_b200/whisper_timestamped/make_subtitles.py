"""Subtitle files from a `.words.json` transcription — host-only mirror of the reference's
`whisper_timestamped/make_subtitles.py` (/root/reference/whisper_timestamped/make_subtitles.py:8-157).

Same public names and behaviour: `split_long_segments` (re-cuts segments longer than `max_length` characters at word
boundaries, preferring the last punctuation mark seen), `format_timestamp`, `write_vtt`, `write_srt`, `cli`.  Pinned by
the reference's own fixtures (tests/expected/split_subtitles/*, replayed by tests/test_subtitles.py).  No device code.
"""
import json
import string

# characters after which a cut is preferred: ASCII punctuation except the ones that live inside words, plus CJK marks
# (make_subtitles.py:6)
_punctuation = "".join(c for c in string.punctuation if c not in "-'") + "。，！？：”、…"

LANGUAGES_WITHOUT_SPACES = ("zh", "ja", "th", "lo", "my")


class _Line:
    """The subtitle line being filled, with the cut it would take if it overflowed."""

    def __init__(self, start):
        self.text = ""
        self.start = start
        self.cut = None            # (characters kept, end time of the word before the cut, start time of the word after)

    def forget_cut(self):
        self.cut = None


def split_long_segments(segments, max_length, use_space=True):
    """Segments whose text fits `max_length` pass through untouched; longer ones are re-cut into
    `{"text", "start", "end"}` pieces along their word list (make_subtitles.py:8-65)."""
    out = []
    for segment in segments:
        if len(segment["text"]) <= max_length:
            out.append(segment)
            continue
        timed = segment["words"]
        # the visible words come from the segment text (punctuation may have been stripped from the word entries)
        shown = segment["text"].split() if use_space else [w["text"] for w in timed]
        if len(shown) != len(timed):
            fallback = [w["text"] for w in timed]
            print(f"WARNING: {' '.join(shown)} != {' '.join(fallback)}")
            shown = fallback
        line = _Line(segment["start"])
        for k, (word, meta) in enumerate(zip(shown, timed)):
            before = line.text
            line.text = before + (" " if before and use_space else "") + word
            if len(line.text) > max_length and before:
                if line.cut is not None:
                    keep, end, next_start = line.cut
                    piece = {"text": line.text[:keep], "start": line.start, "end": end}
                    line.text, line.start = line.text[keep + 1:], next_start
                else:
                    piece = {"text": before, "start": line.start, "end": timed[k - 1]["end"]}
                    line.text, line.start = word, meta["start"]
                line.forget_cut()
                out.append(piece)
            if line.text and line.text[-1] in _punctuation:
                following = timed[k + 1]["start"] if k + 1 < len(timed) else None
                line.cut = (len(line.text), meta["end"], following)
        if line.text:
            out.append({"text": line.text, "start": line.start, "end": segment["end"]})
    return out


def format_timestamp(seconds: float, always_include_hours: bool = False, decimal_marker: str = "."):
    """[hh:]mm:ss.mmm (make_subtitles.py:67-81)."""
    assert seconds >= 0, "non-negative timestamp expected"
    total_ms = round(seconds * 1000.0)
    whole_seconds, ms = divmod(total_ms, 1000)
    whole_minutes, s = divmod(whole_seconds, 60)
    h, m = divmod(whole_minutes, 60)
    hours = f"{h:02d}:" if (always_include_hours or h > 0) else ""
    return f"{hours}{m:02d}:{s:02d}{decimal_marker}{ms:03d}"


def _cue_text(segment):
    return segment["text"].strip().replace("-->", "->")


def write_vtt(result, file):
    """WebVTT cues (make_subtitles.py:83-91)."""
    print("WEBVTT\n", file=file)
    for segment in result:
        print(f"{format_timestamp(segment['start'])} --> {format_timestamp(segment['end'])}\n{_cue_text(segment)}\n",
              file=file, flush=True)


def write_srt(result, file):
    """SubRip cues, numbered from 1 (make_subtitles.py:93-103)."""
    for number, segment in enumerate(result, start=1):
        begin = format_timestamp(segment["start"], always_include_hours=True, decimal_marker=",")
        end = format_timestamp(segment["end"], always_include_hours=True, decimal_marker=",")
        print(f"{number}\n{begin} --> {end}\n{_cue_text(segment)}\n", file=file, flush=True)


_WRITERS = {"srt": write_srt, "vtt": write_vtt}


def convert(input_file, output_files, max_length=200):
    """One `.words.json` file -> the given `.srt` / `.vtt` files."""
    with open(input_file, "r", encoding="utf-8") as f:
        transcript = json.load(f)
    segments = transcript["segments"]
    if max_length:
        segments = split_long_segments(segments, max_length,
                                       use_space=transcript["language"] not in LANGUAGES_WITHOUT_SPACES)
    for output in output_files:
        ext = output.rsplit(".", 1)[-1]
        if ext not in _WRITERS:
            raise RuntimeError(f"Unknown output format for {output}")
        with open(output, "w", encoding="utf-8") as f:
            _WRITERS[ext](segments, file=f)


def cli():
    """`whisper_timestamped_make_subtitles input output [--max_length N] [--format srt|vtt|all]`
    (make_subtitles.py:105-154): input / output may be files or folders."""
    import argparse
    import os

    formats = sorted(_WRITERS)
    parser = argparse.ArgumentParser(
        description="Convert .word.json transcription files (output of whisper_timestamped) to srt or vtt, "
                    "being able to cut long segments",
        formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("input", type=str, help="Input json file, or input folder")
    parser.add_argument("output", type=str, help="Output srt or vtt file, or output folder")
    parser.add_argument("--max_length", default=200, help="Maximum length of a segment in characters", type=int)
    parser.add_argument("--format", type=str, default="all", choices=formats + ["all"],
                        help="Output format (if the output is a folder, i.e. not a file with an explicit extension)")
    args = parser.parse_args()

    output_is_file = any(args.output.endswith(e) for e in formats)
    jobs = []
    if os.path.isdir(args.input) or not output_is_file:
        names = ([f for f in os.listdir(args.input) if f.endswith(".words.json")] if os.path.isdir(args.input)
                 else [os.path.basename(args.input)])
        wanted = formats if args.format == "all" else [args.format]
        for name in names:
            source = os.path.join(args.input, name) if os.path.isdir(args.input) else args.input
            stem = name[:-len(".words.json")]
            jobs.append((source, [os.path.join(args.output, f"{stem}.{e}") for e in wanted]))
        os.makedirs(args.output, exist_ok=True)
    else:
        jobs.append((args.input, [args.output]))
        os.makedirs(os.path.dirname(args.output) or ".", exist_ok=True)
    for source, outputs in jobs:
        convert(source, outputs, args.max_length)


if __name__ == "__main__":
    cli()
