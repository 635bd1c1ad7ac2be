"""Copies the reference's own fixtures for its host-only text writers into tests/golden/subtitles/ (build container
only; the copies are committed because /root/reference does not exist on the GPU box):

* inputs of `TestMakeSubtitles.test_make_subtitles` (/root/reference/tests/test_transcribe.py:619-650):
  tests/data/{smartphone.mp3,no_punctuations.mp3,yes_punctuations.mp3}.words.json, and the expected
  tests/expected/split_subtitles/*_{6,20,50}.{srt,vtt};
* result dicts + the csv / tsv files the reference's command line wrote from them:
  tests/expected/punctuations_{yes,no}/{punctuations.mp3,bonjour.wav}.{words.json,csv,tsv,words.csv,words.tsv}.

These are data fixtures of the reference's test-suite, not source code.
"""
import os
import shutil

REF = "/root/reference/tests"
HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "subtitles")


def main():
    os.makedirs(HERE, exist_ok=True)
    n = 0
    for name in ("smartphone.mp3", "no_punctuations.mp3", "yes_punctuations.mp3"):
        shutil.copyfile(f"{REF}/data/{name}.words.json", f"{HERE}/in_{name}.words.json")
        n += 1
    for name in sorted(os.listdir(f"{REF}/expected/split_subtitles")):
        shutil.copyfile(f"{REF}/expected/split_subtitles/{name}", f"{HERE}/split_{name}")
        n += 1
    for folder in ("punctuations_yes", "punctuations_no"):
        for stem in ("punctuations.mp3", "bonjour.wav"):
            for ext in ("words.json", "csv", "tsv", "words.csv", "words.tsv"):
                shutil.copyfile(f"{REF}/expected/{folder}/{stem}.{ext}", f"{HERE}/{folder}_{stem}.{ext}")
                n += 1
    for f in os.listdir(HERE):
        os.chmod(os.path.join(HERE, f), 0o644)
    print(n, "files")


if __name__ == "__main__":
    main()
