"""Extracts the known-answer vectors of the reference's own `test_split_tokens`
(/root/reference/tests/test_transcribe.py:722-902: token ids -> words / pieces / ids for English, French, Arabic,
a non-timestamp special token and the English-only vocabulary) into tests/golden/split_tokens_vectors.json.
Runs only in the build container (the reference does not exist on the GPU box); the JSON is committed.

The vectors need the real Whisper vocabulary to run as written; tests/test_split_tokens_vectors.py replays them
through a stub tokenizer whose byte pieces are rebuilt from the expected pieces themselves.
"""
import ast
import json
import os

SRC = "/root/reference/tests/test_transcribe.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    tree = ast.parse(open(SRC).read())
    fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "test_split_tokens")
    ns = {"te": "", "_dot": " ."}          # whisper >= 20230314 (current upstream)
    vectors, multilingual, tokens, line = [], True, None, None
    for node in fn.body:
        if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name):
            name = node.targets[0].id
            if name == "tokens":
                tokens, line = ast.literal_eval(node.value), node.lineno
            elif name == "tokenizer":
                multilingual = bool(ast.literal_eval(node.value.args[0]))
        elif isinstance(node, ast.Expr) and isinstance(node.value, ast.Call) and \
                getattr(node.value.func, "attr", "") == "assertEqual":
            expected = eval(compile(ast.Expression(node.value.args[1]), SRC, "eval"), dict(ns))
            words, pieces, ids = expected
            vectors.append({"source_line": line, "multilingual": multilingual, "tokens": tokens,
                            "words": list(words), "pieces": [list(p) for p in pieces], "ids": [list(i) for i in ids]})
    with open(os.path.join(HERE, "split_tokens_vectors.json"), "w") as f:
        json.dump({"source": SRC + ":722-902", "vectors": vectors}, f, indent=1, ensure_ascii=False)
    print(len(vectors), "vectors")


if __name__ == "__main__":
    main()
