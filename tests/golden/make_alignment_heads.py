"""Copies the base85 alignment-head masks (DATA, not code) out of the reference
(/root/reference/whisper_timestamped/transcribe.py:2343-2357, `_ALIGNMENT_HEADS`) into
tests/golden/alignment_heads_b85.json, so that tests/test_split_tokens_vectors.py can tie the product's literal
(layer, head) table (model_zoo.ALIGNMENT_HEADS) to the reference's masks on any box."""
import ast
import json
import os

SRC = "/root/reference/whisper_timestamped/transcribe.py"
HERE = os.path.dirname(os.path.abspath(__file__))

tree = ast.parse(open(SRC).read())
node = next(n for n in tree.body if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") == "_ALIGNMENT_HEADS")
table = {k: v.decode("ascii") for k, v in ast.literal_eval(node.value).items()}
json.dump({"source": f"{SRC}:{node.lineno}-{node.end_lineno}", "masks": table},
          open(os.path.join(HERE, "alignment_heads_b85.json"), "w"), indent=1)
print(len(table), "masks from lines", node.lineno, node.end_lineno)
