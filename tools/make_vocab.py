#!/usr/bin/env python3
"""Convert Mistral Tekken (tekken_240911.json, Apache-2.0, shipped inside the
`mistral_common` wheel of this image) to the ".tiktoken" rank-file format
(base64 token, space, decimal rank per line: tiktoken/load.py:160-172).

Tekken is the only real BPE vocabulary on this box: cl100k_base,
o200k_base and the Llama-3 rank files cannot be downloaded.  The committed file
`vocabs/tekken_240911.tiktoken` holds all 150 000 ranks; a prefix of a BPE rank
file is itself a valid BPE vocabulary, so the benchmark's "cl100k-size" (100 256),
"llama3-size" (128 000), default Tekken (130 072) and "o200k-slot" (150 000)
vocabularies are prefixes of this one file (see cfbpe/vocabs.py).
"""
import base64, hashlib, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/opt/prime-rl/.venv/lib/python3.12/site-packages/mistral_common/data/tekken_240911.json"

def main():
    import mistral_common
    src = os.path.join(os.path.dirname(mistral_common.__file__), "data", "tekken_240911.json")
    d = json.load(open(src))
    out = os.path.join(ROOT, "vocabs", "tekken_240911.tiktoken")
    with open(out, "wb") as f:
        for i, e in enumerate(d["vocab"]):
            assert e["rank"] == i
            tb = base64.b64decode(e["token_bytes"])
            f.write(base64.b64encode(tb) + b" " + str(i).encode() + b"\n")
    data = open(out, "rb").read()
    meta = {"source": "mistral_common/data/tekken_240911.json", "license": "Apache-2.0",
            "pattern": d["config"]["pattern"], "n_ranks": len(d["vocab"]),
            "sha256": hashlib.sha256(data).hexdigest()}
    json.dump(meta, open(os.path.join(ROOT, "vocabs", "tekken_240911.meta.json"), "w"), indent=1)
    print(meta)

if __name__ == "__main__":
    sys.exit(main())
