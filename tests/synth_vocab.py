"""Small adversarial rank files: tokens over a tiny alphabet with RANDOM rank order, so merged tokens may rank
below their parts (tiktoken accepts any byte-string -> rank map).  This is what exercises the guard of the
batched merge rounds and every odd corner of the merge loop (test infrastructure)."""
import base64
import random


def make_rank_file(seed: int, alphabet=b"abc", n_extra=60, max_len=6) -> bytes:
    rng = random.Random(seed)
    toks = [bytes([i]) for i in range(256)]
    seen = set(toks)
    extra = []
    while len(extra) < n_extra:
        ln = rng.randint(2, max_len)
        t = bytes(rng.choice(alphabet) for _ in range(ln))
        if t not in seen:
            seen.add(t)
            extra.append(t)
    rng.shuffle(extra)
    toks += extra
    return b"\n".join(base64.b64encode(t) + b" " + str(i).encode() for i, t in enumerate(toks)) + b"\n"


def make_texts(seed: int, count: int, alphabet="abc", max_len=400):
    rng = random.Random(seed)
    out = []
    for _ in range(count):
        k = rng.randint(0, 3)
        n = rng.randint(1, max_len)
        if k == 0:
            s = "".join(rng.choice(alphabet) for _ in range(n))
        elif k == 1:
            s = rng.choice(alphabet) * n
        elif k == 2:
            per = "".join(rng.choice(alphabet) for _ in range(rng.randint(2, 5)))
            s = (per * (n // len(per) + 1))[:n]
        else:
            s = " ".join("".join(rng.choice(alphabet) for _ in range(rng.randint(1, 50))) for _ in range(rng.randint(1, 8)))
        out.append(s)
    return out
