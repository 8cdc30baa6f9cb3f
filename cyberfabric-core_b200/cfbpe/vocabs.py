"""Vocabulary registry: the "model-registry vocab loader" side of the path.

The reference's `Model` entity has no tokenizer field (modules/model-registry/docs/PRD.md:196-209);
the mapping canonical model id `{provider_slug}::{provider_model_id}` (PRD.md:197) -> VocabSpec
lives here until the registry grows one.

Real OpenAI / Meta rank files are NOT on this box (no network).  A VocabSpec
therefore resolves in this order:
  1. a real rank file found in $CFBPE_VOCAB_DIR or vocabs/ whose sha256 matches the published one;
  2. ONLY when the caller opts in (allow_stand_in=True / CFBPE_ALLOW_STAND_IN=1: benchmarks and tests): the committed
     Mistral Tekken rank file truncated to the same vocabulary size, used with the requested pattern -- labelled
     `stand_in=True` so that every benchmark line and the plugin instance's properties say so.  Otherwise: VocabUnavailable.

Tekken ids here are raw ranks 0..n-1 of the rank file; mistral_common shifts them by its 1000 reserved special-token ids
(`id = rank + 1000`).  A gateway that serves Mistral models adds that offset above this layer.
"""
import hashlib
import os
from dataclasses import dataclass

from . import _native as N

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
VOCAB_DIR = os.path.join(_ROOT, "vocabs")
TEKKEN_FILE = os.path.join(VOCAB_DIR, "tekken_240911.tiktoken")

# published sha256 of the OpenAI files (tiktoken_ext/openai_public.py)
KNOWN_SHA256 = {
    "cl100k_base.tiktoken": "223921b76ee99bde995b7ff738513eef100fb51d18c93597a113bcffe865b2a7",
    "o200k_base.tiktoken": "446a9538cb6c348e3516120d7c08b09f57c36495e2acfffe59a5bf8b0cfb1a2d",
}


@dataclass(frozen=True)
class VocabSpec:
    name: str
    pattern: str          # cl100k | o200k | llama3 | tekken
    n_ranks: int          # vocabulary size of the real thing
    real_file: str = ""   # file name of the real rank file, if one exists publicly
    fmt: int = N.FORMAT_TIKTOKEN


SPECS = {
    "cl100k_base": VocabSpec("cl100k_base", "cl100k", 100256, "cl100k_base.tiktoken"),
    "o200k_base": VocabSpec("o200k_base", "o200k", 199998, "o200k_base.tiktoken"),
    "llama3": VocabSpec("llama3", "llama3", 128000, "llama3.tiktoken"),
    "tekken": VocabSpec("tekken", "tekken", 130072, "tekken_240911.tiktoken"),
}

# canonical model id -> vocab (provider_slug::provider_model_id, modules/model-registry/docs/PRD.md:197)
MODEL_VOCABS = {
    "openai::gpt-4": "cl100k_base", "openai::gpt-3.5-turbo": "cl100k_base", "openai::text-embedding-3-small": "cl100k_base",
    "openai::gpt-4o": "o200k_base", "openai::gpt-4o-mini": "o200k_base", "openai::o1": "o200k_base",
    "meta::llama-3-8b-instruct": "llama3", "meta::llama-3.1-70b-instruct": "llama3",
    "mistral::mistral-nemo": "tekken", "mistral::pixtral-12b": "tekken",
}


@dataclass
class ResolvedVocab:
    spec: VocabSpec
    file_bytes: bytes
    max_ranks: int
    stand_in: bool
    label: str
    sha256: str

    @property
    def pattern_id(self):
        return N.PATTERN_IDS[self.spec.pattern]


def _read(path):
    with open(path, "rb") as f:
        return f.read()


class VocabUnavailable(LookupError):
    """the vocabulary is unknown, or its real rank file is not on this host and stand-ins were not asked for"""


def resolve(name: str, allow_stand_in: bool = False) -> ResolvedVocab:
    """allow_stand_in: serve a vocabulary whose real rank file is absent from the truncated Tekken file (benchmarks and
    tests only -- its token counts are NOT the real model's).  Off by default: a gateway must not bill from a stand-in."""
    if name not in SPECS:
        raise VocabUnavailable("unknown vocabulary %r" % name)
    spec = SPECS[name]
    if name == "tekken":
        data = _read(TEKKEN_FILE)
        return ResolvedVocab(spec, data, spec.n_ranks, False, "tekken_240911[:130072]", hashlib.sha256(data).hexdigest())
    for d in filter(None, [os.environ.get("CFBPE_VOCAB_DIR"), VOCAB_DIR]):
        p = os.path.join(d, spec.real_file)
        if os.path.exists(p):
            data = _read(p)
            sha = hashlib.sha256(data).hexdigest()
            want = KNOWN_SHA256.get(spec.real_file)
            if want and sha != want:
                raise ValueError("%s: sha256 %s does not match the published %s" % (p, sha, want))
            return ResolvedVocab(spec, data, 0, False, spec.real_file, sha)
    if not (allow_stand_in or os.environ.get("CFBPE_ALLOW_STAND_IN") == "1"):
        raise VocabUnavailable("%s: %s is not in $CFBPE_VOCAB_DIR or %s (stand-ins are opt-in: allow_stand_in=True)"
                               % (name, spec.real_file, VOCAB_DIR))
    data = _read(TEKKEN_FILE)
    n = min(spec.n_ranks, 150000)
    return ResolvedVocab(spec, data, n, True, "STAND-IN tekken_240911[:%d] + %s pattern" % (n, spec.pattern),
                         hashlib.sha256(data).hexdigest())


def for_model(canonical_id: str) -> str:
    """vocab name for a model-registry canonical id; VocabUnavailable if the model is unknown"""
    try:
        return MODEL_VOCABS[canonical_id]
    except KeyError:
        raise VocabUnavailable("no vocabulary is registered for model %r" % canonical_id) from None
