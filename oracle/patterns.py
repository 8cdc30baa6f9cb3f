"""Pattern strings of the four pre-tokenizers, exactly as the stand-in oracle engine
(tiktoken 0.12.0) receives them.  TEST INFRASTRUCTURE (see bpe_oracle.c header).

Pattern ids match include/cfbpe.h (CFBPE_PATTERN_*) and oracle/bpe_oracle.c (PAT_*).
"""
PAT_CL100K, PAT_O200K, PAT_LLAMA3, PAT_TEKKEN = 0, 1, 2, 3
PATTERN_NAMES = {PAT_CL100K: "cl100k", PAT_O200K: "o200k", PAT_LLAMA3: "llama3", PAT_TEKKEN: "tekken"}

# tiktoken_ext/openai_public.py:89
CL100K = r"""'(?i:[sdmt]|ll|ve|re)|[^\r\n\p{L}\p{N}]?+\p{L}++|\p{N}{1,3}+| ?[^\s\p{L}\p{N}]++[\r\n]*+|\s++$|\s*[\r\n]|\s+(?!\S)|\s"""
# tiktoken_ext/openai_public.py:104-112
O200K = "|".join([
    r"""[^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]*[\p{Ll}\p{Lm}\p{Lo}\p{M}]+(?i:'s|'t|'re|'ve|'m|'ll|'d)?""",
    r"""[^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]+[\p{Ll}\p{Lm}\p{Lo}\p{M}]*(?i:'s|'t|'re|'ve|'m|'ll|'d)?""",
    r"""\p{N}{1,3}""",
    r""" ?[^\s\p{L}\p{N}]+[\r\n/]*""",
    r"""\s*[\r\n]+""",
    r"""\s+(?!\S)""",
    r"""\s+""",
])
# public Llama-3 tokenizer pattern
LLAMA3 = r"""(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+"""
# mistral_common/data/tekken_240911.json ["config"]["pattern"]
TEKKEN = r"""[^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]*[\p{Ll}\p{Lm}\p{Lo}\p{M}]+|[^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]+[\p{Ll}\p{Lm}\p{Lo}\p{M}]*|\p{N}| ?[^\s\p{L}\p{N}]+[\r\n/]*|\s*[\r\n]+|\s+(?!\S)|\s+"""

PATTERNS = {PAT_CL100K: CL100K, PAT_O200K: O200K, PAT_LLAMA3: LLAMA3, PAT_TEKKEN: TEKKEN}
