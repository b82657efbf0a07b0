"""Tokenization prefix store (host side): the oracle restatement of prefixstore.LRUTokenStore against the
reference's own test cases and independent XXH64 vectors; the C++ host layer (kvidx_host.h) against the oracle."""
import numpy as np
import pytest

from helpers import golden
from kvidx import host
from oracle import kvoracle as ko

KATS = golden("prefix_store_kats.json")
TEXT = b"The capital of France is Paris"
TOKENS = [1, 2, 3, 4, 5, 6]
OFFSETS = [(0, 3), (4, 11), (12, 14), (15, 21), (22, 24), (25, 30)]


def _stores(cache_size=500000, block_size=4):
    return ko.LRUTokenStore(cache_size, block_size), host.PrefixStore(cache_size, block_size)


def test_xxh64_vectors_oracle_and_host():
    for k in KATS["xxh64"]:
        data = bytes.fromhex(k["hex"])
        assert ko.xxhash64(data, k["seed"]) == k["xxh64"]
        assert host.xxhash64(data, k["seed"]) == k["xxh64"]


@pytest.mark.parametrize("bs", [4, 8, 30])
def test_chained_block_keys_match_independent_vectors(bs):
    st = ko.LRUTokenStore(100, bs)
    prev, keys = 0, []
    for s in range(0, len(TEXT) - bs + 1, bs):
        prev = st._hash(prev, TEXT[s:s + bs]); keys.append(prev)
    assert keys == KATS["block_keys"][str(bs)]
    st.add_tokenization(TEXT, TOKENS, OFFSETS)
    assert sorted(st.cache.keys()) == sorted(keys)


@pytest.mark.parametrize("which", ["oracle", "host"])
def test_reference_cases(which):
    """lru_store_test.go:48-162."""
    new = (lambda c, b: ko.LRUTokenStore(c, b)) if which == "oracle" else (lambda c, b: host.PrefixStore(c, b))
    st = new(500000, 4)
    st.add_tokenization(TEXT, TOKENS, OFFSETS)
    assert st.find_longest_contained_tokens(b"The capital of F") == ([1, 2, 3], 1.0)                      # AddAndRetrieve
    for prompt, match, n_min in ((b"The capital of France is Marseille", b"The capital of France is", 4),      # PartialMismatch
                                 (b"The capital of Japan is Tokyo", b"The capital ", 2),
                                 (b"The capital of F", b"The capital of F", 3),
                                 (TEXT, b"The capital of France is Par", 5)):
        toks, ratio = st.find_longest_contained_tokens(prompt)
        assert set(toks) <= set(TOKENS) and len(toks) >= n_min and ratio == len(match) / len(prompt)
    prefix = b""
    for i, word in enumerate(TEXT.split(b" ")):                                                              # PrefixMatch
        prefix += (b" " if i else b"") + word
        toks, ratio = st.find_longest_contained_tokens(prefix)
        assert set(toks) <= set(TOKENS) and i <= len(toks) <= i + 1 and ratio >= (len(prefix) // 4) * 4 / len(prefix)
    st = new(2, 18)                                                                                           # LRUEviction
    texts = [b"abcdefghjiklmno", b"123456789011121314", b"pqrstuvwxyz,./';lp"]
    offs = [[(0, 5), (6, 10), (11, 15)], [(0, 6), (7, 12), (13, 18)], [(0, 6), (7, 12), (13, 18)]]
    for t, tk, of in zip(texts, ([1, 2, 3], [4, 5, 6], [7, 8, 9]), offs):
        st.add_tokenization(t, tk, of)
    assert st.find_longest_contained_tokens(texts[0])[0] == []
    assert st.find_longest_contained_tokens(texts[2])[0] == [7, 8, 9]
    assert st.find_longest_contained_tokens(b"") == ([], 0.0)
    st.add_tokenization(b"", [1], [(0, 1)]); st.add_tokenization(b"abcd", [], [])                             # no-ops (:96-98)


def test_host_store_vs_oracle_random_texts_and_eviction():
    """Differential: random UTF-8 texts sharing prefixes, random token boundaries, a cache small enough to evict; every
    find refreshes recency, so the two LRUs must stay in step call by call."""
    rng = np.random.default_rng(5)
    o, h = _stores(cache_size=40, block_size=16)
    alphabet = "abc défg—hi 😀xyz\n".encode()
    bases = [bytes(rng.choice(list(alphabet), size=int(rng.integers(1, 400))).astype(np.uint8)) for _ in range(12)]
    for step in range(600):
        base = bases[int(rng.integers(0, len(bases)))]
        text = base[: int(rng.integers(0, len(base) + 1))] + bytes(rng.choice(list(alphabet), size=int(rng.integers(0, 40))).astype(np.uint8))
        if rng.random() < 0.5:
            cuts = sorted(set(rng.integers(1, len(text) + 1, size=int(rng.integers(0, 30))).tolist())) if len(text) else []
            offsets, lo = [], 0
            for c in cuts:
                offsets.append((lo, c)); lo = c
            tokens = rng.integers(0, 100000, size=len(offsets)).tolist()
            o.add_tokenization(text, tokens, offsets)
            h.add_tokenization(text, tokens, offsets)
        else:
            assert h.find_longest_contained_tokens(text) == o.find_longest_contained_tokens(text), step
        assert len(h) == len(o.cache) <= 40


def test_process_task_logic():
    """tokenization.Pool.processTask (pool.go:209-225) on top of the store: the second, longer prompt is answered from the
    cache when >= 80 % of it is covered -- with fewer tokens than a fresh tokenization (the history dependence SURVEY 8(a)
    describes, which is why parity starts at token ids)."""
    calls = []

    def encode(p):
        calls.append(p)
        words, offs, pos = p.split(b" "), [], 0
        for wd in words:
            offs.append((pos, pos + len(wd))); pos += len(wd) + 1
        return list(range(1, len(words) + 1)), offs
    st = ko.LRUTokenStore(1000, 8)
    first = b"alpha beta gamma delta epsilon zeta eta theta iota kappa"
    assert ko.tokenize_with_prefix_store(st, first, encode) == list(range(1, 11)) and len(calls) == 1
    longer = first + b" lam"
    got = ko.tokenize_with_prefix_store(st, longer, encode)
    assert len(first) == 56 and len(calls) == 1 and got == list(range(1, 11))     # 56 of 60 bytes covered (93 %): no tokenizer call, and the
    assert len(encode(longer)[0]) == 11; calls.pop()                               # new last word is missing -- a fresh tokenization has 11 tokens
    assert ko.tokenize_with_prefix_store(st, b"something else entirely, nothing cached here", encode) and len(calls) == 2


def test_create_errors():
    with pytest.raises(Exception):
        host.PrefixStore(0, 16)
    with pytest.raises(Exception):
        host.PrefixStore(10, 0)
    with pytest.raises(ValueError):
        ko.LRU(0)
