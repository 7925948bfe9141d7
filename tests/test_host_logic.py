"""Host-side logic of the C++ mirror that needs no device: tokenizer, dictionary bookkeeping, validation errors,
shard planning and the exchange-record layout.  Mirrors the reference's tests where one exists."""
import numpy as np
import pytest

from oracle.tokenizer import tokenize as oracle_tokenize
from trueno_rag_b200 import shard


@pytest.fixture(scope="module")
def api(built_lib):
    from trueno_rag_b200 import api as a
    return a


def test_bm25_index_new(api):                                   # src/index.rs:481-488
    ix = api.BM25Index()
    assert len(ix) == 0 and ix.is_empty()
    assert abs(ix.k1 - 1.2) < 0.01 and abs(ix.b - 0.75) < 0.01


def test_bm25_index_with_params(api):                           # :490-495
    ix = api.BM25Index.with_params(1.5, 0.5)
    assert abs(ix.k1 - 1.5) < 0.01 and abs(ix.b - 0.5) < 0.01


def test_bm25_tokenize(api):                                    # :497-517
    ix = api.BM25Index()
    toks = ix.tokenize("Hello World! This is a test.")
    assert "hello" in toks and "world" in toks and "test" in toks
    assert "this" not in toks and "is" not in toks and "a" not in toks
    toks = ix.tokenize("HELLO World")
    assert "hello" in toks and "world" in toks


@pytest.mark.parametrize("text", [
    "", "   ", "a I x", "Rust's memory-safety: zero_cost abstractions, 100% (no GC)!",
    "The quick brown fox jumps over the lazy dog 42 times", "tab\tseparated\nlines\r\nhere",
    "Ünïcödé Größe ÉCOLE naïve café", "ΑΒΓ δοκιμή Привет МИР", "mixed123abc 7up x9 C3PO", "e-mail@example.com http://a.b/c?d=e",
])
def test_tokenizer_matches_oracle_restatement(api, text):
    assert api.BM25Index().tokenize(text) == oracle_tokenize(text)


def test_bm25_add_chunk_and_batch(api):                         # :519-545
    ix = api.BM25Index()
    ix.add(api.Chunk("Machine learning is fascinating"))
    assert len(ix) == 1 and not ix.is_empty()
    assert ix.contains_term("machine") and ix.contains_term("learning")
    ix2 = api.BM25Index()
    ix2.add_batch([api.Chunk("First document about AI"), api.Chunk("Second document about ML"),
                   api.Chunk("Third document about deep learning")])
    assert len(ix2) == 3


def test_bm25_remove_updates_len(api):                          # :621-634 (search part is a gpu test)
    ix = api.BM25Index()
    c = api.Chunk("Test document")
    ix.add(c)
    assert len(ix) == 1
    ix.remove(c.id)
    assert len(ix) == 0


def test_vector_store_new_and_insert_validation(api):           # :688-726
    store = api.VectorStore.with_dimension(384)
    assert store.dimension == 384 and store.is_empty()
    store = api.VectorStore.with_dimension(3)
    with pytest.raises(api.Error) as e:
        store.insert(api.Chunk("no embedding"))
    assert e.value.kind == "InvalidConfig"
    with pytest.raises(api.Error) as e:
        store.insert(api.Chunk("test", embedding=[1.0, 0.0]))
    assert e.value.kind == "DimensionMismatch" and e.value.expected == 3 and e.value.actual == 2
    store.insert(api.Chunk("ok", embedding=[1.0, 0.0, 0.0]))    # buffered on the host until the first search
    assert len(store) == 1 and not store.is_empty()


def test_vector_store_get_and_remove_nonexistent(api):          # :799-819
    store = api.VectorStore.with_dimension(3)
    c = api.Chunk("test", embedding=[1.0, 0.0, 0.0])
    store.insert(c)
    assert store.get(c.id) == "test"
    assert store.get(api.ChunkId()) is None
    assert store.remove(api.ChunkId()) is False


def test_retrieval_result_best_score(api):                      # src/retrieve.rs:69-75
    r = api.RetrievalResult(api.ChunkId(), "x")
    assert r.best_score() == 0.0
    r.sparse_score = 0.1
    assert r.best_score() == 0.1
    r.dense_score = 0.2
    assert r.best_score() == 0.2
    r.fused_score = 0.3
    assert r.best_score() == 0.3
    r.rerank_score = 0.4
    assert r.best_score() == 0.4


def test_hybrid_config_defaults(api):                           # src/retrieve.rs:91-100, src/fusion.rs:33-37
    c = api.HybridRetrieverConfig()
    assert c.candidates_per_source == 50 and c.use_dense and c.use_sparse
    assert c.fusion.kind == api.RRF and abs(c.fusion.param - 60.0) < 0.01


def test_shard_ranges_are_contiguous_and_cover():
    for n in (0, 1, 7, 8, 10_000_000, 20_000_001):
        for w in (1, 2, 3, 4, 8):
            rs = [shard.shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in rs]
            assert max(sizes) - min(sizes) <= 1


def test_exchange_record_layout_roundtrip(api):
    B, C_ = 5, 7
    rng = np.random.default_rng(0)
    d = (rng.integers(0, 1000, (B, C_)).astype(np.uint32), rng.random((B, C_)).astype(np.float32),
         rng.integers(0, C_ + 1, B).astype(np.uint32))
    s = (rng.integers(0, 1000, (B, C_)).astype(np.uint32), rng.random((B, C_)).astype(np.float32),
         rng.integers(0, C_ + 1, B).astype(np.uint32))
    w = shard.pack_exchange(d, s, B, C_)
    assert w.nbytes == api.exchange_bytes(B, C_)
    ords, scores, n = shard.unpack_exchange(w, B, C_)
    assert np.array_equal(ords[0], d[0]) and np.array_equal(scores[1], s[1]) and np.array_equal(n[1], s[2])


# ---------------------------------------------------------------- product-side synthetic generators vs the oracle's
def test_host_synth_generators_match_oracle(built_lib):
    import ctypes as C
    from oracle import oracle as O
    from trueno_rag_b200._lib import f32p, u32p, u64p
    L = built_lib
    seed, V, D = 0x5EED0004, 3000, 48
    cdf = O.zipf_cdf(V)
    # queries (with and without bf16 rounding of the result)
    q = np.zeros((300, D), np.float32)
    assert L.trr_synth_queries(seed, 5, 300, D, 1000, 1, 1, 0, q.ctypes.data_as(f32p)) == 0
    assert np.array_equal(q, O.synth_queries(seed, 5, 300, D, 1000, corpus_bf16=True, dups=True))
    # query terms
    q_off = np.zeros(41, np.uint32)
    assert L.trr_synth_query_terms(seed, cdf.ctypes.data_as(u64p), V, 0, 40, q_off.ctypes.data_as(u32p), None, 0) == 0
    terms = np.zeros(int(q_off[-1]), np.uint32)
    assert L.trr_synth_query_terms(seed, cdf.ctypes.data_as(u64p), V, 0, 40, q_off.ctypes.data_as(u32p),
                                   terms.ctypes.data_as(u32p), terms.size) == 0
    eo, et = O.synth_query_terms(seed, cdf, 0, 40)
    assert np.array_equal(q_off, eo) and np.array_equal(terms, et)
    # BM25 shard CSR == CSR of the oracle's index restricted to the shard
    n_docs, lo, hi = 5000, 1200, 4100
    doc_off, toks = O.synth_doc_tokens(seed, cdf, 0, n_docs)
    oix = O.BM25(n_terms=V, doc_off=doc_off, tokens=toks)
    term_off, post_doc, post_tf, doc_len, df = oix.csr()
    df_l = np.zeros(V, np.uint32)
    dl = np.zeros(hi - lo, np.uint32)
    tot = C.c_uint64()
    assert L.trr_synth_bm25_count(seed, cdf.ctypes.data_as(u64p), V, lo, hi, df_l.ctypes.data_as(u32p),
                                  dl.ctypes.data_as(u32p), C.byref(tot)) == 0
    assert np.array_equal(dl, doc_len[lo:hi]) and tot.value == int(doc_len[lo:hi].sum())
    keep = (post_doc >= lo) & (post_doc < hi)
    term_of = np.repeat(np.arange(V), np.diff(term_off).astype(np.int64))
    assert np.array_equal(df_l, np.bincount(term_of[keep], minlength=V))
    s_off = np.zeros(V + 1, np.uint64)
    np.cumsum(df_l, out=s_off[1:])
    pd = np.zeros(int(s_off[-1]), np.uint32)
    ptf = np.zeros(int(s_off[-1]), np.uint32)
    assert L.trr_synth_bm25_fill(seed, cdf.ctypes.data_as(u64p), V, lo, hi, s_off.ctypes.data_as(u64p),
                                 pd.ctypes.data_as(u32p), ptf.ctypes.data_as(u32p)) == 0
    assert np.array_equal(pd, post_doc[keep] - lo) and np.array_equal(ptf, post_tf[keep])


@pytest.mark.parametrize("text", [
    "ΟΔΥΣΣΕΥΣ ΣΟΦΟΣ Σ ΑΣ ΑΣ́ ΑΣ́Β Σ́Α",                 # Final_Sigma: word-final, lone, with case-ignorable marks around
    "İstanbul IŞIK ǅungla ǈ ǋ ẞ STRASSE straße",        # multi-code-point and title-case mappings
    "日本語のテキスト 漢字 ｶﾀｶﾅ ㈱ ㊙",                    # CJK letters, half-width kana, enclosed ideographs (No / So)
    "٣٤٥ ۴۵۶ ①②③ Ⅻ ⅻ ½ x² ৳৭",                        # Nd / Nl / No digits of several scripts
    "x̀y áb क्षत्रिय ไทย ภาษา",                            # combining marks: Mn that are / are not Alphabetic
    "ˆˇ ˂˃ ʰʱ ªº µ ·",                                   # modifier letters vs modifier symbols
    "😀emoji😀 𝒜𝓁𝓅𝒽𝒶 𝟘𝟙𝟚 𐐀𐐨 🄰",                        # astral: symbols, math alphanumerics, Deseret case pair
    "ᾈ ᾘ ᾨ ᾼ ῌ ῼ Ω K Å",                                # Greek titlecase with iota, Ohm / Kelvin / Angstrom signs
])
def test_tokenizer_unicode_semantics_match_oracle(api, text):
    assert api.BM25Index().tokenize(text) == oracle_tokenize(text)


def test_tokenizer_random_unicode_matches_oracle(api):
    """the product's generated tables + Final_Sigma code against the oracle (regex UCD properties + CPython str.lower)"""
    import random
    rng = random.Random(20261018)
    blocks = [(0x20, 0x7E), (0xA0, 0x24F), (0x250, 0x36F), (0x370, 0x3FF), (0x400, 0x52F), (0x530, 0x6FF), (0x900, 0xDFF),
              (0xE00, 0x10FF), (0x1D00, 0x1FFF), (0x2000, 0x2BFF), (0x2C00, 0x2DFF), (0x3000, 0x30FF), (0x4E00, 0x4E80),
              (0xA640, 0xA7FF), (0xFB00, 0xFB4F), (0xFF00, 0xFFEF), (0x10400, 0x1044F), (0x1D400, 0x1D7FF), (0x1E900, 0x1E95F),
              (0x1F100, 0x1F1FF), (0x1F600, 0x1F64F)]
    ix = api.BM25Index()
    for _ in range(400):
        chars = []
        for _ in range(rng.randint(1, 60)):
            r = rng.random()
            if r < 0.15:
                chars.append(rng.choice(" Σσςİ.-'"))
            elif r < 0.25:
                cp = rng.randrange(1, 0x110000)  # (U+0000 cannot cross the C-string API of the test binding)
                chars.append(chr(cp) if not 0xD800 <= cp <= 0xDFFF else " ")
            else:
                lo, hi = rng.choice(blocks)
                chars.append(chr(rng.randint(lo, hi)))
        text = "".join(chars)
        assert ix.tokenize(text) == oracle_tokenize(text), [hex(ord(c)) for c in text]


def test_every_code_point_is_classified_like_the_oracle(api):
    """exhaustive over all planes: a code point either joins its neighbours into one token or splits them"""
    import unicodedata
    import regex
    alnum_re = regex.compile(r"[\p{Alphabetic}\p{N}]")

    class alnum:  # the oracle's definition (oracle/tokenizer.py)
        @staticmethod
        def fullmatch(ch):
            return unicodedata.category(ch) != "Cn" and alnum_re.fullmatch(ch)
    ix = api.BM25Index()
    step = 4096
    for base in range(0, 0x110000, step):
        cps = [cp for cp in range(max(base, 1), min(base + step, 0x110000)) if not 0xD800 <= cp <= 0xDFFF]
        if not cps:
            continue
        text = " ".join("qq" + chr(cp) + "zz" for cp in cps)
        toks = ix.tokenize(text)
        i = 0
        for cp in cps:
            if alnum.fullmatch(chr(cp)):
                assert toks[i] == ("qq" + chr(cp) + "zz").lower(), hex(cp)
                i += 1
            else:
                assert toks[i] == "qq" and toks[i + 1] == "zz", hex(cp)
                i += 2
        assert i == len(toks)


def test_unicode_tables_are_generated():
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "gen_unicode_tables.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
