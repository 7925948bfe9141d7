"""The CLI's persisted index (crates/trueno-rag-cli/src/main.rs:133-154): `index.json` -> embedding slab, and
`run_query`'s brute-force cosine scan (:479-495) through K1.  SURVEY 8(f) ranks 2 and 4.

CPU tests: the library's serde_json-compatible reader against Python's json module (strings with escapes, surrogate
pairs, optional / defaulted / unknown fields, numbers narrowed f64 -> f32).  GPU tests: the query scan against the CPU
oracle's cosine (the CLI's `cosine_similarity`, :528-542, has the op order of src/index.rs:440-462) with the stable-sort
tie rule of :494.
"""
import json
import math

import numpy as np
import pytest

from oracle import oracle as O
from trueno_rag_b200 import api


def f32_repr(x) -> str:
    """shortest decimal that round-trips as f32 (what serde_json / ryu prints for an f32)"""
    return np.format_float_positional(np.float32(x), unique=True, trim="0") if abs(x) >= 1e-4 or x == 0 else \
        np.format_float_scientific(np.float32(x), unique=True, trim="0")


def make_json(E, chunks=None, dimension=None, embedder_type="tfidf", model_name=None, pretty=True, shortest=True,
              extra=None):
    n = len(E)
    chunks = chunks if chunks is not None else [
        {"content": f"chunk {i} text", "title": None if i % 3 == 0 else f"doc {i}", "source": f"docs/{i}.md"} for i in range(n)]
    rows = "[" + ",".join("[" + ",".join(f32_repr(v) if shortest else repr(float(v)) for v in r) + "]" for r in E) + "]"
    doc = {"chunks": chunks, "embeddings": "@@", "dimension": dimension if dimension is not None else (len(E[0]) if n else 0)}
    if embedder_type is not None:
        doc["embedder_type"] = embedder_type
    if model_name is not False:
        doc["model_name"] = model_name
    if extra:
        doc.update(extra)
    return json.dumps(doc, indent=2 if pretty else None, ensure_ascii=False).replace('"@@"', rows)


def test_parses_what_serde_json_writes():
    rng = np.random.default_rng(1)
    E = rng.standard_normal((17, 24)).astype(np.float32)
    E[3, 5] = 0.0
    E[4, 0] = np.float32(1e-30)
    E[4, 1] = np.float32(-3.4e38)
    E[4, 2] = np.float32(1.0)
    E[4, 3] = np.float32(1.17549435e-38)   # smallest normal
    E[4, 4] = np.float32(1e-45)            # smallest subnormal
    for pretty in (True, False):
        for shortest in (True, False):
            ix = api.PersistedIndex.from_json(make_json(E, pretty=pretty, shortest=shortest, model_name="BAAI/bge-small-en-v1.5",
                                                        embedder_type="semantic"))
            assert len(ix) == 17 and ix.n_embeddings == 17 and ix.dimension == 24
            assert ix.embedder_type == "semantic" and ix.model_name == "BAAI/bge-small-en-v1.5"
            for i in range(17):
                assert ix.embedding(i).tobytes() == E[i].tobytes()
            c = ix.chunk(4)
            assert (c.content, c.title, c.source) == ("chunk 4 text", "doc 4", "docs/4.md")
            assert ix.chunk(3).title is None


def test_strings_escapes_and_unicode():
    text = 'quote " backslash \\ slash / tab\t newline\n cr\r bs\b ff\f nul-ish  é ☃ 😀 end'
    chunks = [{"content": text, "title": "Ünï", "source": None}]
    for ensure_ascii in (False, True):  # raw UTF-8, and \uXXXX escapes with surrogate pairs
        doc = {"chunks": chunks, "embeddings": [[1.0, 2.0]], "dimension": 2}
        ix = api.PersistedIndex.from_json(json.dumps(doc, ensure_ascii=ensure_ascii))
        c = ix.chunk(0)
        assert c.content == text and c.title == "Ünï" and c.source is None
    ix = api.PersistedIndex.from_json('{"chunks":[{"content":"a\\/b \\u00e9"}],"embeddings":[[0.5]],"dimension":1}')
    assert ix.chunk(0).content == "a/b é" and ix.chunk(0).title is None


def test_defaults_unknown_fields_and_key_order():
    # #[serde(default)] on embedder_type / model_name; unknown keys are ignored; members may come in any order
    t = ('{"future_field": {"a": [1, 2, {"b": null}], "c": "x"}, "dimension": 3, "embeddings": [[1, 2, 3], [4.5e0, -6E-1, 7]], '
         '"chunks": [{"source": null, "content": "one", "more": [true, false]}, {"content": "two"}]}')
    ix = api.PersistedIndex.from_json(t)
    assert ix.embedder_type == "" and ix.model_name is None and ix.dimension == 3
    assert ix.embedding(0).tolist() == [1.0, 2.0, 3.0]           # integers are valid f32 values for serde_json
    assert ix.embedding(1).tobytes() == np.array([4.5, -0.6, 7.0], np.float32).tobytes()
    assert ix.chunk(1).content == "two"


def test_numbers_are_narrowed_like_serde_json():
    # f64 first, then `as f32`: more digits than f32 holds, exponents, values that overflow f32
    toks = ["0.1", "0.30000000000000004", "16777217", "1e-46", "3.4028235e38", "3.5e38", "-3.5e38", "1E2", "-0.0",
            "123456789.123456789"]
    ix = api.PersistedIndex.from_json('{"chunks":[],"embeddings":[[' + ",".join(toks) + ']],"dimension":10}')
    with np.errstate(over="ignore"):
        want = np.array([float(t) for t in toks], np.float64).astype(np.float32)
    got = ix.embedding(0)
    assert got.tobytes() == want.tobytes()
    assert math.isinf(got[5]) and got[5] > 0 and math.isinf(got[6]) and got[6] < 0 and np.signbit(got[8])


def test_empty_index_and_ragged_rows():
    ix = api.PersistedIndex.from_json('{"chunks": [], "embeddings": [], "dimension": 384}')
    assert len(ix) == 0 and ix.n_embeddings == 0 and ix.dimension == 384
    ix = api.PersistedIndex.from_json('{"chunks": [{"content":"a"},{"content":"b"}], "embeddings": [[1,2,3],[]], "dimension": 3}')
    assert ix.embedding(0).size == 3 and ix.embedding(1).size == 0


@pytest.mark.parametrize("bad", [
    "", "{", "[]", '{"chunks": [], "embeddings": []}',                                   # missing field `dimension`
    '{"chunks": [], "dimension": 1}', '{"embeddings": [], "dimension": 1}',
    '{"chunks": [{"title": "t"}], "embeddings": [], "dimension": 1}',                   # missing field `content`
    '{"chunks": [], "embeddings": [[1, null]], "dimension": 2}',                        # serde_json writes null for NaN
    '{"chunks": [], "embeddings": [[1,]], "dimension": 1}', '{"chunks": [], "embeddings": [[01]], "dimension": 1}',
    '{"chunks": [], "embeddings": [[1.]], "dimension": 1}', '{"chunks": [], "embeddings": [[1e]], "dimension": 1}',
    '{"chunks": [], "embeddings": [], "dimension": -1}', '{"chunks": [], "embeddings": [], "dimension": 1.5}',
    '{"chunks": [], "embeddings": [], "dimension": 1} x', '{"chunks": [], "embeddings": [], "dimension": 1, "dimension": 2}',
    '{"chunks": [{"content": "a\\q"}], "embeddings": [], "dimension": 1}',
    '{"chunks": [{"content": "\\ud83d"}], "embeddings": [], "dimension": 1}',
    '{"chunks": [{"content": "a\nb"}], "embeddings": [], "dimension": 1}',               # raw control character
    '{"chunks": "no", "embeddings": [], "dimension": 1}',
])
def test_malformed_input_is_a_serialization_error(bad):
    with pytest.raises(api.Error) as e:
        api.PersistedIndex.from_json(bad)
    assert e.value.kind == "SerializationError"


# ------------------------------------------------------------------------------------------------
# run_query's scan on the device
# ------------------------------------------------------------------------------------------------
def oracle_query(E, q, top_k):
    """:479-495 with the oracle's cosine: stable descending sort, truncate"""
    q = np.asarray(q, np.float32)
    sims = []
    for row in E:
        row = np.asarray(row, np.float32)
        sims.append(0.0 if row.size != q.size or row.size == 0 else float(O.cosine(q, row)))
    order = sorted(range(len(E)), key=lambda i: (-sims[i], i))
    return [(i, sims[i]) for i in order[:top_k]]


@pytest.mark.gpu
@pytest.mark.parametrize("n,d,k", [(1, 8, 5), (50, 16, 5), (300, 384, 10), (2000, 64, 1000)])
def test_query_matches_reference_scan(n, d, k):
    rng = np.random.default_rng(n + d)
    E = rng.standard_normal((n, d)).astype(np.float32)
    if n >= 50:
        E[7] = E[3]          # exact tie: the stable sort keeps index order
        E[11] = 0.0          # zero norm scores 0.0 (:537-538)
    ix = api.PersistedIndex.from_json(make_json(E))
    for s in range(3):
        q = rng.standard_normal(d).astype(np.float32)
        got = ix.query(q, k)
        want = oracle_query(E, q, k)
        assert [i for i, _ in got] == [i for i, _ in want]
        assert np.array([v for _, v in got], np.float32).tobytes() == np.array([v for _, v in want], np.float32).tobytes()


@pytest.mark.gpu
def test_query_with_rows_of_other_lengths_and_wrong_query_length():
    rng = np.random.default_rng(5)
    E = [rng.standard_normal(6).astype(np.float32) for _ in range(12)]
    E[2] = E[2][:4]                      # scores 0.0 against a 6-dimensional query (:529-531)
    E[9] = np.zeros(0, np.float32)
    doc = {"chunks": [{"content": str(i)} for i in range(12)], "embeddings": [[float(v) for v in r] for r in E], "dimension": 6}
    ix = api.PersistedIndex.from_json(json.dumps(doc))
    q = rng.standard_normal(6).astype(np.float32)
    got, want = ix.query(q, 12), oracle_query(E, q, 12)
    assert [i for i, _ in got] == [i for i, _ in want]
    assert np.array([v for _, v in got], np.float32).tobytes() == np.array([v for _, v in want], np.float32).tobytes()
    # a query of a length no row has: every score is 0.0 and the order is the index order
    assert ix.query(np.ones(5, np.float32), 4) == [(0, 0.0), (1, 0.0), (2, 0.0), (3, 0.0)]
    # a 4-dimensional query only matches row 2
    q4 = E[2].copy()
    got = ix.query(q4, 3)
    assert got[0][0] == 2 and abs(got[0][1] - 1.0) < 1e-6 and [i for i, _ in got[1:]] == [0, 1]
    assert ix.query(q, 0) == []


@pytest.mark.gpu
def test_top_k_beyond_the_device_limit_is_reported():
    E = np.random.default_rng(3).standard_normal((1500, 8)).astype(np.float32)
    ix = api.PersistedIndex.from_json(make_json(E))
    with pytest.raises(api.Error) as e:  # the C ABI serves k <= 1024 per query; no host-side scan is substituted
        ix.query(E[0], 1500)
    assert e.value.kind == "Unsupported"


# ------------------------------------------------------------------------------------------------
# writing index.json (run_index, :407-424)
# ------------------------------------------------------------------------------------------------
def _build(E, chunks):
    ix = api.PersistedIndex.new(len(E[0]) if len(E) else 0, "tfidf", None)
    for c, e in zip(chunks, E):
        ix.push(api.PersistedChunk(c["content"], c["title"], c["source"]), e)
    return ix


def test_to_json_round_trips_and_matches_pretty_layout():
    rng = np.random.default_rng(9)
    E = rng.standard_normal((7, 12)).astype(np.float32)
    E[0, :6] = [0.0, -0.0, 1.0, 1e21, np.float32(1e-45), np.float32(3.4028235e38)]
    chunks = [{"content": f'c{i} "q" \\ / \n\t\r\b\f \x01 é ☃ 😀', "title": None if i % 2 else f"T{i}", "source": f"s{i}.md"} for i in range(7)]
    text = _build(E, chunks).to_json()
    doc = json.loads(text)
    assert doc["chunks"] == chunks and doc["dimension"] == 12 and doc["embedder_type"] == "tfidf" and doc["model_name"] is None
    assert list(doc) == ["chunks", "embeddings", "dimension", "embedder_type", "model_name"]   # struct field order
    with np.errstate(over="ignore"):
        assert np.array(doc["embeddings"], np.float64).astype(np.float32).tobytes() == E.tobytes()
    # the layout of serde_json's PrettyFormatter is the layout of json.dumps(indent=2); only number spelling may differ
    skeleton = lambda s: __import__("re").sub(r"-?\d[0-9.eE+-]*", "#", s)  # noqa: E731
    assert skeleton(text) == skeleton(json.dumps(doc, indent=2, ensure_ascii=False))
    again = api.PersistedIndex.from_json(text)
    assert again.to_json() == text
    for i in range(7):
        assert again.embedding(i).tobytes() == E[i].tobytes()
        assert again.chunk(i).content == chunks[i]["content"]


def test_to_json_prints_the_shortest_round_tripping_digits():
    rng = np.random.default_rng(10)
    vals = np.concatenate([rng.standard_normal(2000).astype(np.float32), (rng.standard_normal(500) * 1e-20).astype(np.float32),
                           (rng.standard_normal(500) * 1e20).astype(np.float32),
                           np.array([0.1, 0.2, 0.3, 1 / 3, 16777216.0, 1e-38, 123456.79, 5e-324], np.float32)])
    ix = api.PersistedIndex.new(len(vals))
    ix.push(api.PersistedChunk("x"), vals)
    toks = [t.strip().rstrip(",") for t in ix.to_json().split('"embeddings": [')[1].split("]")[0].split("\n") if t.strip() not in ("", "[")]
    assert len(toks) == len(vals)
    for t, v in zip(toks, vals):
        assert np.float32(float(t)).tobytes() == np.float32(v).tobytes(), (t, v)
        digits = lambda s: len(s.lower().split("e")[0].replace("-", "").replace(".", "").strip("0")) or 1  # noqa: E731
        want = np.format_float_scientific(v, unique=True, trim="-")
        assert digits(t) <= digits(want), (t, want)          # never more digits than the shortest representation
        assert "." in t or "e" in t                            # always spelled as a float


def test_to_json_empty_and_non_finite():
    assert json.loads(api.PersistedIndex.new(384, "semantic", "BAAI/bge-small-en-v1.5").to_json()) == {
        "chunks": [], "embeddings": [], "dimension": 384, "embedder_type": "semantic", "model_name": "BAAI/bge-small-en-v1.5"}
    ix = api.PersistedIndex.new(2)
    ix.push(api.PersistedChunk("x"), np.array([np.nan, np.inf], np.float32))
    text = ix.to_json()
    assert json.loads(text)["embeddings"] == [[None, None]]        # serde_json writes null for NaN / infinity ...
    with pytest.raises(api.Error):
        api.PersistedIndex.from_json(text)                          # ... and cannot read it back (neither can the reference)
