"""BM25Index persistence in the REFERENCE's on-disk format.

Mirrors the reference's own tests in src/compressed.rs:110-300 and pins the byte layout against an independent
pure-Python restatement of the formats involved (and, for ZSTD, against libzstd itself through pyarrow):
  * bincode 1.3 default options of `#[derive(Serialize)] struct BM25Index` (src/index.rs:30-51): little-endian,
    fixed-width integers, u64 lengths, maps as (len, entries), `ChunkId(uuid::Uuid)` as u64 16 + the 16 big-endian bytes;
  * lz4_flex 0.11 `compress_prepend_size`: u32 little-endian uncompressed size + one LZ4 block;
  * ZSTD: standard frames (RFC 8878) - the library's decoder must read what libzstd writes at any level, and libzstd
    must read the library's (store-only) frames.
There is no Rust toolchain in the image, so the layout is pinned to those published formats (and to a hand-spelled golden
record below), not to bytes produced by the reference itself.
"""
import random
import struct
import uuid

import pytest

from trueno_rag_b200 import api

DEFAULT_STOPWORDS = 90  # src/index.rs:93-106


# ------------------------------------------------------------------------------------------------
# independent restatements (test infrastructure)
# ------------------------------------------------------------------------------------------------
class Rd:
    def __init__(self, b):
        self.b, self.i = b, 0

    def take(self, n):
        assert self.i + n <= len(self.b)
        v = self.b[self.i:self.i + n]
        self.i += n
        return v

    def u8(self): return self.take(1)[0]
    def u32(self): return struct.unpack("<I", self.take(4))[0]
    def u64(self): return struct.unpack("<Q", self.take(8))[0]
    def f32(self): return struct.unpack("<f", self.take(4))[0]
    def s(self): return self.take(self.u64()).decode()

    def cid(self):
        assert self.u64() == 16
        return uuid.UUID(bytes=self.take(16))


def read_bm25(b):
    r = Rd(b)
    inv = {}
    for _ in range(r.u64()):
        t = r.s()
        inv[t] = [(r.cid(), r.u32()) for _ in range(r.u64())]
    dfs = {}
    for _ in range(r.u64()):
        t = r.s()
        dfs[t] = r.u32()
    lens = {}
    for _ in range(r.u64()):
        c = r.cid()
        lens[c] = r.u32()
    out = dict(inv=inv, dfs=dfs, lens=lens, avg=r.f32(), count=r.u32(), k1=r.f32(), b=r.f32(), lowercase=r.u8())
    out["stop"] = {r.s() for _ in range(r.u64())}
    assert r.i == len(b)
    return out


def write_bm25(d, rng=None):
    def s(x):
        e = x.encode()
        return struct.pack("<Q", len(e)) + e

    def cid(c):
        return struct.pack("<Q", 16) + c.bytes

    def items(m):
        it = list(m.items() if isinstance(m, dict) else m)
        if rng:
            rng.shuffle(it)  # a HashMap serialises in arbitrary order
        return it

    out = struct.pack("<Q", len(d["inv"]))
    for t, pl in items(d["inv"]):
        out += s(t) + struct.pack("<Q", len(pl)) + b"".join(cid(c) + struct.pack("<I", tf) for c, tf in pl)
    out += struct.pack("<Q", len(d["dfs"])) + b"".join(s(t) + struct.pack("<I", v) for t, v in items(d["dfs"]))
    out += struct.pack("<Q", len(d["lens"])) + b"".join(cid(c) + struct.pack("<I", v) for c, v in items(d["lens"]))
    out += struct.pack("<fIffB", d["avg"], d["count"], d["k1"], d["b"], d["lowercase"])
    st = items([(x, None) for x in d["stop"]])
    out += struct.pack("<Q", len(st)) + b"".join(s(x) for x, _ in st)
    return out


def lz4_block_decode(b, out_size):
    out, i = bytearray(), 0
    while i < len(b):
        tok = b[i]; i += 1
        lit = tok >> 4
        if lit == 15:
            while True:
                x = b[i]; i += 1; lit += x
                if x != 255:
                    break
        out += b[i:i + lit]; i += lit
        if i == len(b):
            break
        off = b[i] | (b[i + 1] << 8); i += 2
        ml = tok & 15
        if ml == 15:
            while True:
                x = b[i]; i += 1; ml += x
                if x != 255:
                    break
        ml += 4
        assert 0 < off <= len(out)
        for _ in range(ml):
            out.append(out[-off])
    assert len(out) == out_size
    return bytes(out)


def lz4_block_encode(data, min_match=4):
    """greedy encoder with a dictionary of 4-byte windows; obeys the end-of-block rules (12 / 5 bytes)"""
    n, out, anchor, i, seen = len(data), bytearray(), 0, 0, {}

    def length(v):
        r = bytearray()
        while v >= 255:
            r.append(255); v -= 255
        r.append(v)
        return r

    def emit(lit, ml, off):
        tok = (min(len(lit), 15) << 4) | (min(ml - 4, 15) if ml else 0)
        out.append(tok)
        if len(lit) >= 15:
            out.extend(length(len(lit) - 15))
        out.extend(lit)
        if ml:
            out.extend(struct.pack("<H", off))
            if ml - 4 >= 15:
                out.extend(length(ml - 4 - 15))

    while n > 12 and i <= n - 12:
        w = data[i:i + 4]
        c = seen.get(w)
        seen[w] = i
        if c is not None and i - c <= 65535:
            ml = 4
            while i + ml < n - 5 and data[c + ml] == data[i + ml]:
                ml += 1
            if ml >= min_match:
                emit(data[anchor:i], ml, i - c)
                i += ml
                anchor = i
                continue
        i += 1
    emit(data[anchor:], 0, 0)
    return bytes(out)


def prepend_size(block, n):
    return struct.pack("<I", n) + block


# ------------------------------------------------------------------------------------------------
# src/compressed.rs tests
# ------------------------------------------------------------------------------------------------
def test_compression_as_str():  # :122-126
    assert api.Compression.as_str(api.Compression.Lz4) == "lz4"
    assert api.Compression.as_str(api.Compression.Zstd) == "zstd"


def test_compression_default():  # :128-131
    assert api.Compression.default() == api.Compression.Lz4


def test_lz4_compress_decompress():  # :133-139
    data = b"hello world hello world hello world"
    c = api.compress(data, api.Compression.Lz4)
    assert api.decompress(c, api.Compression.Lz4) == data
    # layout: u32 size prefix, then a block any LZ4 decoder accepts
    assert struct.unpack("<I", c[:4])[0] == len(data)
    assert lz4_block_decode(c[4:], len(data)) == data


def test_empty_data_compression():  # :149-163
    assert api.compress(b"", api.Compression.Lz4) == b""
    assert api.decompress(b"", api.Compression.Lz4) == b""
    assert api.compress(b"", api.Compression.Zstd) == b""
    assert api.decompress(b"", api.Compression.Zstd) == b""


def test_lz4_compresses_repeated_data():  # :165-171
    data = bytes(10000)
    c = api.compress(data)
    assert len(c) < len(data) // 10
    assert lz4_block_decode(c[4:], len(data)) == data


def test_zstd_compress_decompress():  # :141-147
    data = b"hello world hello world hello world"
    c = api.compress(data, api.Compression.Zstd)
    assert api.decompress(c, api.Compression.Zstd) == data
    assert c[:4] == bytes.fromhex("28b52ffd")                       # a standard frame ...
    assert _real_zstd_decompress(c, len(data)) == data              # ... that libzstd reads


def test_zstd_compresses_repeated_data():  # :173-179
    data = bytes(10000)
    c = api.compress(data, api.Compression.Zstd)
    assert len(c) < len(data) // 10
    assert _real_zstd_decompress(c, len(data)) == data


# ---- the ZSTD frame decoder against frames written by libzstd (through pyarrow, test infrastructure) ----
def _real_zstd_compress(data, level=3):
    import pyarrow as pa
    return pa.Codec("zstd", compression_level=level).compress(data, asbytes=True)


def _real_zstd_decompress(frame, n):
    import pyarrow as pa
    return pa.Codec("zstd").decompress(frame, decompressed_size=n, asbytes=True) if n else b""


def _corpus(seed):
    rng = random.Random(seed)
    words = [bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyz") for _ in range(rng.randint(2, 9))) for _ in range(300)]
    text = b" ".join(rng.choice(words) for _ in range(60000))                      # Huffman literals + FSE sequences
    noise = bytes(rng.getrandbits(8) for _ in range(200000))                       # raw blocks, > 128 KB
    runs = b"".join(bytes([rng.getrandbits(8)]) * rng.randint(1, 5000) for _ in range(200))   # RLE, long matches
    skew = bytes(rng.choice(b"aaaaaaaabbbbccd") for _ in range(150000))            # few symbols: short Huffman codes
    period = (b"0123456789abcdef" * 3 + b"XY") * 4000                               # repeat offsets
    ix = api.BM25Index()
    for i in range(200):
        ix.add(api.Chunk(" ".join(rng.choice(words).decode() for _ in range(rng.randint(5, 40)))))
    return dict(text=text, noise=noise, runs=runs, skew=skew, period=period, bincode=ix.to_bytes(),
                tiny=b"a", small=b"abcabcabcabc", mixed=text[:70000] + noise[:70000] + runs[:70000] + text[:70000])


@pytest.mark.parametrize("level", [1, 3, 9, 19])
def test_zstd_decoder_reads_libzstd_frames(level):
    for name, data in _corpus(level).items():
        frame = _real_zstd_compress(data, level)
        assert api.decompress(frame, api.Compression.Zstd) == data, name


@pytest.mark.parametrize("n", [1, 2, 3, 7, 31, 32, 33, 255, 256, 257, 65535, 65536, 65791, 65792, 131071, 131072, 131073, 400000])
def test_zstd_sizes_both_ways(n):
    rng = random.Random(n)
    for data in (bytes(rng.choice(b"ab ") for _ in range(n)), bytes(n), bytes(rng.getrandbits(8) for _ in range(n))):
        assert api.decompress(_real_zstd_compress(data), api.Compression.Zstd) == data
        mine = api.compress(data, api.Compression.Zstd)
        assert _real_zstd_decompress(mine, n) == data
        assert api.decompress(mine, api.Compression.Zstd) == data


def _xxh64(data, seed=0):
    M = (1 << 64) - 1
    P1, P2, P3, P4, P5 = 11400714785074694791, 14029467366897019727, 1609587929392839161, 9650029242287828579, 2870177450012600261
    rotl = lambda x, r: ((x << r) | (x >> (64 - r))) & M
    rnd = lambda acc, v: (rotl((acc + v * P2) & M, 31) * P1) & M
    n, i = len(data), 0
    if n >= 32:
        v = [(seed + P1 + P2) & M, (seed + P2) & M, seed, (seed - P1) & M]
        while i + 32 <= n:
            for k in range(4):
                v[k] = rnd(v[k], struct.unpack_from("<Q", data, i + 8 * k)[0])
            i += 32
        h = (rotl(v[0], 1) + rotl(v[1], 7) + rotl(v[2], 12) + rotl(v[3], 18)) & M
        for k in range(4):
            h = ((h ^ rnd(0, v[k])) * P1 + P4) & M
    else:
        h = (seed + P5) & M
    h = (h + n) & M
    while i + 8 <= n:
        h = (rotl(h ^ rnd(0, struct.unpack_from("<Q", data, i)[0]), 27) * P1 + P4) & M
        i += 8
    if i + 4 <= n:
        h = (rotl(h ^ (struct.unpack_from("<I", data, i)[0] * P1) & M, 23) * P2 + P3) & M
        i += 4
    while i < n:
        h = (rotl(h ^ (data[i] * P5) & M, 11) * P1) & M
        i += 1
    h ^= h >> 33
    h = (h * P2) & M
    h ^= h >> 29
    h = (h * P3) & M
    return h ^ (h >> 32)


def test_zstd_checksum_skippable_and_concatenated_frames():
    a, b = b"first frame " * 500, bytes(range(256)) * 40
    fa, fb = bytearray(_real_zstd_compress(a)), _real_zstd_compress(b)
    assert _xxh64(b"") == 0xEF46DB3751D8E999 and _xxh64(b"abc") == 0x44BC2CF5AD770999   # published XXH64 vectors
    assert not fa[4] & 4
    fa[4] |= 4                                                       # Content_Checksum_flag
    with_sum = bytes(fa) + struct.pack("<I", _xxh64(a) & 0xFFFFFFFF)
    assert api.decompress(with_sum, api.Compression.Zstd) == a
    skippable = struct.pack("<II", 0x184D2A53, 5) + b"hello"
    assert api.decompress(skippable + with_sum + skippable + fb, api.Compression.Zstd) == a + b
    bad_sum = bytes(fa) + struct.pack("<I", (_xxh64(a) + 1) & 0xFFFFFFFF)
    with pytest.raises(api.Error) as e:
        api.decompress(bad_sum, api.Compression.Zstd)
    assert e.value.kind == "SerializationError" and "checksum" in str(e.value)


def test_zstd_corrupt_frames_never_crash():
    """truncations and byte flips of valid frames: an error or (rarely) some bytes, never a crash or a runaway allocation"""
    rng = random.Random(11)
    data = _corpus(3)["mixed"]
    frame = _real_zstd_compress(data)
    for k in (0, 1, 3, 4, 5, 6, 9, 12, 40, len(frame) // 2, len(frame) - 1):   # every truncation is an error
        if k == 0:
            assert api.decompress(frame[:0], api.Compression.Zstd) == b""         # (:54-56: empty in, empty out)
            continue
        with pytest.raises(api.Error) as e:
            api.decompress(frame[:k], api.Compression.Zstd)
        assert e.value.kind == "SerializationError"
    cases = []
    for _ in range(300):
        f = bytearray(frame)
        for _ in range(rng.randint(1, 4)):
            f[rng.randrange(len(f))] ^= 1 << rng.randrange(8)
        cases.append(bytes(f))
    # the same on a frame that is all entropy-coded (a flip in a raw block is undetectable without a checksum)
    text = _corpus(3)["text"]
    tframe = _real_zstd_compress(text)
    for _ in range(300):
        f = bytearray(tframe)
        f[rng.randrange(5, len(f))] ^= 1 << rng.randrange(8)
        cases.append(bytes(f))
    cases += [b"\x00" * 16, bytes.fromhex("28b52ffd") + bytes(rng.getrandbits(8) for _ in range(64)), bytes.fromhex("28b52ffd2000")]
    errors = 0
    for c in cases:
        try:
            out = api.decompress(c, api.Compression.Zstd)
            assert len(out) <= 16 * len(data)
        except api.Error as e:
            assert e.kind == "SerializationError"
            errors += 1
    assert errors > 100


def test_bm25_zstd_roundtrip():  # :201-211
    ix = api.BM25Index()
    ix.add(api.Chunk("rust programming language"))
    ix.add(api.Chunk("systems programming with rust"))
    c = ix.to_compressed_bytes(api.Compression.Zstd)
    r = api.BM25Index.from_compressed_bytes(c, api.Compression.Zstd)
    assert len(ix) == len(r)
    # and an index file as the reference would write it: bincode through libzstd at the reference's level 3 (:43)
    r2 = api.BM25Index.from_compressed_bytes(_real_zstd_compress(ix.to_bytes(), 3), api.Compression.Zstd)
    assert read_bm25(r2.to_bytes()) == read_bm25(r.to_bytes())


@pytest.mark.parametrize("seed", range(6))
def test_lz4_cross_decoding(seed):
    """library encoder -> python decoder, python encoder -> library decoder, on data with long literal runs, long matches
    and overlapping (run-length) matches"""
    rng = random.Random(seed)
    parts = []
    for _ in range(rng.randint(1, 40)):
        kind = rng.randint(0, 3)
        if kind == 0:
            parts.append(bytes(rng.getrandbits(8) for _ in range(rng.randint(0, 700))))
        elif kind == 1:
            parts.append(bytes([rng.getrandbits(8)]) * rng.randint(1, 3000))
        elif kind == 2 and parts:
            parts.append(rng.choice(parts)[: rng.randint(0, 400)])
        else:
            parts.append(b"the quick brown fox " * rng.randint(1, 30))
    data = b"".join(parts)
    c = api.compress(data)
    if data:
        assert struct.unpack("<I", c[:4])[0] == len(data)
        assert lz4_block_decode(c[4:], len(data)) == data
    assert api.decompress(c) == data
    mine = prepend_size(lz4_block_encode(data), len(data)) if data else b""
    assert api.decompress(mine) == data


@pytest.mark.parametrize("n", [1, 4, 5, 11, 12, 13, 14, 15, 16, 17, 19, 270, 271, 4096])
def test_lz4_small_and_boundary_sizes(n):
    for data in (bytes(n), bytes(range(256)) * (n // 256 + 1)):
        data = data[:n]
        c = api.compress(data)
        assert lz4_block_decode(c[4:], n) == data
        assert api.decompress(c) == data


def test_lz4_corrupt_input_is_an_error():
    good = api.compress(b"abcdabcdabcdabcdabcdabcdabcdabcd-abcdabcdabcd")
    for bad in (good[:3], good[:-3], struct.pack("<I", 7) + good[4:], struct.pack("<I", 1000) + good[4:],
                struct.pack("<I", 8) + bytes([0x10, 65, 9, 0])):  # offset 9 reaches before the start of the output
        with pytest.raises(api.Error) as e:
            api.decompress(bad)
        assert e.value.kind == "SerializationError"


# ------------------------------------------------------------------------------------------------
# bincode layout of BM25Index
# ------------------------------------------------------------------------------------------------
def _index3():
    ix = api.BM25Index()
    chunks = [api.Chunk("machine learning is great"), api.Chunk("deep learning neural networks"),
              api.Chunk("natural language processing learning learning")]
    for c in chunks:
        ix.add(c)
    return ix, chunks


def test_to_bytes_layout_matches_the_struct():
    ix, chunks = _index3()
    d = read_bm25(ix.to_bytes())
    ids = [c.id.value for c in chunks]
    assert d["count"] == 3 and d["k1"] == pytest.approx(1.2) and d["b"] == pytest.approx(0.75) and d["lowercase"] == 1
    assert len(d["stop"]) == DEFAULT_STOPWORDS and "the" in d["stop"]
    assert d["lens"] == {ids[0]: 3, ids[1]: 4, ids[2]: 5}       # "is" is a stopword
    assert struct.pack("<f", d["avg"]) == struct.pack("<f", 12 / 3)
    assert d["dfs"]["learning"] == 3 and d["dfs"]["machine"] == 1
    assert d["inv"]["learning"] == [(ids[0], 1), (ids[1], 1), (ids[2], 2)]  # push order = insertion order (:196)
    assert set(d["inv"]) == set(d["dfs"]) == {"machine", "learning", "great", "deep", "neural", "networks", "natural",
                                              "language", "processing"}


def test_golden_record_spelled_by_hand():
    """one chunk, one term, no stopwords - every byte written out"""
    cid = uuid.UUID("00112233-4455-6677-8899-aabbccddeeff")
    rec = b"".join([
        struct.pack("<Q", 1),                                    # inverted_index: 1 entry
        struct.pack("<Q", 2), b"ab",                             #   key "ab"
        struct.pack("<Q", 1),                                    #   Vec of 1 posting
        struct.pack("<Q", 16), bytes.fromhex("00112233445566778899aabbccddeeff"), struct.pack("<I", 3),  # (ChunkId, tf 3)
        struct.pack("<Q", 1), struct.pack("<Q", 2), b"ab", struct.pack("<I", 1),                          # doc_freqs
        struct.pack("<Q", 1), struct.pack("<Q", 16), bytes.fromhex("00112233445566778899aabbccddeeff"),
        struct.pack("<I", 3),                                    # doc_lengths
        struct.pack("<f", 3.0), struct.pack("<I", 1),            # avg_doc_length, doc_count
        struct.pack("<f", 1.5), struct.pack("<f", 0.5),          # k1, b
        b"\x00",                                                 # lowercase = false
        struct.pack("<Q", 0),                                    # stopwords: empty set
    ])
    ix = api.BM25Index.from_bytes(rec)
    assert len(ix) == 1 and ix.k1 == 1.5 and ix.b == 0.5 and ix.avg_doc_length == 3.0
    assert ix.contains_term("ab") and not ix.contains_term("the")
    assert ix.tokenize("The AB ab") == ["The", "AB", "ab"]       # no lowercasing, no stopwords
    assert ix.to_bytes() == rec                                  # a single-entry index has one possible encoding
    assert read_bm25(rec)["inv"]["ab"] == [(cid, 3)]


@pytest.mark.parametrize("seed", range(4))
def test_reads_any_map_order_and_round_trips(seed):
    rng = random.Random(100 + seed)
    ids = [uuid.UUID(int=rng.getrandbits(128)) for _ in range(rng.randint(1, 40))]
    vocab = ["t%03d" % i for i in range(rng.randint(1, 60))]
    inv, lens = {}, {}
    for c in ids:
        terms = rng.sample(vocab, rng.randint(1, min(8, len(vocab))))
        tfs = [rng.randint(1, 5) for _ in terms]
        lens[c] = sum(tfs)
        for t, tf in zip(terms, tfs):
            inv.setdefault(t, []).append((c, tf))
    for pl in inv.values():
        rng.shuffle(pl)
    total = sum(lens.values())
    avg = struct.unpack("<f", struct.pack("<f", total / len(ids)))[0]
    d = dict(inv=inv, dfs={t: len(pl) for t, pl in inv.items()}, lens=lens, avg=avg, count=len(ids), k1=0.9, b=0.4,
             lowercase=1, stop={"foo", "bar"})
    ix = api.BM25Index.from_bytes(write_bm25(d, rng))
    assert len(ix) == len(ids) and ix.avg_doc_length == avg
    back = read_bm25(ix.to_bytes())
    assert back["lens"] == lens and back["dfs"] == d["dfs"] and back["stop"] == d["stop"]
    assert back["count"] == len(ids) and back["lowercase"] == 1
    assert struct.pack("<f", back["avg"]) == struct.pack("<f", avg)
    assert {t: sorted(pl) for t, pl in back["inv"].items()} == {t: sorted(pl) for t, pl in inv.items()}
    # the library numbers chunks by ascending ChunkId: postings come back in that order
    for pl in back["inv"].values():
        assert [c.int for c, _ in pl] == sorted(c.int for c, _ in pl)
    # through LZ4 as well
    again = api.BM25Index.from_compressed_bytes(ix.to_compressed_bytes(api.Compression.Lz4), api.Compression.Lz4)
    assert read_bm25(again.to_bytes()) == back


def test_stored_average_is_kept_until_the_next_mutation():
    ix, _ = _index3()
    d = read_bm25(ix.to_bytes())
    d["avg"] = 7.25  # inconsistent with doc_lengths: the reference would score with it as stored
    r = api.BM25Index.from_bytes(write_bm25(d))
    assert r.avg_doc_length == 7.25
    r.add(api.Chunk("extra machine learning text"))  # update_avg_doc_length (:203)
    assert struct.pack("<f", r.avg_doc_length) == struct.pack("<f", (12 + 4) / 4)


def test_bm25_empty_index_compression():  # :236-244
    ix = api.BM25Index()
    c = ix.to_compressed_bytes(api.Compression.Lz4)
    r = api.BM25Index.from_compressed_bytes(c, api.Compression.Lz4)
    assert r.is_empty()
    assert len(read_bm25(r.to_bytes())["stop"]) == DEFAULT_STOPWORDS


def test_bm25_compression_reduces_size():  # :213-233
    ix = api.BM25Index()
    for i in range(100):
        ix.add(api.Chunk(f"document number {i} about machine learning and artificial intelligence"))
    uncompressed = ix.to_bytes()
    lz4 = ix.to_compressed_bytes(api.Compression.Lz4)
    zstd = ix.to_compressed_bytes(api.Compression.Zstd)
    assert len(lz4) < len(uncompressed) and len(zstd) < len(uncompressed)
    assert len(zstd) <= len(lz4)                                   # "ZSTD typically achieves better compression than LZ4"
    assert _real_zstd_decompress(zstd, len(uncompressed)) == uncompressed


@pytest.mark.parametrize("seed", [1, 2])
def test_zstd_compressor_output_is_read_by_libzstd(seed):
    """the library's frames (LZ77 + predefined-FSE sequences, raw literals, RLE / stored blocks) through libzstd and through
    the library's own decoder"""
    for name, data in _corpus(seed).items():
        c = api.compress(data, api.Compression.Zstd)
        assert _real_zstd_decompress(c, len(data)) == data, name
        assert api.decompress(c, api.Compression.Zstd) == data, name
        if name in ("text", "runs", "skew", "period", "bincode", "mixed"):
            assert len(c) < len(data), name
            assert len(c) <= len(api.compress(data, api.Compression.Lz4)), name


def test_removed_chunks_and_emptied_terms_are_not_written():
    ix, chunks = _index3()
    ix.remove(chunks[0].id)  # "machine" and "great" lose their last document (:262-271)
    d = read_bm25(ix.to_bytes())
    assert d["count"] == 2 and chunks[0].id.value not in d["lens"]
    assert "machine" not in d["inv"] and "machine" not in d["dfs"] and d["dfs"]["learning"] == 2
    r = api.BM25Index.from_bytes(ix.to_bytes())
    assert len(r) == 2 and not r.contains_term("machine")


def test_truncated_or_trailing_bytes_are_errors():
    ix, _ = _index3()
    b = ix.to_bytes()
    for bad in (b[:-1], b[:40], b + b"\x00", b""):
        with pytest.raises(api.Error) as e:
            api.BM25Index.from_bytes(bad)
        assert e.value.kind == "SerializationError"
    huge = struct.pack("<Q", 1 << 60) + b"\x00" * 32  # a length prefix the input cannot hold must not allocate
    with pytest.raises(api.Error):
        api.BM25Index.from_bytes(huge)


# ------------------------------------------------------------------------------------------------
# device behaviour of a restored index
# ------------------------------------------------------------------------------------------------
def _by_score_then_id(res):
    return sorted(((s, c.value.int) for c, s in res), key=lambda x: (-x[0], x[1]))


@pytest.mark.gpu
def test_bm25_lz4_roundtrip_search():  # :185-199
    ix, _ = _index3()
    r = api.BM25Index.from_compressed_bytes(ix.to_compressed_bytes(api.Compression.Lz4), api.Compression.Lz4)
    assert len(ix) == len(r)
    assert _by_score_then_id(ix.search("machine learning", 10)) == _by_score_then_id(r.search("machine learning", 10))


@pytest.mark.gpu
def test_bm25_preserved_search_behavior():  # :246-270
    ix = api.BM25Index()
    for t in ("python programming language scripting", "javascript web development frontend",
              "rust systems programming performance"):
        ix.add(api.Chunk(t))
    r = api.BM25Index.from_compressed_bytes(ix.to_compressed_bytes(api.Compression.Lz4), api.Compression.Lz4)
    a, b = ix.search("programming language", 3), r.search("programming language", 3)
    assert len(a) == len(b) == 2
    assert _by_score_then_id(a) == _by_score_then_id(b)  # bit-identical scores


@pytest.mark.gpu
def test_restored_index_keeps_accepting_adds_and_removes():
    rng = random.Random(7)
    vocab = ["w%02d" % i for i in range(40)]
    chunks = [api.Chunk(" ".join(rng.choice(vocab) for _ in range(rng.randint(3, 12)))) for _ in range(60)]
    ix = api.BM25Index()
    for c in chunks[:40]:
        ix.add(c)
    r = api.BM25Index.from_bytes(ix.to_bytes())
    for c in chunks[40:]:
        ix.add(c)
        r.add(c)
    ix.remove(chunks[3].id)
    r.remove(chunks[3].id)
    for q in ("w01 w02 w03", "w10 w10 w39", "w05"):
        assert _by_score_then_id(ix.search(q, 20)) == _by_score_then_id(r.search(q, 20))


# ------------------------------------------------------------------------------------------------
# committed fixtures (tests/golden/make_persistence_golden.py): a libzstd frame and an independently encoded LZ4 block
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,compression", [("bm25_index.zst", 1), ("bm25_index.lz4", 0)])
def test_golden_index_files_load(name, compression):
    import json
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    want = json.load(open(os.path.join(here, "bm25_index.json")))
    data = open(os.path.join(here, name), "rb").read()
    raw = api.decompress(data, compression)
    assert len(raw) == want["bincode_len"]
    ix = api.BM25Index.from_compressed_bytes(data, compression)
    got = read_bm25(ix.to_bytes())
    assert len(ix) == want["count"] == got["count"]
    assert struct.unpack("<I", struct.pack("<f", got["avg"]))[0] == want["avg_bits"]
    assert struct.unpack("<I", struct.pack("<f", got["k1"]))[0] == want["k1_bits"] and got["b"] == want["b"]
    assert got["lowercase"] == want["lowercase"] and sorted(got["stop"]) == want["stop"]
    assert got["dfs"] == want["dfs"]
    assert {c.hex: v for c, v in got["lens"].items()} == want["lens"]
    assert {t: [[c.hex, tf] for c, tf in sorted(pl)] for t, pl in got["inv"].items()} == want["inv"]
