"""Oracle-side restatement of BM25Index::tokenize (reference src/index.rs:93-124).  TEST INFRASTRUCTURE ONLY.

split on `!char::is_alphanumeric`, drop empties, lowercase, drop the 90 stopwords, drop tokens whose UTF-8 byte
length is < 2.  Rust's `char::is_alphanumeric` is `Alphabetic || Nd || Nl || No`, taken here from the `regex` package's
UCD properties; `str::to_lowercase` (full mapping + Final_Sigma) is what CPython's `str.lower()` implements.  The product's
tokenizer uses its own generated tables and its own Final_Sigma code, so the comparison is between two implementations.
Without the `regex` package the split falls back to `str.isalnum()` (letters and digits only: exact on ASCII/Latin/Greek/
Cyrillic text, not on marks and letter-like symbols)."""
import unicodedata as _ud

try:
    import regex as _regex
    _ALNUM = _regex.compile(r"[\p{Alphabetic}\p{N}]")

    def _is_alnum(ch):  # one Unicode version: only code points CPython's UCD (the source of str.lower) assigns
        return _ud.category(ch) != "Cn" and _ALNUM.fullmatch(ch) is not None
except ImportError:  # pragma: no cover
    def _is_alnum(ch):
        return ch.isalnum()

STOPWORDS = frozenset("""a an the is are was were be been being have has had do does did will would could should may
might must shall can need dare ought used to of in for on with at by from as into through during before after above
below between under again further then once here there when where why how all each few more most other some such no
nor not only own same so than too very just and but if or because until while this that these those it its""".split())


def tokenize(text: str, stopwords=STOPWORDS, lowercase=True):
    out, cur = [], []

    def emit():
        if cur:
            tok = "".join(cur)
            if lowercase:
                tok = tok.lower()
            if tok not in stopwords and len(tok.encode("utf-8")) >= 2:
                out.append(tok)
            cur.clear()

    for ch in text:
        if ch < "\x80" and ch.isalnum() or ch >= "\x80" and _is_alnum(ch):
            cur.append(ch)
        else:
            emit()
    emit()
    return out


class TextIndex:
    """Text-level wrapper that feeds the C oracle: builds the String -> id dictionary in first-seen order."""

    def __init__(self, k1=1.2, b=0.75):
        self.k1, self.b = k1, b
        self.dict, self.docs = {}, []

    def add(self, text: str):
        ids = []
        for t in tokenize(text):
            ids.append(self.dict.setdefault(t, len(self.dict)))
        self.docs.append(ids)

    def build(self):
        from . import oracle as O
        return O.BM25(self.docs, max(len(self.dict), 1), self.k1, self.b)

    def query_ids(self, text: str):
        return [self.dict.get(t, 0xFFFFFFFF) for t in tokenize(text)]
