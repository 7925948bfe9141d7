"""Oracle-side restatement of BM25Index::tokenize (reference src/index.rs:93-124).  TEST INFRASTRUCTURE ONLY.

split on `!char::is_alphanumeric`, drop empties, lowercase, drop the 90 stopwords, drop tokens whose UTF-8 byte
length is < 2.  Python's str.isalnum()/lower() agree with Rust's char::is_alphanumeric/to_lowercase on ASCII and on
the Latin/Greek/Cyrillic letters used in the tests; exotic code points may differ (documented in DESIGN.md)."""

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
        if ch.isalnum():
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
