"""Corpus builder (mirror of setup_docs_db.py): the restated langchain 0.0.286 splitter against
hand-derived known answers and size/coverage properties; insert_doc against sqlite.  CPU only."""
import random

import pytest

from vietnamese_qa_system_b200 import corpus, db
from vietnamese_qa_system_b200.corpus import RecursiveCharacterTextSplitter as Splitter


def test_short_text_is_one_chunk():
    assert corpus.reference_splitter().split_text("Hà Nội là thủ đô.") == ["Hà Nội là thủ đô."]
    assert corpus.reference_splitter().split_text("") == []
    assert corpus.reference_splitter().split_text("   \n\n  ") == []


def test_keep_separator_glues_the_delimiter_to_the_following_piece():
    # pieces: "a", ".b", ".c" (lengths 1, 2, 2); chunk_size 3 -> "a.b" fits, ".c" starts a new chunk
    s = Splitter(separators=["."], chunk_size=3, chunk_overlap=0, keep_separator=True)
    assert s.split_text("a.b.c") == ["a.b", ".c"]
    # without keep_separator the delimiter is dropped at the split and re-inserted by the join
    s2 = Splitter(separators=["."], chunk_size=3, chunk_overlap=0, keep_separator=False)
    assert s2.split_text("a.b.c") == ["a.b", "c"]


def test_overlap_carries_trailing_pieces_into_the_next_chunk():
    # pieces " aa"-style words of 3 chars ("aa", " bb", " cc", " dd", " ee"): chunk_size 8, overlap 3
    s = Splitter(separators=[" "], chunk_size=8, chunk_overlap=3, keep_separator=True)
    # "aa bb cc" = 8 fits; adding " dd" overflows -> emit, keep the last piece (3 <= overlap) -> " cc dd ee" is 9 > 8,
    # so after " cc dd" (6) adding " ee" overflows -> emit "cc dd" (stripped), carry " dd"
    assert s.split_text("aa bb cc dd ee") == ["aa bb cc", "cc dd", "dd ee"]
    s0 = Splitter(separators=[" "], chunk_size=8, chunk_overlap=0, keep_separator=True)
    assert s0.split_text("aa bb cc dd ee") == ["aa bb cc", "dd ee"]


def test_recursion_falls_through_to_finer_separators():
    s = Splitter(separators=["\n\n", ".", " "], chunk_size=10, chunk_overlap=0, keep_separator=True)
    text = "one two three four.five\n\nsix"
    got = s.split_text(text)
    # paragraph 1 (23 chars) is too long -> split on "." -> "one two three four" (18) too long -> split on " ":
    # pieces "one", " two", " three", " four" (3, 4, 6, 5); lengths are measured BEFORE the strip, so
    # " three" + " four" = 11 does not fit 10
    assert got == ["one two", "three", "four", ".five", "six"]
    assert all(len(c) <= 10 for c in got)


def test_piece_without_any_separator_is_kept_whole():
    s = Splitter(separators=[" "], chunk_size=5, chunk_overlap=0)
    assert s.split_text("abcdefghij kl") == ["abcdefghij", "kl"]      # langchain emits the oversize piece as is


def test_overlap_larger_than_chunk_is_rejected():
    with pytest.raises(ValueError):
        Splitter(chunk_size=10, chunk_overlap=11)


def _article(rng, n_sent):
    words = ["Hà_Nội", "thủ_đô", "Việt_Nam", "là", "của", "thành_phố", "lịch_sử", "văn_hóa", "năm", "1010",
             "sông", "Hồng", "người", "dân", "trung_tâm", "kinh_tế", "chính_trị", "và", "một", "trong", "những"]
    out = []
    for i in range(n_sent):
        k = rng.randint(4, 30)
        sent = " ".join(rng.choice(words) for _ in range(k))
        out.append(sent + rng.choice([".", ".", "!", "?", ";", ","]))
        if rng.random() < 0.15:
            out.append("\n\n" if rng.random() < 0.5 else "\n")
        else:
            out.append(" ")
    return "".join(out)


def test_reference_configuration_properties():
    rng = random.Random(7)
    sp = corpus.reference_splitter()
    for _ in range(20):
        text = _article(rng, rng.randint(5, 120))
        chunks = sp.split_text(text)
        assert chunks and all(c and c == c.strip() for c in chunks)
        assert all(len(c) <= 512 for c in chunks)
        # keep_separator=True joins pieces with "", so every chunk is a contiguous slice of the article
        at, prev_end = 0, 0
        for c in chunks:
            pos = text.find(c, at)
            assert pos >= 0
            assert pos <= prev_end + 2 or text[prev_end:pos].strip() == ""     # nothing but whitespace is skipped
            if prev_end > pos:
                assert prev_end - pos <= 51.2 + 1                               # overlap budget (512 * 0.1)
            at, prev_end = pos + 1, pos + len(c)
        assert text[prev_end:].strip() == ""
        # only the final chunk of a run may be short: consecutive chunks cannot both fit one chunk without overlap
        assert corpus.chunk_corpus([text]) == [c.replace("_", " ") for c in chunks]


def test_insert_doc_builds_documents_db(tmp_path):
    rng = random.Random(3)
    texts = [_article(rng, 60) for _ in range(5)]
    path = str(tmp_path / "documents.db")
    docs = corpus.insert_doc(path, texts=texts)
    assert docs == corpus.chunk_corpus(texts) and len(docs) > 5 and not any("_" in d for d in docs)
    rows = db.query(path, "SELECT * FROM documents", fetch_size=50000)           # heavy_ranker.py:70-72
    assert [r[0] for r in rows] == list(range(1, len(docs) + 1))                 # AUTOINCREMENT ids, ascending
    assert [r[1] for r in rows] == docs and {r[2] for r in rows} == {corpus.REFERENCE_SOURCE}
    # re-running drops and recreates the table (setup_docs_db.py:42-50); max_examples truncates the articles
    docs2 = corpus.insert_doc(path, max_examples=2, texts=texts)
    rows2 = db.query(path, "SELECT * FROM documents", fetch_size="all")
    assert docs2 == corpus.chunk_corpus(texts[:2]) and [r[1] for r in rows2] == docs2
