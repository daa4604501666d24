"""Corpus builder -- mirror of the reference's ``inference_pipeline/db_utils/setup_docs_db.py``.

The reference fills ``documents.db`` by chunking a Vietnamese Wikipedia dump
(``insert_doc`` :16-50): ``RecursiveCharacterTextSplitter(chunk_size=512, chunk_overlap=51.2,
separators=["\\n\\n", "\\n", ".", ",", ";", "!", "?", " "], keep_separator=True)`` over each
article (:27-35), underscores of the word-segmented text replaced by spaces (:24-25,36), one row
``{"doc", "source"}`` per chunk (:38-40), table dropped / recreated / filled (:42-50).  Those
rows are what ``heavy_ranker.py:70-76`` reads back and indexes, so the chunker decides what a
"passage" is on the retrieval path.

The splitter itself is third-party (``langchain==0.0.286``, ``requirements.txt:14``; not
installable here): ``RecursiveCharacterTextSplitter`` below restates that release's algorithm
(``_split_text_with_regex`` / ``_split_text`` / ``_merge_splits`` / ``_join_docs``).  The
dataset (``EddieChen372/vietnamese-wiki-segmented``) needs the network, so ``insert_doc`` also
accepts the article texts directly.  Host-side string work; nothing here touches the GPU.
"""
from __future__ import annotations

import os
import re
from typing import Callable, Iterable, List, Optional, Sequence

from .db import drop_tables, insert_data, setup_database

REFERENCE_SEPARATORS = ["\n\n", "\n", ".", ",", ";", "!", "?", " "]        # setup_docs_db.py:32
REFERENCE_SOURCE = "EddieChen372/vietnamese-wiki-segmented"                 # setup_docs_db.py:18,39


def _split_text_with_regex(text: str, separator: str, keep_separator: bool) -> List[str]:
    if separator:
        if keep_separator:
            # the capture group keeps the delimiters; each one is glued to the FRONT of the piece after it
            parts = re.split(f"({separator})", text)
            splits = [parts[i] + parts[i + 1] for i in range(1, len(parts), 2)]
            if len(parts) % 2 == 0:
                splits += parts[-1:]
            splits = [parts[0]] + splits
        else:
            splits = re.split(separator, text)
    else:
        splits = list(text)
    return [s for s in splits if s != ""]


class RecursiveCharacterTextSplitter:
    """langchain 0.0.286 ``RecursiveCharacterTextSplitter`` (with ``TextSplitter``'s merge logic)."""

    def __init__(self, separators: Optional[Sequence[str]] = None, keep_separator: bool = True,
                 is_separator_regex: bool = False, chunk_size: int = 4000, chunk_overlap: float = 200,
                 length_function: Callable[[str], int] = len, add_start_index: bool = False,
                 strip_whitespace: bool = True):
        if chunk_overlap > chunk_size:
            raise ValueError(f"Got a larger chunk overlap ({chunk_overlap}) than chunk size ({chunk_size}), "
                             f"should be smaller.")
        self._separators = list(separators) if separators is not None else ["\n\n", "\n", " ", ""]
        self._keep_separator = keep_separator
        self._is_separator_regex = is_separator_regex
        self._chunk_size = chunk_size
        self._chunk_overlap = chunk_overlap
        self._length_function = length_function
        self._add_start_index = add_start_index
        self._strip_whitespace = strip_whitespace

    # -- TextSplitter ------------------------------------------------------------------------
    def _join_docs(self, docs: List[str], separator: str) -> Optional[str]:
        text = separator.join(docs)
        if self._strip_whitespace:
            text = text.strip()
        return None if text == "" else text

    def _merge_splits(self, splits: Iterable[str], separator: str) -> List[str]:
        separator_len = self._length_function(separator)
        docs: List[str] = []
        current: List[str] = []
        total = 0
        for d in splits:
            n = self._length_function(d)
            if total + n + (separator_len if len(current) > 0 else 0) > self._chunk_size:
                if len(current) > 0:
                    doc = self._join_docs(current, separator)
                    if doc is not None:
                        docs.append(doc)
                    # drop pieces from the front until the carried-over tail fits the overlap budget
                    # (or until the next piece fits the chunk)
                    while total > self._chunk_overlap or (
                            total + n + (separator_len if len(current) > 0 else 0) > self._chunk_size and total > 0):
                        total -= self._length_function(current[0]) + (separator_len if len(current) > 1 else 0)
                        current = current[1:]
            current.append(d)
            total += n + (separator_len if len(current) > 1 else 0)
        doc = self._join_docs(current, separator)
        if doc is not None:
            docs.append(doc)
        return docs

    # -- RecursiveCharacterTextSplitter ----------------------------------------------------------
    def _split_text(self, text: str, separators: List[str]) -> List[str]:
        final_chunks: List[str] = []
        separator = separators[-1]
        new_separators: List[str] = []
        for i, s in enumerate(separators):
            pattern = s if self._is_separator_regex else re.escape(s)
            if s == "":
                separator = s
                break
            if re.search(pattern, text):
                separator = s
                new_separators = separators[i + 1:]
                break
        pattern = separator if self._is_separator_regex else re.escape(separator)
        splits = _split_text_with_regex(text, pattern, self._keep_separator)
        good: List[str] = []
        joiner = "" if self._keep_separator else separator
        for s in splits:
            if self._length_function(s) < self._chunk_size:
                good.append(s)
            else:
                if good:
                    final_chunks.extend(self._merge_splits(good, joiner))
                    good = []
                if not new_separators:
                    final_chunks.append(s)
                else:
                    final_chunks.extend(self._split_text(s, new_separators))
        if good:
            final_chunks.extend(self._merge_splits(good, joiner))
        return final_chunks

    def split_text(self, text: str) -> List[str]:
        return self._split_text(text, self._separators)

    def create_documents(self, texts: Iterable[str]) -> List[str]:
        """langchain returns ``Document`` objects; the reference only reads ``page_content``
        (setup_docs_db.py:36), so the chunks themselves are returned."""
        out: List[str] = []
        for text in texts:
            out.extend(self.split_text(text))
        return out


def reference_splitter() -> RecursiveCharacterTextSplitter:
    """The splitter exactly as setup_docs_db.py:27-34 configures it."""
    return RecursiveCharacterTextSplitter(chunk_size=512, chunk_overlap=512 * 0.1, length_function=len,
                                          add_start_index=False, separators=REFERENCE_SEPARATORS,
                                          keep_separator=True)


def rm_underscore(data: str) -> str:
    """setup_docs_db.py:24-25: the dump is word-segmented with ``_`` inside compounds."""
    return re.sub("_", " ", data)


def chunk_corpus(texts: Iterable[str], splitter: Optional[RecursiveCharacterTextSplitter] = None) -> List[str]:
    """Articles -> passages (setup_docs_db.py:27-36)."""
    splitter = splitter or reference_splitter()
    return [rm_underscore(c) for c in splitter.create_documents(texts)]


def insert_doc(database_path: str, max_examples: int = 50000, texts: Optional[Sequence[str]] = None,
               database_dir: Optional[str] = None, source: str = REFERENCE_SOURCE, verbose: bool = False) -> List[str]:
    """Build ``documents.db`` (setup_docs_db.py:16-52).

    ``database_path``: the ``.db`` file to (re)create -- the reference ignores its own argument and always
    writes ``inference_pipeline/dbs/documents.db`` (:42-50); here the argument is honoured.  ``texts``: the
    article texts (``ctx_wiki_dataset['segmented_text']``); when omitted they are loaded with
    ``datasets.load_dataset`` as the reference does, which needs the network.  Returns the passages."""
    if texts is None:
        from datasets import load_dataset  # network; reference :18-19

        texts = load_dataset(source, split="train")[:max_examples]["segmented_text"]
    texts = list(texts)[:max_examples]
    docs = chunk_corpus(texts)
    data_to_insert = [{"doc": doc, "source": source} for doc in docs]
    database_dir = database_dir or os.path.dirname(os.path.abspath(database_path))
    name = os.path.splitext(os.path.basename(database_path))[0]
    if os.path.isfile(database_path):
        drop_tables(database_path, tables_to_drop=["documents"], verbose=verbose)
    created = setup_database(name, table_names=["documents"],
                             fields=['''(id INTEGER PRIMARY KEY AUTOINCREMENT, doc TEXT, source TEXT)'''],
                             database_dir=database_dir, verbose=verbose)
    if data_to_insert:
        insert_data(created, table_name="documents", data=data_to_insert, verbose=verbose)
    return docs
