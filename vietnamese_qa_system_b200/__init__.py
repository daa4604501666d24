"""B200-native dense-retrieval engine for vTuanpham/Vietnamese_QA_System's retriever path.

``import vietnamese_qa_system_b200 as txtai`` gives the names the reference uses at
inference_pipeline/db_utils/heavy_ranker.py:78-101 (``Embeddings``); the sqlite helper
names of ``setup_db.py`` live in ``vietnamese_qa_system_b200.db`` and the corpus builder of
``setup_docs_db.py`` in ``vietnamese_qa_system_b200.corpus``.  All arithmetic is in
the in-tree CUDA library ``libvqa_b200.so`` (sm_100a); nothing here runs on the CPU.
"""
from . import _native
from .embeddings import Embeddings
from .ann import B200Flat
from .ops import FlatShard, agree, merge_topk, normalize_rows, pool_normalize
from .sharded import ShardedFlat, ShardedSearch, shard_bounds
from .ranker import HeavyRanker, load_passages, straighten_docs
from .scoring import BM25, Tokenizer
from . import corpus

__version__ = "0.1.0"
__all__ = ["Embeddings", "B200Flat", "FlatShard", "ShardedFlat", "ShardedSearch", "shard_bounds", "HeavyRanker",
           "straighten_docs", "load_passages", "BM25", "Tokenizer", "corpus", "agree", "merge_topk", "normalize_rows",
           "pool_normalize", "_native"]
