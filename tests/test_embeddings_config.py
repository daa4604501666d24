"""Host-side configuration of the drop-in Embeddings class that does not need a GPU."""
import inspect

import torch

from vietnamese_qa_system_b200 import vectors
from vietnamese_qa_system_b200.embeddings import Embeddings


def test_encoder_runs_in_fp32_unless_asked_otherwise(monkeypatch):
    """The reference (txtai / sentence-transformers) runs the encoder forward in fp32; a bf16 forward moves the
    embeddings by ~1e-2 relative and can flip the top-1 hit or the 0.4 agreement rule (heavy_ranker.py:110).  The
    drop-in constructor therefore defaults to fp32; `encoderdtype` opts in to 16-bit and is saved with the config."""
    assert inspect.signature(vectors.HFEncoder.__init__).parameters["dtype"].default is torch.float32
    seen = {}

    class Stub:
        def __init__(self, path, **kw):
            seen["path"], seen["kw"] = path, kw

    monkeypatch.setattr(vectors, "HFEncoder", Stub)
    e = Embeddings(hybrid=True, content=True, path="sentence-transformers/paraphrase-multilingual-mpnet-base-v2")
    e._encoder_fn()
    assert seen["kw"]["dtype"] == "fp32" and seen["path"].endswith("mpnet-base-v2")
    e2 = Embeddings(path="m", encoderdtype="bf16", encodebatch=8)
    e2._encoder_fn()
    assert seen["kw"]["dtype"] == "bf16" and seen["kw"]["batch"] == 8
    assert e2.config["encoderdtype"] == "bf16"            # part of the config -> written by save(), read by load()
