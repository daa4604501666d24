"""Text -> embedding: HuggingFace encoder forward (torch) + K1 fused pool/normalise.

Replaces txtai's vectors pipeline under Embeddings.index / Embeddings.search
(heavy_ranker.py:86,88,98,100): tokenizer(padding=True, truncation=True) ->
AutoModel forward -> MeanPooling -> normalize.  The encoder forward itself is the
model's own (out of scope); everything after the last hidden state is
``libvqa_b200.so``'s ``vqa_pool_normalize`` (K1).  Inputs are sorted by length and
batched by ``encodebatch`` (32), order restored, as txtai does.

Precision: the encoder forward runs in **float32** by default, like the reference's
(txtai / sentence-transformers load the checkpoint as stored and never down-cast).  A
bf16 forward moves the embeddings by ~1e-2 relative -- enough to flip a top-1 hit or the
``score_a + score_b > 0.4`` agreement rule (heavy_ranker.py:110) -- so it is opt-in:
``Embeddings(..., encoderdtype="bf16")``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import ops


class HFEncoder:
    normalized = True  # output is already pooled + L2-normalised on device

    def __init__(self, path: str, device=None, batch: int = 32, maxlength: Optional[int] = None,
                 model=None, tokenizer=None, dtype: torch.dtype = torch.float32):
        from . import _native

        _native.require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if model is None or tokenizer is None:
            from transformers import AutoModel, AutoTokenizer

            tokenizer = tokenizer or AutoTokenizer.from_pretrained(path)
            model = model or AutoModel.from_pretrained(path)
        self.tokenizer = tokenizer
        if isinstance(dtype, str):
            dtype = ops.DTYPES[dtype.lower()]
        self.dtype = dtype
        self.model = model.to(self.device, dtype=dtype).eval()
        self.batch = int(batch)
        self.maxlength = maxlength

    @torch.no_grad()
    def hidden_states(self, texts: Sequence[str]):
        enc = self.tokenizer(list(texts), padding=True, truncation=True, max_length=self.maxlength,
                             return_tensors="pt")
        enc = {k: v.to(self.device) for k, v in enc.items()}
        out = self.model(**enc)
        return out[0], enc["attention_mask"]

    def __call__(self, texts: Sequence[str]) -> torch.Tensor:
        texts = list(texts)
        order = sorted(range(len(texts)), key=lambda i: len(texts[i]))
        out: List[Optional[torch.Tensor]] = [None] * len(texts)
        for i in range(0, len(order), self.batch):
            idx = order[i:i + self.batch]
            hidden, mask = self.hidden_states([texts[j] for j in idx])
            emb = ops.pool_normalize(hidden, mask, normalize=True)
            for row, j in enumerate(idx):
                out[j] = emb[row]
        return torch.stack(out)  # type: ignore[arg-type]
