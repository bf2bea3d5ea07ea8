"""Vocabulary-free stand-in for the HF tokenizer — TEST INFRASTRUCTURE ONLY.

No tokenizer files exist offline, so tests and golden-vector generation use this
deterministic word-hash tokenizer with the PubMedBERT conventions ([PAD]=0, [CLS]=2, [SEP]=3) and the same
call signature the reference uses (`tokenizer(texts, max_length=256, padding='max_length', truncation=True,
return_tensors='pt')`, quick_start/keep_inference.py:99, WSI_evaluation/utils.py:73)."""
from __future__ import annotations

import re
import zlib

import torch
from transformers import BatchEncoding


class FakeTokenizer:
    def __init__(self, vocab_size: int = 30522):
        self.vocab_size = vocab_size

    def _ids(self, text: str):
        words = re.findall(r"[a-z0-9]+|[^\sa-z0-9]", text.lower())
        return [5 + zlib.crc32(w.encode()) % (self.vocab_size - 5) for w in words]

    def __call__(self, texts, max_length=256, padding="max_length", truncation=True, return_tensors="pt"):
        if isinstance(texts, str):
            texts = [texts]
        rows = []
        for t in texts:
            ids = [2] + self._ids(t)[: max_length - 2] + [3]
            rows.append(ids)
        width = max_length if padding == "max_length" else max(len(r) for r in rows)
        input_ids = torch.zeros(len(rows), width, dtype=torch.long)
        mask = torch.zeros(len(rows), width, dtype=torch.long)
        for i, r in enumerate(rows):
            input_ids[i, : len(r)] = torch.tensor(r)
            mask[i, : len(r)] = 1
        return BatchEncoding({"input_ids": input_ids, "token_type_ids": torch.zeros_like(input_ids), "attention_mask": mask})
