"""Seeded synthetic collections in the shape SURVEY.md §8(d) names (harness code: tests + bench.py).

Everything is generated with torch ops so the same code runs on the CPU (tests) and on the GPU (bench, 10 M docs);
the result is the flattened posting form that both tsgpu_index_load_field and the CPU oracle consume, i.e. what a
mirror of Index::index_field_in_memory (src/index.cpp:700) would export: per token an ascending seq_id list with the
reference's offset encoding (src/index.cpp:1323-1395).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np
import torch

from .structs import FlatField


def zipf_cdf(vocab: int, s: float, device) -> torch.Tensor:
    w = torch.arange(1, vocab + 1, dtype=torch.float64, device=device).pow(-s)
    return torch.cumsum(w / w.sum(), 0)


@dataclass
class FieldData:
    flat: FlatField
    # per-doc token stream (CSR) kept for query generation: tokens of doc d = doc_tok[doc_off[d]:doc_off[d+1]]
    doc_off: np.ndarray
    doc_tok: np.ndarray


def _flatten(tok: torch.Tensor, doc: torch.Tensor, poskey: torch.Tensor, vocab: int) -> Tuple[np.ndarray, ...]:
    """Sort (token, doc, poskey) triples and cut them into lists / postings. poskey 0xFFFF encodes the trailing 0."""
    key = (tok.to(torch.int64) << 42) | (doc.to(torch.int64) << 16) | poskey.to(torch.int64)
    key, _ = torch.sort(key)
    raw = (key & 0xFFFF)
    raw = torch.where(raw == 0xFFFF, torch.zeros_like(raw), raw).to(torch.int32)
    td = key >> 16
    new_post = torch.ones_like(td, dtype=torch.bool)
    new_post[1:] = td[1:] != td[:-1]
    post_start = torch.nonzero(new_post).flatten()
    ids = ((td[post_start]) & ((1 << 26) - 1)).to(torch.int32)
    ptok = (td[post_start] >> 26)
    counts = torch.bincount(ptok, minlength=vocab)
    list_off = torch.zeros(vocab + 1, dtype=torch.int64, device=tok.device)
    list_off[1:] = torch.cumsum(counts, 0)
    pos_off = torch.cat([post_start, torch.tensor([key.numel()], device=tok.device, dtype=post_start.dtype)])
    return (list_off.cpu().numpy().astype(np.uint64), ids.cpu().numpy().astype(np.uint32),
            pos_off.cpu().numpy().astype(np.uint64), raw.cpu().numpy().astype(np.uint32))


def make_string_field(n_docs: int, vocab: int, len_lo: int, len_hi: int, seed: int, s: float = 1.07,
                      device="cpu") -> FieldData:
    """Plain string field: positions 1-based, the doc's last token gets a trailing 0 (src/index.cpp:1341-1348)."""
    assert n_docs < (1 << 26) and vocab < (1 << 21)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    lens = torch.randint(len_lo, len_hi + 1, (n_docs,), generator=g, device=device)
    off = torch.zeros(n_docs + 1, dtype=torch.int64, device=device)
    off[1:] = torch.cumsum(lens, 0)
    N = int(off[-1])
    doc = torch.repeat_interleave(torch.arange(n_docs, device=device), lens)
    pos = torch.arange(N, device=device) - off[doc] + 1
    cdf = zipf_cdf(vocab, s, device)
    u = torch.rand(N, generator=g, device=device, dtype=torch.float64)
    tok = torch.searchsorted(cdf, u).clamp_(max=vocab - 1)
    last = off[1:] - 1                                    # index of each doc's last token
    tok_all = torch.cat([tok, tok[last]])
    doc_all = torch.cat([doc, torch.arange(n_docs, device=device)])
    pos_all = torch.cat([pos, torch.full((n_docs,), 0xFFFF, device=device, dtype=pos.dtype)])
    list_off, ids, pos_off, raw = _flatten(tok_all, doc_all, pos_all, vocab)
    return FieldData(FlatField(list_off, ids, pos_off, raw, False), off.cpu().numpy(), tok.cpu().numpy().astype(np.uint32))


def make_array_field(n_docs: int, vocab: int, elems_lo: int, elems_hi: int, len_lo: int, len_hi: int, seed: int,
                     s: float = 1.07) -> FieldData:
    """string[] field (python loop: tests only). Encoding per src/index.cpp:1357-1393: per element and token
    positions..., last position repeated, array index; the element's last token also gets 0."""
    rng = np.random.default_rng(seed)
    w = np.arange(1, vocab + 1, dtype=np.float64) ** (-s)
    cdf = np.cumsum(w / w.sum())
    per_tok: List[List[Tuple[int, List[int]]]] = [[] for _ in range(vocab)]
    doc_off = [0]
    doc_tok: List[int] = []
    for d in range(n_docs):
        t2o = {}
        for ai in range(int(rng.integers(elems_lo, elems_hi + 1))):
            L = int(rng.integers(len_lo, len_hi + 1))
            toks = np.minimum(np.searchsorted(cdf, rng.random(L)), vocab - 1)
            seen = []
            for p, t in enumerate(toks):
                t = int(t)
                t2o.setdefault(t, []).append(p + 1)
                if t not in seen:
                    seen.append(t)
                doc_tok.append(t)
            for t in sorted(seen):
                t2o[t].append(t2o[t][-1])
                t2o[t].append(ai)
            t2o[int(toks[-1])].append(0)
        doc_off.append(len(doc_tok))
        for t, offs in t2o.items():
            per_tok[t].append((d, offs))
    return FieldData(FlatField.from_postings(per_tok, True), np.asarray(doc_off, np.int64), np.asarray(doc_tok, np.uint32))


def make_points(n_docs: int, seed: int, hi: int = 1_000_000, missing_frac: float = 0.0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    v = rng.integers(0, hi, n_docs, dtype=np.int64)
    if missing_frac > 0:
        v[rng.random(n_docs) < missing_frac] = np.iinfo(np.int64).min
    return v


def make_vectors(n: int, dim: int, seed: int, device="cpu", normalize: bool = True) -> torch.Tensor:
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    v = torch.randn(n, dim, generator=g, device=device, dtype=torch.float32)
    if normalize:
        v = v / (v.norm(dim=1, keepdim=True) + 1e-30)
    return v


def sample_queries(fd: FieldData, n_queries: int, n_terms: int, seed: int) -> np.ndarray:
    """Pick a doc and take n_terms of its tokens (SURVEY §8d cfg 2). Returns [n_queries, n_terms] token ids."""
    rng = np.random.default_rng(seed)
    n_docs = len(fd.doc_off) - 1
    out = np.zeros((n_queries, n_terms), np.uint32)
    for i in range(n_queries):
        while True:
            d = int(rng.integers(0, n_docs))
            a, b = int(fd.doc_off[d]), int(fd.doc_off[d + 1])
            if b - a >= n_terms:
                break
        sel = np.sort(rng.choice(b - a, n_terms, replace=False))
        out[i] = fd.doc_tok[a + sel]
    return out
