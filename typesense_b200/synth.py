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


def make_vectors_clustered(n: int, dim: int, n_clusters: int, seed: int, device="cpu", spread: float = 0.7,
                            latent: int = 8, centers_seed: Optional[int] = None, center_latent: int = 16):
    """Synthetic embeddings with the structure real ones have — low intrinsic dimension at every scale: cluster centres
    live in a `center_latent`-dimensional subspace, members vary along a further `latent`-dimensional subspace, and
    members are scattered over seq_ids at random. (i.i.d. Gaussian vectors in 768-d have no neighbourhood structure:
    no ANN index, hnswlib included, retrieves meaningful neighbours from them, so recall would say nothing.)
    x = normalize(A u_c + spread * B z / sqrt(latent)), u_c ~ N(0, I)/sqrt(center_latent), z ~ N(0, I).
    Returns (vectors, cluster key) where the key orders clusters along the first latent axis (a locality hint for the
    bulk graph builder). Queries are drawn with the same centers_seed and a different seed."""
    gc = torch.Generator(device=device)
    gc.manual_seed(seed if centers_seed is None else centers_seed)
    basis = torch.linalg.qr(torch.randn(dim, center_latent + latent, generator=gc, device=device, dtype=torch.float32))[0]
    A, B = basis[:, :center_latent], basis[:, center_latent:]
    u = torch.randn(n_clusters, center_latent, generator=gc, device=device, dtype=torch.float32)
    u = u / u.norm(dim=1, keepdim=True)
    centers = u @ A.T                                                   # unit vectors
    rank_of = torch.empty(n_clusters, dtype=torch.int64, device=device)
    rank_of[torch.argsort(u[:, 0])] = torch.arange(n_clusters, device=device)
    g = torch.Generator(device=device)
    g.manual_seed(seed + 7919)
    cid = torch.randint(0, n_clusters, (n,), generator=g, device=device)
    v = torch.empty(n, dim, device=device, dtype=torch.float32)
    step = 1 << 20
    for s0 in range(0, n, step):
        e = min(n, s0 + step)
        z = torch.randn(e - s0, latent, generator=g, device=device, dtype=torch.float32)
        x = centers[cid[s0:e]] + spread * (z @ B.T) / latent ** 0.5
        v[s0:e] = x / x.norm(dim=1, keepdim=True)
    return v, rank_of[cid]


# ----------------------------------------------------------------------------------------------------------------
# Bulk graph builder for bench-scale vector sets (10 M x 768 cannot be inserted point by point on the CPU within a
# bench run: hnswlib-style construction is hours at that size). This is HARNESS code for the write path, which is out
# of scope: it produces an HNSW-SHAPED graph (same level distribution, 2M links at level 0, M above,
# entry point at the top level) from exact local kNN inside overlapping windows of a locality-preserving order (the
# generating cluster when known, a random projection otherwise); small upper levels get an exact global kNN graph.
# The CPU oracle and the CUDA path traverse the SAME exported graph, so parity and the CPU/GPU ratio do not depend on
# how it was built; recall against brute force is reported next to every number that uses it.
def _local_knn_links(vec: torch.Tensor, ids: torch.Tensor, m: int, order: torch.Tensor, chunk: int, shift: int,
                     budget_elems: int = 1 << 29, n_random: int = 0, gen=None) -> torch.Tensor:
    """For the node subset `ids` visited in `order` (a permutation of range(len(ids))), the m nearest (inner product)
    neighbours inside windows of `chunk` consecutive nodes, windows offset by `shift`. Returns [len(ids), m] global ids."""
    dev = vec.device
    n = ids.numel()
    out = torch.empty(n, m, dtype=torch.int64, device=dev)
    if n <= m:
        for j in range(m):
            out[:, j] = ids[(torch.arange(n, device=dev) + 1 + j % max(n - 1, 1)) % n]
        return out
    c = min(chunk, n)
    n_chunks = (n - shift + c - 1) // c + (1 if shift else 0)
    starts = (torch.arange(n_chunks, device=dev) * c - (c - shift if shift else 0)).clamp_(min=0, max=n - c)
    ar = torch.arange(c, device=dev)
    batch_chunks = max(1, budget_elems // (c * c))
    for b0 in range(0, n_chunks, batch_chunks):
        st = starts[b0:b0 + batch_chunks]
        loc = order[(st[:, None] + ar[None, :])]          # [B, c] positions in ids
        gid = ids[loc]                                    # [B, c] global ids
        x = vec[gid].to(torch.bfloat16 if vec.is_cuda else torch.float32)
        sims = torch.bmm(x, x.transpose(1, 2)).float()
        sims.diagonal(dim1=1, dim2=2).fill_(float("-inf"))
        nb = torch.topk(sims, m, dim=2).indices           # [B, c, m] local
        if n_random:                                      # small-world shortcuts: random members of the window
            rnd = torch.randint(0, c, (nb.shape[0], c, n_random), generator=gen, device=dev)
            nb[:, :, m - n_random:] = rnd
        out[loc.reshape(-1)] = torch.gather(gid[:, None, :].expand(-1, c, -1), 2, nb).reshape(-1, m)
        del x, sims, nb
    return out


def build_graph_bulk(vec: torch.Tensor, M: int = 16, seed: int = 100, max_level_cap: int = 6,
                     order_key: Optional[torch.Tensor] = None, n_random_links: int = 0):
    """Returns (levels u8 [n], links0 i32 [n*(2M+1)], upper_off i64 [n+1], links_up i32 [R*(M+1)], max_level, entry).
    order_key: per-node locality key (e.g. generating cluster id); None -> a random projection."""
    dev = vec.device
    n, d = vec.shape
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    u = torch.rand(n, generator=gen, device=dev, dtype=torch.float64).clamp_(min=1e-300)
    levels = torch.floor(-torch.log(u) * (1.0 / np.log(M))).clamp_(max=max_level_cap).to(torch.int64)
    if order_key is None:
        proj = torch.randn(d, generator=gen, device=dev, dtype=torch.float32)
        key = torch.empty(n, device=dev, dtype=torch.float64)
        step = 1 << 20
        for s0 in range(0, n, step):
            key[s0:s0 + step] = (vec[s0:s0 + step] @ proj).double()
    else:
        key = order_key.double() + torch.rand(n, generator=gen, device=dev, dtype=torch.float64) * 0.999
    all_ids = torch.arange(n, device=dev)

    def merged_rows(ids, m_total, chunk):
        order = torch.argsort(key[ids])
        nn = ids.numel()
        if nn <= 65536:                                   # small level: exact global kNN graph
            rows = _local_knn_links(vec, ids, m_total, order, nn, 0)
        else:
            a = _local_knn_links(vec, ids, m_total - m_total // 2, order, chunk, 0)
            b = _local_knn_links(vec, ids, m_total // 2 + m_total // 4, order, chunk, chunk // 2,
                                 n_random=n_random_links if nn == n else 0, gen=gen)
            rows = torch.cat([a, b], dim=1)
        rows, _ = torch.sort(rows, dim=1)
        dup = torch.zeros_like(rows, dtype=torch.bool)
        dup[:, 1:] = rows[:, 1:] == rows[:, :-1]
        rows = torch.where(dup, torch.full_like(rows, n), rows)
        rows, _ = torch.sort(rows, dim=1)
        rows = rows[:, :m_total]
        cnt = (rows < n).sum(1)
        rows = torch.where(rows < n, rows, torch.zeros_like(rows))
        return rows, cnt

    rows, cnt = merged_rows(all_ids, 2 * M, 4096)
    links0 = torch.zeros(n, 2 * M + 1, dtype=torch.int32, device=dev)
    links0[:, 0] = cnt.to(torch.int32)
    links0[:, 1:] = rows.to(torch.int32)
    del rows, cnt
    max_level = int(levels.max())
    upper_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    upper_off[1:] = torch.cumsum(levels, 0)
    R = int(upper_off[-1])
    links_up = torch.zeros(max(R, 1), M + 1, dtype=torch.int32, device=dev)
    for l in range(1, max_level + 1):
        ids = torch.nonzero(levels >= l).flatten()
        rows, cnt = merged_rows(ids, M, 8192)
        rec = upper_off[ids] + (l - 1)
        links_up[rec, 0] = cnt.to(torch.int32)
        links_up[rec, 1:] = rows.to(torch.int32)
    entry = int(torch.nonzero(levels == max_level).flatten()[0])
    return (levels.to(torch.uint8), links0.reshape(-1), upper_off, links_up.reshape(-1), max_level, entry)


# ----------------------------------------------------------------------------------------------------------------
# Vocabulary strings for the token ids (rank r -> a 7-letter word; a bijection, so no two ranks share a word, and scrambled so
# that neighbouring ranks are not lexical neighbours). 26^7 = 8e9 words for <= 2^21 tokens: as in a natural vocabulary an
# edit-distance-1 neighbour of a word is almost never a word itself, so a misspelt query token has one or two candidates.
_W_LEN, _W_MOD = 7, 26 ** 7


def vocab_word(rank: int) -> bytes:
    x = (rank * 2654435761 + 97) % _W_MOD          # odd multiplier, gcd(2654435761, 26^7) == 1: a bijection on [0, 26^7)
    out = bytearray(_W_LEN)
    for i in range(_W_LEN):
        out[_W_LEN - 1 - i] = 97 + x % 26
        x //= 26
    return bytes(out)


def vocab_words(vocab: int):
    x = (np.arange(vocab, dtype=np.uint64) * np.uint64(2654435761) + np.uint64(97)) % np.uint64(_W_MOD)
    m = np.zeros((vocab, _W_LEN), np.uint8)
    for i in range(_W_LEN):
        m[:, _W_LEN - 1 - i] = (x % np.uint64(26)).astype(np.uint8) + 97
        x //= np.uint64(26)
    return [bytes(r) for r in m]


def misspell(word: bytes, rng, taken=None) -> bytes:
    """One substituted letter (an edit-distance-1 typo), never itself a vocabulary word when `taken` (a set) is given."""
    while True:
        i = int(rng.integers(0, len(word)))
        c = 97 + int(rng.integers(0, 26))
        if c == word[i]:
            continue
        w = word[:i] + bytes([c]) + word[i + 1:]
        if taken is None or w not in taken:
            return w
