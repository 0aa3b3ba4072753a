"""tsgpu_index_build_hnsw (csrc/hnsw_build.cuh) on the GPU: hnswlib's addPoint as batched rounds.

Parity: construction is UNPINNED against the reference itself (it inserts with 4 threads, src/index.cpp:1009, so no two
reference builds agree); the anchor is the oracle's single-threaded restatement of hnswlib's addPoint. With one node per
round the device build must reproduce that graph link for link — same levels, same rows in the same order, same entry
point. Batched rounds relax the order exactly like a multi-threaded hnswlib build: there the checks are structural
invariants, search parity against the oracle on the exported graph, and recall against brute force."""
import numpy as np
import pytest

import oracle_lib as ol
from typesense_b200 import capi, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,dim,M,efc", [(700, 32, 8, 40), (500, 128, 16, 60), (300, 50, 4, 30)])
def test_sequential_device_build_equals_oracle_build(n, dim, M, efc):
    vec = synth.make_vectors_clustered(n, dim, 12, seed=n, spread=0.6, latent=6, center_latent=6)[0].numpy()
    ref = ol.hnsw_build(vec, M, efc, 100)
    gi = capi.GpuIndex(n, 0)
    info = gi.build_hnsw(vec, M, efc, 100, max_batch=1)
    g = gi.export_hnsw(vec)
    assert info["max_level"] == ref.max_level and info["entry_point"] == ref.entry_point
    assert (g.levels == ref.levels).all() and (g.upper_off == ref.upper_off).all()
    L0 = 2 * M + 1
    a, b = g.links0.reshape(n, L0), ref.links0.reshape(n, L0)
    assert (a[:, 0] == b[:, 0]).all(), np.nonzero(a[:, 0] != b[:, 0])[0][:10]
    for i in range(n):                                   # rows beyond the count are scratch in both builds
        assert (a[i, 1:1 + a[i, 0]] == b[i, 1:1 + b[i, 0]]).all(), (i, a[i], b[i])
    if info["n_upper"]:
        au, bu = g.links_up.reshape(-1, M + 1), ref.links_up.reshape(-1, M + 1)
        assert (au[:, 0] == bu[:, 0]).all()
        for i in range(len(au)):
            assert (au[i, 1:1 + au[i, 0]] == bu[i, 1:1 + bu[i, 0]]).all(), (i, au[i], bu[i])
    gi.close()


def test_batched_device_build_invariants_recall_and_search_parity():
    n, dim, M, efc = 30000, 64, 16, 100
    vec = synth.make_vectors_clustered(n, dim, 40, seed=3, spread=0.5, latent=8, center_latent=8)[0].numpy()
    gi = capi.GpuIndex(n, 0)
    info = gi.build_hnsw(vec, M, efc, 100, max_batch=1024)
    assert info["build"]["rounds"] < n // 8
    g = gi.export_hnsw(vec)
    L0 = 2 * M + 1
    rows = g.links0.reshape(n, L0)
    cnt = rows[:, 0]
    assert cnt.max() <= 2 * M and cnt[1:].min() >= 1
    for i in range(0, n, 97):
        r = rows[i, 1:1 + cnt[i]]
        assert (r < n).all() and i not in r and len(set(r.tolist())) == len(r)
    up = g.links_up.reshape(-1, M + 1)
    assert up[:, 0].max() <= M
    # the library searches the graph it built; the oracle walks the exported copy: identical answers
    oi = ol.OracleIndex(n, [], [], g)
    qv = synth.make_vectors_clustered(200, dim, 40, seed=91, spread=0.5, latent=8, center_latent=8, centers_seed=3)[0].numpy()
    d, l, cn = gi.knn(qv, 10, 100)
    od, olab, ocn, _ = oi.knn(qv, 10, 100)
    assert cn.tolist() == ocn.tolist() and l.tolist() == olab.tolist() and (d == od).all()
    exact = np.argsort(-(qv @ vec.T), axis=1)[:, :10]
    recall = np.mean([len(set(l[i].tolist()) & set(exact[i].tolist())) / 10 for i in range(len(qv))])
    assert recall >= 0.95, recall
    # same data through the oracle's sequential build: the batched graph must search about as well
    ref = ol.hnsw_build(vec[:6000], M, efc, 100)
    gi2 = capi.GpuIndex(6000, 0)
    gi2.build_hnsw(vec[:6000], M, efc, 100, max_batch=256)
    g2 = gi2.export_hnsw(vec[:6000])
    ex2 = np.argsort(-(qv @ vec[:6000].T), axis=1)[:, :10]
    rec = []
    for graph in (ref, g2):
        o = ol.OracleIndex(6000, [], [], graph)
        _, ll, _, _ = o.knn(qv, 10, 50)
        rec.append(np.mean([len(set(ll[i].tolist()) & set(ex2[i].tolist())) / 10 for i in range(len(qv))]))
    assert rec[1] >= rec[0] - 0.02, rec
    gi.close(); gi2.close()


def _rows_equal(a, b, n, width):
    a, b = a.reshape(-1, width)[:n], b.reshape(-1, width)[:n]
    if not (a[:, 0] == b[:, 0]).all():
        return False
    return all((a[i, 1:1 + a[i, 0]] == b[i, 1:1 + b[i, 0]]).all() for i in range(len(a)))


def test_append_continues_the_sequential_build_link_for_link():
    """SURVEY 8 f-4, vector half: build(n1) + append(n2) with one node per round == build(n1 + n2) == the oracle's sequential build."""
    n, n1, dim, M, efc = 900, 520, 64, 8, 40
    vec = synth.make_vectors_clustered(n, dim, 10, seed=5, spread=0.6, latent=6, center_latent=6)[0].numpy()
    ref = ol.hnsw_build(vec, M, efc, 100)
    gi = capi.GpuIndex(n, 0)
    gi.build_hnsw(vec[:n1], M, efc, 100, max_batch=1)
    info = gi.append_hnsw(vec[n1:], efc, 100, max_batch=1)
    assert info["n"] == n and info["max_level"] == ref.max_level and info["entry_point"] == ref.entry_point
    g = gi.export_hnsw(vec)
    assert (g.levels == ref.levels).all() and (g.upper_off == ref.upper_off).all()
    assert _rows_equal(g.links0, ref.links0, n, 2 * M + 1)
    if info["n_upper"]:
        assert _rows_equal(g.links_up, ref.links_up, info["n_upper"], M + 1)
    # the searches of the appended index see the new rows' vectors
    qv = vec[n1 + 5:n1 + 25] + 0.001
    d, l, cn = gi.knn(qv, 5, 50)
    od, olab, ocn, _ = ol.OracleIndex(n, [], [], ref).knn(qv, 5, 50)
    assert l.tolist() == olab.tolist() and (d == od).all()
    gi.close()


def test_batched_append_search_parity_and_recall_then_mark_deleted():
    n, n1, dim, M, efc = 20000, 12000, 64, 16, 100
    vec = synth.make_vectors_clustered(n, dim, 30, seed=8, spread=0.5, latent=8, center_latent=8)[0].numpy()
    gi = capi.GpuIndex(n, 0)
    gi.build_hnsw(vec[:n1], M, efc, 100, max_batch=512)
    gi.append_hnsw(vec[n1:n1 + 3000], efc, 100, max_batch=512)
    gi.append_hnsw(vec[n1 + 3000:], efc, 100, max_batch=512)
    g = gi.export_hnsw(vec)
    assert len(g.levels) == n
    oi = ol.OracleIndex(n, [], [], g)
    qv = synth.make_vectors_clustered(150, dim, 30, seed=77, spread=0.5, latent=8, center_latent=8, centers_seed=8)[0].numpy()
    d, l, cn = gi.knn(qv, 10, 100)
    od, olab, ocn, _ = oi.knn(qv, 10, 100)
    assert cn.tolist() == ocn.tolist() and l.tolist() == olab.tolist() and (d == od).all()
    exact = np.argsort(-(qv @ vec.T), axis=1)[:, :10]
    recall = np.mean([len(set(l[i].tolist()) & set(exact[i].tolist())) / 10 for i in range(len(qv))])
    assert recall >= 0.95, recall
    # markDelete: a deleted label is never returned but still traversed == the oracle's walk with those labels excluded
    dead = np.unique(np.concatenate([l[:40, :3].ravel(), np.arange(0, n, 50, dtype=np.uint32)])).astype(np.uint32)
    gi.mark_deleted(dead)
    d2, l2, cn2 = gi.knn(qv, 10, 100)
    alive = np.setdiff1d(np.arange(n, dtype=np.uint32), dead).astype(np.uint32)      # the same gate as a filter of the live labels
    od2, ol2, ocn2, _ = oi.knn(qv, 10, 100, np.zeros(len(qv), np.int32), [alive])
    assert cn2.tolist() == ocn2.tolist() and l2.tolist() == ol2.tolist() and (d2 == od2).all()
    assert not (set(l2[:, :].ravel().tolist()) & set(dead.tolist()))
    gi.mark_deleted(dead, deleted=False)
    d3, l3, _ = gi.knn(qv, 10, 100)
    assert l3.tolist() == l.tolist()
    gi.close()
