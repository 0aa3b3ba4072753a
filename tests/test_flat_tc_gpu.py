"""K7 on the tensor cores (csrc/flat_tc.cu): the batched flat scan — process_results_bruteforce (src/index.cpp:3345-3374) for queries
that share one candidate set — against the oracle's fp32 loop (tso_flat_distances). The bar for float distances is 1e-4 relative
(BASELINE.json north_star); measured deviations are ~1e-7 absolute, the class of a re-ordered fp32 sum."""
import os

import numpy as np
import pytest

import oracle_lib as ol
from typesense_b200 import capi, structs as S, synth

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-4, 2e-6
DEVICE = os.environ.get("TSGPU_TEST_DOUBLE") != "1"        # the dry run against the oracle double has no device counters


def _unit(n, dim, seed):
    return synth.make_vectors(n, dim, seed).numpy()


def _empty_graph(vec):
    n = vec.shape[0]
    return S.HnswGraph(vec, np.zeros(n, np.uint8), np.zeros((n, 33), np.uint32).reshape(-1), np.zeros(n + 1, np.uint64), np.zeros(1, np.uint32), 16, 0, 0)


@pytest.mark.parametrize("dim", [96, 128, 768])
def test_flat_distances_batch_vs_oracle(dim):
    n = 6000
    vec = _unit(n, dim, 5)
    g = _empty_graph(vec)
    gi = capi.GpuIndex(n, 0)
    gi.load_hnsw(g)
    oi = ol.OracleIndex(n, [], [], g)
    rng = np.random.default_rng(dim)
    for nq, n_ids in [(8, 128), (37, 1000), (205, 3000), (300, 2001)]:
        q = _unit(nq, dim, 100 + nq)
        q[1::2] = vec[rng.integers(0, n, len(q[1::2]))] + 0.02 * rng.standard_normal((len(q[1::2]), dim)).astype(np.float32)     # near-duplicates: small distances
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        ids = np.sort(rng.choice(n, n_ids, replace=False)).astype(np.uint32)
        d = gi.flat_distances_batch(q, ids)
        assert not DEVICE or gi.stats()["flat_tc_queries"] == nq, "the tensor-core scan did not run"
        ref = np.stack([oi.flat_distances(q[i], ids) for i in range(nq)])
        assert d.shape == ref.shape
        assert np.allclose(d, ref, rtol=RTOL, atol=ATOL), f"nq={nq}: max abs dev {np.abs(d - ref).max():.3g}"
        # and against the scalar device path (bit-equal to the oracle): the same numbers within the tolerance
        d0 = gi.flat_distances(q[0], ids)
        assert np.allclose(d[0], d0, rtol=RTOL, atol=ATOL)
    gi.close()


def test_unsupported_dim_takes_the_scalar_path():
    n, dim = 2000, 50                                # not a multiple of 32: pair-by-pair kernel, bit-equal to the oracle
    vec = _unit(n, dim, 9)
    g = _empty_graph(vec)
    gi = capi.GpuIndex(n, 0)
    gi.load_hnsw(g)
    oi = ol.OracleIndex(n, [], [], g)
    q = _unit(20, dim, 10)
    ids = np.arange(0, n, 3, dtype=np.uint32)
    d = gi.flat_distances_batch(q, ids)
    assert gi.stats()["flat_tc_queries"] == 0
    ref = np.stack([oi.flat_distances(q[i], ids) for i in range(len(q))])
    assert (d == ref).all()
    gi.close()


def test_vector_search_flat_groups_on_tensor_cores():
    """tsgpu_vector_search_batch with TSGPU_VEC_FLAT_TENSOR: flat-path queries sharing a filter go through the tensor-core scan;
    the results are the fp32 path's — same ids (up to swaps of candidates closer than the tolerance), distances within 1e-4."""
    n, dim = 8000, 128
    vec = _unit(n, dim, 21)
    graph = ol.hnsw_build(vec, 16, 100, 100)
    fd = synth.make_string_field(n, 300, 3, 8, seed=3)
    pts = synth.make_points(n, 4)
    gi = capi.GpuIndex(n, 0)
    gi.load_field(fd.flat); gi.load_sort_column(pts); gi.load_hnsw(graph)
    rng = np.random.default_rng(77)
    filters = [np.unique(rng.integers(0, n, 900)).astype(np.uint32), np.arange(0, n, 40, dtype=np.uint32), np.arange(5, 60, dtype=np.uint32)]
    sort = ((S.SORT_VECTOR_DISTANCE, -1, -1, 0), (S.SORT_NUMERIC, 0, 1, 0), (S.SORT_NONE, -1, 1, 0))
    nq = 90
    qs = []
    for i in range(nq):
        q = S.Query([], topk=250, sort=sort)
        q.filter = i % 3
        qs.append(q)
    b = S.KwBatch(qs, [0], filters)
    qv = _unit(nq, dim, 23)
    kv0, cnt0, found0 = gi.vector_search(b, qv, S.vec_params(k=50, ef=80, flat_search_cutoff=2000, fetch_size=10), 256)
    assert gi.stats()["flat_tc_queries"] == 0
    kv1, cnt1, found1 = gi.vector_search(b, qv, S.vec_params(k=50, ef=80, flat_search_cutoff=2000, fetch_size=10, flags=S.VEC_FLAT_TENSOR), 256)
    assert not DEVICE or gi.stats()["flat_tc_queries"] == 60           # the two filters with >= 128 ids, 30 queries each; the 55-id filter stays scalar
    assert cnt0.tolist() == cnt1.tolist() and found0.tolist() == found1.tolist()
    same_order = 0
    for q in range(nq):
        c = int(cnt0[q])
        assert sorted(kv0["key"][q, :c].tolist()) == sorted(kv1["key"][q, :c].tolist()) or \
            len(set(kv0["key"][q, :c].tolist()) ^ set(kv1["key"][q, :c].tolist())) <= 2, f"query {q}: result sets differ"
        d0 = np.sort(kv0["vector_distance"][q, :c]); d1 = np.sort(kv1["vector_distance"][q, :c])
        assert np.allclose(d0, d1, rtol=RTOL, atol=ATOL)
        same_order += int(kv0["key"][q, :c].tolist() == kv1["key"][q, :c].tolist())
    assert same_order >= nq - 5, f"only {same_order} of {nq} queries kept the exact order"
    gi.close()
