"""SURVEY §8 f-1 on the GPU: tsgpu_art_walk_batch (one thread per search, art_kernels.cu) against the host walk of the same
ART mirror (art_mirror_t::walk_hits, which tests/test_art_mirror.py pins on the reference's compiled art.cpp). It passed on the driver's B200 in round 1 and gates since round 2."""
import ctypes as C

import numpy as np
import pytest

import test_art_mirror as T
from test_art_mirror import am  # noqa: F401  (fixture)

DTYPES = [np.uint32, np.uint16, np.uint8, np.uint8, np.uint8, np.int32, np.uint64, np.uint8]


def flat_arrays(am, h):
    am.am_flat.restype = C.c_size_t
    am.am_flat.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    am.am_root.restype = C.c_int32
    am.am_root.argtypes = [C.c_void_p]
    out = []
    for which, dt in enumerate(DTYPES):
        n = am.am_flat(h, which, None)
        a = np.zeros(max(n, 1), dt)
        am.am_flat(h, which, a.ctypes.data)
        out.append(a[:n])
    return am.am_root(h), out


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["dfs", "frontier"])
def test_art_walk_batch_matches_host_walk(mode):
    """Runs the check below in a child process: a kernel that has never run may fault, and a poisoned CUDA context must not
    reach the fixtures of the tests that gate the suite. dfs: one thread per search (art_walk_kernel); frontier: one thread per
    node visit, level by level, hits sorted back by pre-order rank (art_frontier_kernel; small chunks so several are needed)."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, TSGPU_ART_CHILD="1")
    if mode == "frontier":        # the default since round 2; tiny chunks and item buffers so that several chunks and the split-on-overflow path run
        env.update(TSGPU_ART_MODE="frontier", TSGPU_ART_CHUNK="64", TSGPU_ART_CHUNK2="16", TSGPU_ART_ITEMS="4096")
    else:
        env.update(TSGPU_ART_MODE="dfs")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-m", "gpu", "--runxfail", "-p", "no:cacheprovider",
                        "-k", "child_art_walk"], env=env, capture_output=True, text=True, timeout=300,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and " passed" in r.stdout, (r.stdout + r.stderr)[-3000:]


@pytest.mark.gpu
def test_child_art_walk_batch_matches_host_walk(am):
    import os
    if os.environ.get("TSGPU_ART_CHILD") != "1":
        pytest.skip("runs inside test_art_walk_batch_matches_host_walk's child process")
    from typesense_b200 import capi
    am.am_walk.restype = C.c_size_t
    am.am_walk.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_size_t, C.POINTER(C.c_int)]
    rng = np.random.default_rng(4711)
    cap = 512
    buf = np.zeros(1 << 15, np.int32)
    so = C.c_int(0)
    total = 0
    for trial in (0, 1, 3, 5, 9, 11):
        coll = T.make_collection(rng, trial)
        toks = sorted(coll.vocab, key=coll.vocab.get)
        df = np.diff(coll.flat.list_off.astype(np.int64)).astype(np.uint32)
        ms = np.zeros(len(toks), np.int64)
        if trial % 2 and T.ol.have_ref() and hasattr(T.ol.ref(), "ref_art_new"):      # a mirror LOADED from the reference's live tree
            R = T.ol.ref()
            rt = T.ref_tree(R, coll)
            blob = T.export(R, rt)
            R.ref_art_free(rt)
            h = am.am_load(blob, len(blob))
            assert h
        else:
            h = am.am_build("\n".join(toks).encode(), ms.ctypes.data_as(C.POINTER(C.c_int64)), T.ol.p32(df), len(toks))
        root, arrs = flat_arrays(am, h)
        gi = capi.GpuIndex(coll.n_docs, 0)
        fid = gi.load_field(coll.flat)
        gi.load_art(fid, root, *arrs)
        qs = list(T.queries(rng, coll, 300))
        hits, flags = gi.art_walk(fid, [q["term"].encode() for q in qs], [q["cost"] for q in qs], [q["cost"] for q in qs],
                                  [q["prefix"] for q in qs], cap)
        for q, got, fl in zip(qs, hits, flags):
            n = am.am_walk(h, 0, q["term"].encode(), q["cost"], q["cost"], q["prefix"], buf.ctypes.data_as(C.POINTER(C.c_int32)), len(buf), C.byref(so))
            want = buf[:n].tolist()
            if fl == 4:
                assert n > cap and got == want[:cap]
            else:
                assert fl == 0 and got == want, (trial, q["term"], q["cost"], q["prefix"], fl)
            total += n
        long_term = b"x" * 40
        hits, flags = gi.art_walk(fid, [long_term], [1], [1], [0], cap)
        assert flags[0] == 2 and hits[0] == []
        gi.close()
        am.am_free(h)
    assert total > 2000


@pytest.mark.gpu
def test_host_layer_scenarios_with_device_walk():
    """tests/cpp/host_scenarios (the reference's typo / prefix / ranking scenarios through the C++ host layer) with every
    candidate walk routed through tsgpu_art_walk_batch, plus the rank-fusion KATs on 3-vector graphs through
    Index::hybrid_search / vector_search (a graph size the GPU hybrid path has not been run on yet)."""
    import os
    import subprocess
    import test_cpp_host as tch
    if os.environ.get("TSGPU_TEST_DOUBLE") == "1":
        pytest.skip("links the real libtsgpu.so")
    tch.build()
    r = subprocess.run([tch.BIN, os.path.join(tch.ROOT, "tests", "golden", "documents.jsonl")], capture_output=True, text=True, cwd=tch.ROOT,
                       env=dict(os.environ, TSGPU_HOST_DEVICE_ART="1", TSGPU_HOST_HYBRID_KAT="1"), timeout=300)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
