"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/tsgpu.h declares, and refuses to run
without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from typesense_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "tsgpu.h")).read()
    declared = set(re.findall(r"^(?:tsgpu_status|const char\*|int|void)\s+(tsgpu_[a-z0-9_]+)\(", hdr, re.M))
    assert len(declared) >= 17
    L = capi.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/tsgpu.h but not exported"
    assert set(capi.EXPORTS) <= declared


def test_struct_sizes_match_header():
    assert capi.KV_DTYPE.itemsize == 56
    assert C.sizeof(capi.FieldStruct) == 8 + 4 * 8
    assert C.sizeof(capi.VecParamsStruct) == 28          # k, ef, flat_search_cutoff, distance_threshold, alpha, fetch_size, flags


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    L = capi.lib()
    assert L.tsgpu_device_count() == 0
    h = C.c_void_p()
    rc = L.tsgpu_index_create(10, 0, C.byref(h))
    assert rc == 1   # TSGPU_ERR_NO_DEVICE
    assert b"no CPU fallback" in L.tsgpu_last_error()


def test_opt_in_regscore_kernel_adds_no_local_memory_traffic():
    """SASS of the built library (cuobjdump, no GPU): the opt-in kw_search_kernel<true> exists in its own namespace and has no
    more local-memory instructions than the default kernel — its register-resident scoring branch neither spills nor indexes
    a local array (it still carries the default scoring code as the fallback for long queries / array fields)."""
    import shutil
    import sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_funcs
    import re
    fn = sass_funcs.funcs(os.path.join(ROOT, "typesense_b200", "libtsgpu.so"))
    dflt = [v for k, v in fn.items() if "kw_search_kernelILb0" in k]
    regs = [v for k, v in fn.items() if "kw_search_kernelILb1" in k and "tsk_rs" in k]
    assert len(dflt) == 1 and len(regs) == 2                        # <true, false> and the single-field <true, true>
    cnt = lambda ins, pat: sum(1 for i in ins if re.search(pat, i))
    for r in regs:
        assert cnt(r, r"\bLDL") <= cnt(dflt[0], r"\bLDL") and cnt(r, r"\bSTL") <= cnt(dflt[0], r"\bSTL")
        assert cnt(r, r"\bLDG") > cnt(dflt[0], r"\bLDG")          # the extra branch is really there
