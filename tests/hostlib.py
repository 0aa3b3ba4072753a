"""Builds the C++ host layer's C wrapper against the oracle-backed test double of the C-ABI (TEST INFRASTRUCTURE: the CPU arm
of bench.py and the CPU runs of the host-layer tests; never part of the product)."""
import os
import subprocess

import oracle_lib as ol

ROOT = ol.ROOT
DOUBLE_SO = os.path.join(ROOT, "tests", "cpp", "libtsgpu_double.so")
HOST_CPU_SO = os.path.join(ROOT, "tests", "cpp", "libtshost_cpu.so")


def _stale(target, deps):
    return not os.path.exists(target) or any(os.path.getmtime(d) > os.path.getmtime(target) for d in deps)


def build_double() -> str:
    ol.build_oracle()
    src = os.path.join(ROOT, "tests", "cpp", "tsgpu_oracle_double.cpp")
    deps = [src, os.path.join(ROOT, "include", "tsgpu.h"), os.path.join(ROOT, "oracle", "liboracle.so")]
    if _stale(DOUBLE_SO, deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wno-unused", "-fPIC", "-shared", src, "-o", DOUBLE_SO + ".tmp", "-L", os.path.join(ROOT, "oracle"),
                               "-l:liboracle.so", f"-Wl,-rpath,{os.path.join(ROOT, 'oracle')}", "-pthread"])
        os.replace(DOUBLE_SO + ".tmp", DOUBLE_SO)
    return DOUBLE_SO


def build_host_cpu() -> str:
    dbl = build_double()
    hd = os.path.join(ROOT, "typesense_b200", "host")
    src = os.path.join(hd, "tshost_capi.cpp")
    deps = [src, os.path.join(hd, "tsgpu_host.hpp"), os.path.join(hd, "art_mirror.hpp"), dbl]
    if _stale(HOST_CPU_SO, deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", src, "-o", HOST_CPU_SO + ".tmp", "-L", os.path.dirname(dbl), "-l:libtsgpu_double.so",
                               "-Wl,-rpath,$ORIGIN", "-pthread"])
        os.replace(HOST_CPU_SO + ".tmp", HOST_CPU_SO)
    return HOST_CPU_SO
