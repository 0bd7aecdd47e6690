"""CPU: the C-ABI library loads, exports every symbol include/npb200.h declares, fails loudly without a GPU,
and its pure host helpers behave."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "npb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(npb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(npb):
    lib = npb.load_library()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(npb.EXPORTS) == syms


def test_no_cpu_fallback(npb):
    import torch
    if torch.cuda.is_available():
        return
    try:
        npb.Context(0)
    except npb.NpbError as e:
        assert e.status == -2
    else:
        raise AssertionError("a context was created without a CUDA device")


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "noparama_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f)).read()
                assert "np_oracle" not in text and "oracle." not in text and "from oracle" not in text, f


def test_scan_order_is_a_permutation(npb):
    for N in (1, 2, 3, 200, 1000, 4097, 100000):
        for sweep in (0, 1, 7):
            o = npb.scan_order(20261018, sweep, N)
            assert np.array_equal(np.sort(o), np.arange(N))
    a, b = npb.scan_order(1, 0, 5000), npb.scan_order(1, 1, 5000)
    assert not np.array_equal(a, b)
    assert np.array_equal(a, npb.scan_order(1, 0, 5000))
    # looks like a shuffle: displacement and successive differences are not structured
    assert abs(np.corrcoef(np.arange(5000), a)[0, 1]) < 0.05
    assert abs(np.corrcoef(a[:-1], a[1:])[0, 1]) < 0.05


def test_status_strings(npb):
    lib = npb.load_library()
    assert lib.npb_status_str(0) == b"ok"
    assert b"Kmax" in lib.npb_status_str(-3)
