"""CPU tests (-m "not gpu"): the C-ABI library loads and exports every symbol include/ema_b200.h
declares; compute entry points fail loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "ema_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(emab_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    import ema_b200
    L = ema_b200.lib()
    names = declared_functions()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_wire_struct_sizes():
    import ema_b200
    from ema_b200 import _lib
    assert _lib.CAND_DTYPE.itemsize == 56
    sizes = (C.c_int32 * 4)()
    ema_b200.lib().emab_abi_sizes(sizes)   # the C side's sizeof: the ctypes mirrors must agree with the header as compiled
    assert list(sizes) == [56, C.sizeof(_lib.Stats), C.sizeof(_lib.RunStats), C.sizeof(_lib.IndexBuildStats)]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import ema_b200
    with pytest.raises(ema_b200.EmabError):
        ema_b200.Context()
    with pytest.raises(ema_b200.EmabError):
        ema_b200.Index(os.path.join(ROOT, "tests", "golden", "tiny_rep", "ref.fa"))
    with pytest.raises(ema_b200.EmabError):
        ema_b200.Session(os.path.join(ROOT, "tests", "golden", "tiny_rep", "ref.fa"))


def test_cli_usage_without_gpu():
    import subprocess
    cli = os.path.join(ROOT, "ema_b200", "ema-b200")
    r = subprocess.run([cli], capture_output=True, text=True)
    assert r.returncode == 0 and "EMA version" in r.stderr
    r = subprocess.run([cli, "align", "-r", "x.fa"], capture_output=True, text=True)
    assert r.returncode != 0 and "exactly one" in r.stderr


def test_host_selftest():
    """host logic without a GPU (emab_host_selftest): the SSE2 tokenizer, the bucket parser (stable barcode sort, fields),
    the integer writer, the inline candidate lists and the output block pool against plain restatements, and
    reverse-complement / reverse / nt4 encoding: the AVX2 paths (picked at run time) against the scalar loops that
    define them, on every byte value, three alignments and every length up to 300 (emab_host_selftest)"""
    import ema_b200
    assert ema_b200.lib().emab_host_selftest() == 0
