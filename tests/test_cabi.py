"""The C-ABI library loads and exports every symbol include/hlala_b200.h declares; compute entry points fail loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import harness as H


def declared_functions():
    txt = open(os.path.join(H.REPO, "include", "hlala_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(hlala_[a-z_0-9]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported():
    lib = C.CDLL(H.LIB_PRODUCT)
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n


def test_no_oracle_symbols_or_dependencies_in_product():
    out = os.popen("ldd %s" % H.LIB_PRODUCT).read() + os.popen("nm -D %s" % H.LIB_PRODUCT).read()
    assert "oracle" not in out and "hlala_ref" not in out


def test_compute_without_gpu_fails_loudly(dataset):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    d, b, mu, sd = dataset("small")
    P = H.Product(d)
    with pytest.raises(RuntimeError, match="-3"):
        P.to_gpu(0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.chains(b, 512)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.pairs(b, mu, sd, 512)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.long_reads(b, 512)          # long-read entry: the same rule
    P.close()


def test_bad_arguments_are_rejected(dataset):
    d, b, mu, sd = dataset("small")
    P = H.Product(d)
    sb = H.make_batch_struct(b); sb.n_reads = 3
    co = H.ChainOut(); co.max_columns = 512
    assert P.lib.hlala_align_chains(P.g, C.byref(sb), C.byref(co)) == -1
    po = H.PairOut(); po.max_columns = 20000          # long reads: max_columns beyond 16384 is refused before anything runs
    assert P.lib.hlala_align_long_reads(P.g, C.byref(H.make_batch_struct(b)), C.byref(po), None) == -1
    assert P.lib.hlala_evaluate_types(None, None, None, None, 0, None, 0, None, 0) == -1
    P.close()
