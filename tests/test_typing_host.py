"""Host side of the typing stage (no GPU): table loading through the C ABI against the oracle's tables, loud failure without a device."""
import ctypes as C

import numpy as np
import pytest

import harness as H


def test_typer_tables_match_oracle(dataset, tmp_path):
    d, b, mu, sd = dataset("typing")
    P = H.Product(d); T = H.ProductTyping(P, d)
    dims = T.table_dims()
    assert [x[0] for x in dims] == ["A", "B", "C", "DQA1", "DQB1", "DRB1", "DPA1", "DPB1", "DRA", "DRB3", "DRB4", "E", "F", "G", "H", "K", "V"]
    aln = H.Oracle(d).pairs(b, mu, sd, 512)
    O = H.OracleTyping(d, b, aln, mu, sd, str(tmp_path / "hla"))
    for i, (name, Cn, Pn) in enumerate(dims):
        assert O.locus(i)["C"] == Cn, name
        assert Pn in (270, 546), (name, Pn)
    O.close(); T.close(); P.close()


def test_typing_without_gpu_fails_loudly(dataset):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    d, b, mu, sd = dataset("typing")
    P = H.Product(d); T = H.ProductTyping(P, d)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        T.infer([b"\0" * 64], mu, sd, None)
    T.close(); P.close()


@pytest.mark.parametrize("name", ["typing", "typing250"])
def test_host_typing_logic_writes_the_oracles_files(dataset, tmp_path, name):
    """hla_typing.cpp (projection, filters, calls, QC, writers) driven by a test-only loop stand-in for the two kernels
    (tests/native/typing_host.cpp): all 73 files byte-identical to the oracle's (which is pinned to the compiled reference)."""
    import filecmp
    import os
    d, b, mu, sd = dataset(name)
    aln = H.Oracle(d).pairs(b, mu, sd, 512)
    or_dir = str(tmp_path / "oracle" / "hla")
    O = H.OracleTyping(d, b, aln, mu, sd, or_dir); O.close()
    lib = C.CDLL(os.path.join(H.REPO, "tests", "native", "build", "libtyping_host.so")); lib.typing_host_last_error.restype = C.c_char_p
    rmq = np.ascontiguousarray(aln["read_mapq"], np.float64)
    for roundtrip in (0, 1):
        out = str(tmp_path / ("host%d" % roundtrip) / "hla"); os.makedirs(out)
        n = lib.typing_host_run(d.encode(), C.c_longlong(len(b["read_off"]) - 1), H.p(b["read_off"]), H.p(b["bases"]), H.p(b["quals"]), C.c_int(512), H.p(aln["n_cols"]), H.p(aln["level"]),
                                H.p(aln["gchar"]), H.p(aln["schar"]), H.p(aln["mapq"]), H.p(aln["read_reverse"]), H.p(rmq), C.c_double(mu), C.c_double(sd), out.encode(), C.c_int(roundtrip))
        assert n == 17, lib.typing_host_last_error().decode()
        fo = sorted(os.listdir(or_dir)); assert sorted(os.listdir(out)) == fo and len(fo) == 73
        bad = [f for f in fo if not filecmp.cmp(os.path.join(or_dir, f), os.path.join(out, f), shallow=False)]
        assert not bad, bad
