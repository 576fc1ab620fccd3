"""Pins the typing restatement (oracle/hlala_oracle_typing.cpp) against the UNMODIFIED reference HLATyper::HLATypeInference
(oracle/_ref): every file the reference writes into outDir/hla must be byte-identical, for all 17 loci."""
import filecmp
import os

import numpy as np
import pytest

import harness as H

pytestmark = pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref/libhlala_ref.so not built (reference tree absent)")


def test_typing_files_byte_identical(dataset, tmp_path):
    d, b, mu, sd = dataset("typing")
    R = H.quiet(H.Ref, d)
    aln = H.quiet(R.pairs, b, mu, sd, 512)
    ref_dir = str(tmp_path / "ref" / "hla"); or_dir = str(tmp_path / "oracle" / "hla")
    r = H.quiet(R.type, d, b, mu, sd, ref_dir)
    assert r["n_used"] > 500
    T = H.OracleTyping(d, b, aln, mu, sd, or_dir)
    try:
        fr = sorted(os.listdir(ref_dir)); fo = sorted(os.listdir(or_dir))
        assert fr == fo and len(fr) == 5 + 4 * 17
        bad = [f for f in fr if not filecmp.cmp(os.path.join(ref_dir, f), os.path.join(or_dir, f), shallow=False)]
        assert not bad, "files differ from the reference's: %s" % bad
        assert T.n_loci == 17
        dims = [T.locus(i) for i in range(T.n_loci)]
        assert all(x["C"] >= 8 and x["R"] >= 20 for x in dims)
        assert all(np.isfinite(x["pair_ll"]).all() for x in dims)
    finally:
        T.close()
