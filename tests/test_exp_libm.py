"""The exp of the mapping-quality posteriors (processBAM.cpp:4074) is the one transcendental the alignment path evaluates on the device. Its last bit shows in the
output (a column all chains agree on gets phred 255 when the posteriors sum to exactly 1 and 190 when the sum is one ulp short), so the device computes it the way the
host's libm does (hla-la_b200/csrc/exp_libm.cuh). Here: the host restatement of that operation sequence against this host's libm (no GPU), and the device function
itself against libm through hlala_exp_probe (GPU)."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

import harness as H

SRC = os.path.join(H.REPO, "tests", "native", "exp_libm_check.c")
HDR = os.path.join(H.PKG, "csrc", "exp_libm.cuh")


def test_operation_sequence_equals_the_hosts_libm(tmp_path):
    exe = str(tmp_path / "exp_libm_check")
    subprocess.run(["/usr/bin/gcc", "-O2", "-mfma", "-o", exe, SRC, "-lm"], check=True)
    r = subprocess.run([exe, HDR, "40000000"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "0 mismatches of 40000000", r.stdout + r.stderr


@pytest.mark.gpu
def test_device_exp_equals_the_hosts_libm():
    rng = np.random.RandomState(7)
    x = np.concatenate([-rng.rand(150000) * 60, -rng.rand(50000) * 2, -rng.rand(30000) * 700, -rng.rand(20000) * 1e-3, -np.ldexp(rng.rand(2000), -rng.randint(0, 70, 2000)),
                        np.array([0.0, -0.0, -1.0, -745.2, -800.0, -1e-300, -0.6931471805599453])])
    y = np.zeros_like(x)
    L = C.CDLL(H.LIB_PRODUCT); L.hlala_last_error.restype = C.c_char_p
    assert L.hlala_exp_probe(C.c_int(0), C.c_int64(len(x)), H.p(x), H.p(y)) == 0, L.hlala_last_error()
    want = np.array([math.exp(v) for v in x])          # math.exp is libm's exp (numpy has its own vector loops)
    main = (np.abs(x) < 512)                            # where the sums of the posteriors can feel the last bit; beyond, results are below 1e-222
    bad = np.nonzero(y[main].view(np.uint64) != want[main].view(np.uint64))[0]
    assert len(bad) == 0, "%d of %d differ, first at x = %r: device %r, libm %r" % (len(bad), main.sum(), x[main][bad[0]], y[main][bad[0]], want[main][bad[0]])
    assert np.allclose(y[~main], want[~main], rtol=1e-12, atol=0)


def _fixture(tmp_path):
    import sys
    sys.path.insert(0, os.path.join(H.REPO, "tests", "golden"))
    from make_exp_lastbit import PRG_KW
    gold = np.load(os.path.join(H.REPO, "tests", "golden", "exp_lastbit_pairs.npz"))
    d = str(tmp_path / "prg"); H.synth_prg(d, **PRG_KW)
    return d, gold, {k[3:]: gold[k] for k in gold.files if k.startswith("in_")}


def _same_as_gold(got, gold):
    pack = lambda cols, n: np.concatenate([cols[r, :n[r]] for r in range(len(n))])
    assert np.array_equal(got["n_cols"], gold["n_cols"])
    for k in ("level", "edge", "gchar", "schar", "from_seed", "mapq"):
        assert np.array_equal(pack(got[k], got["n_cols"]), gold[k]), k
    assert np.allclose(got["pair_mapq"], gold["pair_mapq"], rtol=0, atol=1e-12) and np.allclose(got["read_mapq"], gold["read_mapq"], rtol=0, atol=1e-12)


def test_last_bit_fixture_on_the_oracle(tmp_path):
    """the five pairs of tests/golden/exp_lastbit_pairs.npz (expected values from the compiled reference): the restatement agrees, and the read in question has the pattern
    that makes the last bit visible: 110 columns of phred 255 (posteriors summing to exactly 1) next to 3 columns two of its four chains do not share"""
    d, gold, b = _fixture(tmp_path)
    _same_as_gold(H.quiet(H.Oracle(d).pairs, b, 100.0, 10.0, 640), gold)
    n = gold["n_cols"]; q = gold["mapq"][int(n[:4].sum()):int(n[:5].sum())]
    assert int((q == 255).sum()) == 110 and int((q == 38).sum()) == 3
