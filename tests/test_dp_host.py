"""The extension DP the GPU executes (hla-la_b200/csrc/extend_dp.h, compiled for the host by tests/native/dp_host.cpp) against the
oracle's independent restatement, on every chain of the datasets that needs an extension."""
import ctypes as C
import os

import numpy as np
import pytest

import harness as H

LIB = os.path.join(H.REPO, "tests", "native", "build", "libdp_host.so")


def check_dp_host(d, b, oc):
    """every extension of the chains in `oc` (oracle / reference output of `chains`) recomputed by extend_dp.h on the host; returns how many were compared"""
    lib = C.CDLL(LIB); lib.dp_host_open.restype = C.c_void_p
    h = C.c_void_p(lib.dp_host_open(d.encode())); assert h
    slot_read = np.repeat(np.arange(len(b["chain_off"]) - 1), np.diff(b["chain_off"]))
    tested = 0
    for i in np.nonzero(oc["status"] == 0)[0]:
        r = slot_read[i]; seq = b["bases"][b["read_off"][r]:b["read_off"][r + 1]].copy(); L = len(seq)
        sb, se = oc["seed_begin"][i], oc["seed_end"][i]
        if sb == 0 and se == L - 1:
            continue
        n = oc["n_cols"][i]; idx = np.nonzero(oc["from_seed"][i, :n])[0]; s0, s1 = idx[0], idx[-1]
        parts = []
        for pos, need, start_seq, eord in ((0, sb != 0, sb, oc["edge"][i, s0]), (1, se != L - 1, se + 1, oc["edge"][i, s1])):
            oe = np.zeros(512, np.int32); os_ = np.zeros(512, np.uint8); nc = C.c_int32(); fy = C.c_int32(); app = C.c_int32()
            if need:
                assert lib.dp_host_extend(h, H.p(seq), len(seq), int(start_seq), int(eord), pos, H.p(oe), H.p(os_), C.byref(nc), C.byref(fy), C.byref(app)) == 0
            parts.append((oe[:nc.value].copy(), os_[:nc.value].copy(), fy.value if need else start_seq))
        (le, ls, lfar), (re_, rs, rfar) = parts
        padL = lfar if sb != 0 else 0; padR = (L - rfar) if se != L - 1 else 0
        edge = np.concatenate([np.full(padL, -1, np.int32), le, oc["edge"][i, s0:s1 + 1], re_, np.full(padR, -1, np.int32)])
        sch = np.concatenate([seq[:padL], ls, oc["schar"][i, s0:s1 + 1], rs, seq[L - padR:] if padR else np.zeros(0, np.uint8)])
        assert len(edge) == n and np.array_equal(edge, oc["edge"][i, :n]) and np.array_equal(sch, oc["schar"][i, :n]), "slot %d" % i
        tested += 1
    return tested


def check_lean_host(d, b, oc, lib_path=None, big=False):
    """every extension of the chains in `oc`: the first GPU tier (extend_lean.h, host build) against the scalar DP (extend_dp.h);
    returns (extensions compared, extensions the tier deferred)"""
    lib = C.CDLL(lib_path or LIB); lib.dp_host_open.restype = C.c_void_p
    h = C.c_void_p(lib.dp_host_open(d.encode())); assert h
    slot_read = np.repeat(np.arange(len(b["chain_off"]) - 1), np.diff(b["chain_off"]))
    tested = deferred = 0
    for i in np.nonzero(oc["status"] == 0)[0]:
        r = slot_read[i]; seq = b["bases"][b["read_off"][r]:b["read_off"][r + 1]].copy(); L = len(seq)
        sb, se = oc["seed_begin"][i], oc["seed_end"][i]
        if sb == 0 and se == L - 1:
            continue
        n = oc["n_cols"][i]; idx = np.nonzero(oc["from_seed"][i, :n])[0]; s0, s1 = idx[0], idx[-1]
        for pos, need, start_seq, eord in ((0, sb != 0, sb, oc["edge"][i, s0]), (1, se != L - 1, se + 1, oc["edge"][i, s1])):
            if not need:
                continue
            res = []
            for fn in (lib.dp_host_extend, lib.dp_host_extend_lean_big if big else lib.dp_host_extend_lean):
                oe = np.zeros(512, np.int32); os_ = np.zeros(512, np.uint8); nc = C.c_int32(); fy = C.c_int32(); app = C.c_int32()
                rc = fn(h, H.p(seq), len(seq), int(start_seq), int(eord), pos, H.p(oe), H.p(os_), C.byref(nc), C.byref(fy), C.byref(app))
                res.append((rc, oe[:nc.value].copy(), os_[:nc.value].copy(), fy.value, app.value))
            assert res[0][0] == 0
            if res[1][0] == -100:
                deferred += 1
                continue
            assert res[1][0] == 0, "slot %d side %d: rc %d" % (i, pos, res[1][0])
            assert res[0][4] == res[1][4] and res[0][3] == res[1][3] and np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2]), "slot %d side %d" % (i, pos)
            tested += 1
    return tested, deferred


@pytest.mark.parametrize("name", ["S", "genes"])
def test_lean_tier_matches_scalar_dp(dataset, name):
    d, b, mu, sd = dataset(name)
    oc = H.Oracle(d).chains(b, 1024)
    tested, deferred = check_lean_host(d, b, oc)
    assert tested > 50 and deferred < tested
    tested_big, deferred_big = check_lean_host(d, b, oc, big=True)
    assert tested_big >= tested and deferred_big <= deferred


def test_lean_key_order_matches_string_order():
    lib = C.CDLL(LIB)
    rng = np.random.default_rng(5)
    xs = [0, 1, 9, 10, 11, 19, 99, 100, 101, 109, 110, 999, 1000, 1001, 1099, 4999999, 5000000, 49, 490, 4900, 12, 120, 1200, 121]
    zs = [0, 1, 2, 9, 10, 11, 15, 19, 100, 255]
    for x1 in xs:
        for x2 in xs:
            for z1 in zs:
                for z2 in zs:
                    assert lib.dp_host_key_less_check(x1, z1, x2, z2) == 1, (x1, z1, x2, z2)
    for _ in range(20000):
        a = [int(v) for v in rng.integers(0, 10 ** int(rng.integers(1, 8)), 2)] + [int(v) for v in rng.integers(0, 256, 2)]
        assert lib.dp_host_key_less_check(a[0], a[2], a[1], a[3]) == 1, a


@pytest.mark.parametrize("name", ["S", "genes"])
def test_extension_dp_matches_oracle(dataset, name):
    d, b, mu, sd = dataset(name)
    oc = H.Oracle(d).chains(b, 1024)
    assert check_dp_host(d, b, oc) > 50
