"""The product's PRG loader (hla-la_b200/host/prg_graph.cpp, through the C ABI) against the oracle: canonical order, flat layout,
gap-edge paths, jump lists, gap stretches, level anchors; binary cache round trip; error reporting."""
import ctypes as C
import os
import shutil

import numpy as np
import pytest

import harness as H


def test_flat_graph_matches_oracle(dataset):
    d, b, mu, sd = dataset("S")
    og = H.Oracle(d).graph()
    P = H.Product(d)
    node_ord = P.array("node_ord"); node_level = P.array("node_level")
    assert np.array_equal(og["node_level"][node_ord], node_level)
    assert (np.diff(node_level) >= 0).all(), "flat nodes are sorted by level"
    eo = P.array("edge_ord"); ef = P.array("edge_from"); et = P.array("edge_to"); em = P.array("edge_emis")
    assert np.array_equal(og["edge_from"][eo], node_ord[ef]) and np.array_equal(og["edge_to"][eo], node_ord[et]) and np.array_equal(og["edge_emis"][eo], em)
    # inside a level, flat order == canonical order (what "first taken edge in set<Edge*> order" resolves against)
    leo = P.array("level_edge_off")
    for l in range(0, len(leo) - 1, 97):
        seg = eo[leo[l]:leo[l + 1]]
        assert (np.diff(seg) > 0).all()
    assert np.array_equal(P.array("path_off"), og["path_off"]) and np.array_equal(eo[P.array("path_edges")], og["path_edges"])
    assert np.array_equal(P.array("gap_stretch"), og["gap_stretch"])
    P.close()


def test_jump_lists_sorted_by_target_ordinal(dataset):
    d, b, mu, sd = dataset("S")
    P = H.Product(d); node_ord = P.array("node_ord")
    for name, tgt in (("jump_fwd", "path_to"), ("jump_bwd", "path_from")):
        off = P.array(name + "_off"); lst = P.array(name + "_path"); t = node_ord[P.array(tgt)]
        for n in np.nonzero(np.diff(off) > 1)[0]:
            assert (np.diff(t[lst[off[n]:off[n + 1]]]) > 0).all()
    P.close()


def test_level_anchor_quirk(dataset):
    """translation files end with a newline, so the reference's parser appends a 0: level 0 carries position = contig length
    for every contig (processBAM.cpp:4409-4456)."""
    d, b, mu, sd = dataset("small")
    P = H.Product(d)
    ao = P.array("anchor_off"); ap = P.array("anchor_pos"); aid = P.array("anchor_prg_id"); co = P.array("contig_off"); cid = P.array("contig_prg_id")
    lens = dict(zip(cid.tolist(), np.diff(co).tolist()))
    for k in range(ao[0], ao[1]):
        assert ap[k] == lens[int(aid[k])]
    P.close()


def test_cache_round_trip(dataset, tmp_path):
    d, b, mu, sd = dataset("small")
    d2 = str(tmp_path / "prg"); shutil.copytree(d, d2)
    cache = os.path.join(d2, "PRG", "graph.hlala_b200.cache")
    if os.path.exists(cache):
        os.remove(cache)
    A = H.Product(d2); assert os.path.exists(cache)
    Bp = H.Product(d2)
    for k in H.Product.INT32 + H.Product.UINT8 + ["contig_off"]:
        assert np.array_equal(A.array(k), Bp.array(k)), k
    A.close(); Bp.close()
    # an edited translation file (same mtime, other size) invalidates the cache: contig_level and the anchors are baked into it.
    # The edit repeats a level, which the loader refuses — so a load that succeeds would have come from the stale cache.
    tr = os.path.join(d2, "translation", sorted(os.listdir(os.path.join(d2, "translation")))[0])
    st = os.stat(tr); lines = open(tr).read().split("\n"); lines[1] = lines[0]
    open(tr, "w").write("\n".join(lines) + "\n\n"); os.utime(tr, (st.st_atime, st.st_mtime))
    with pytest.raises(RuntimeError, match="translation"):
        H.Product(d2)
    assert not [f for f in os.listdir(os.path.join(d2, "PRG")) if ".tmp" in f]


def test_missing_directory_reports_error(tmp_path):
    lib = C.CDLL(H.LIB_PRODUCT); lib.hlala_last_error.restype = C.c_char_p
    g = C.c_void_p()
    rc = lib.hlala_graph_load(str(tmp_path / "nope").encode(), C.byref(g))
    assert rc == -2 and b"graph" in lib.hlala_last_error()


def test_prg5_is_injected_not_looked_up(dataset, tmp_path):
    """sequences.txt lists contig 5 but the PRG-only FASTA does not hold PRG_5: the reference asserts exactly that and injects "N" (processBAM.cpp:86-88)"""
    d, b, mu, sd = dataset("small")
    d2 = str(tmp_path / "prg"); shutil.copytree(d, d2)
    for f in ("PRG/graph.hlala_b200.cache",):
        if os.path.exists(os.path.join(d2, f)):
            os.remove(os.path.join(d2, f))
    with open(os.path.join(d2, "sequences.txt"), "a") as f:
        f.write("5\tpgf\tPRG_5\t\t1\t1\n")
    with open(os.path.join(d2, "translation", "5.txt"), "w") as f:
        f.write("7\n")
    os.environ["HLALA_NO_GRAPH_CACHE"] = "1"
    try:
        A = H.Product(d); Bp = H.Product(d2)
        off = Bp.array("contig_off"); seq = Bp.array("contig_seq")
        assert len(off) == len(A.array("contig_off")) + 1
        assert bytes(seq[off[-2]:off[-1]]) == b"N" and int(Bp.array("contig_prg_id")[-1]) == 5
        assert int(Bp.array("contig_level")[off[-2]]) == 7
        A.close(); Bp.close()
        # a FASTA that does hold PRG_5 is refused like the reference's assert
        with open(os.path.join(d2, "mapping_PRGonly", "referenceGenome.fa"), "a") as f:
            f.write(">PRG_5\nACGT\n")
        lib = C.CDLL(H.LIB_PRODUCT); lib.hlala_last_error.restype = C.c_char_p
        g = C.c_void_p()
        assert lib.hlala_graph_load(d2.encode(), C.byref(g)) != 0 and b"PRG_5" in lib.hlala_last_error()
    finally:
        os.environ.pop("HLALA_NO_GRAPH_CACHE", None)
