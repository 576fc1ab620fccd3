"""Parity at scale on the GPU, driver-run: tools/scale_parity.py (own process) on a PRG built with the bench generator's parameters
(8 haplotypes, 17 gene blocks x 1000 alleles): the production cascade against the scalar DP path with duplicate chains aligned over the
whole batch, and against the compiled reference on every differing pair plus 3000 sampled pairs; and the random-graph fuzz with its GPU leg."""
import json
import os
import subprocess
import sys

import pytest

import harness as H

pytestmark = pytest.mark.gpu


def test_cascade_scalar_and_reference_agree_at_scale(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("scale"))
    r = subprocess.run([sys.executable, os.path.join(H.REPO, "tools", "scale_parity.py"), "--levels", "600000", "--genes", "17", "--alleles", "1000", "--haps", "8", "--pairs", "200000",
                        "--check-pairs", "3000", "--dir", d], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=1500)
    assert r.stdout.strip(), r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert r.returncode == 0, line
    assert line["pairs_differing_cascade_vs_scalar"] == 0 and line["pairs_differing_from_checker"] == [] and line["pair_mapq_differing_from_checker"] == [] and line["full_batch_vs_subbatch_mismatches"] == 0
    assert line["pairs_checked"] >= 3000
    if H.have_ref():
        assert line["checked_against"].startswith("compiled reference")


@pytest.mark.parametrize("seed", [3, 12, 25])
def test_random_graph_fuzz_gpu_leg(seed):
    if not H.have_ref():
        pytest.skip("needs oracle/_ref")
    r = subprocess.run([sys.executable, os.path.join(H.REPO, "tests", "align_fuzz.py"), str(seed), "--gpu"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "GPU kernels ok" in r.stdout, r.stdout[-2000:]
