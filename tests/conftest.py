import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def built():
    """Build everything once per session if a product of the build is missing (the driver normally calls build() first)."""
    import harness as H
    need = [H.LIB_PRODUCT, H.LIB_ORACLE, H.SYNTH, os.path.join(H.REPO, "tests", "native", "build", "libdp_host.so"),
            os.path.join(H.REPO, "tests", "native", "build", "libtyping_host.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()
    return True


DATASETS = {
    # name: (prg kwargs, reads kwargs, insert-size mean, sd)
    "small": (dict(levels=6000, haps=4, genes=1, alleles=16, seed=7), dict(pairs=150, len=100, seed=7), 100.0, 10.0),
    "S": (dict(levels=25000, haps=4, genes=2, alleles=64), dict(pairs=1200, len=100, clip_frac=0.15), 100.0, 10.0),
    "typing": (dict(levels=40000, haps=4, genes=17, alleles=24, seed=11), dict(pairs=1500, len=100, seed=11, gene_frac=0.8), 100.0, 10.0),
    "typing1k": (dict(levels=60000, haps=4, genes=17, alleles=1000, seed=31), dict(pairs=700, len=150, seed=31, gene_frac=0.85), 100.0, 10.0),   # 1000 alleles per locus like the bench PRG
    "typing250": (dict(levels=35000, haps=4, genes=17, alleles=24, seed=24), dict(pairs=800, len=250, seed=24, gene_frac=0.8, clip_frac=0.2, gap_mean=300, gap_sd=40), 300.0, 40.0),
    "genes": (dict(levels=30000, haps=8, genes=8, alleles=300), dict(pairs=250, len=150, clip_frac=0.15, gene_frac=1.0), 100.0, 10.0),
    # long-read mode (HLA-LA.pl --longReads): single reads of 3-6 kb, 3 % indels, half of them clipped by up to 400 bases, half starting in gene blocks
    "long": (dict(levels=40000, haps=6, genes=3, alleles=32, seed=41), dict(pairs=160, len=4000, single=1, indel_rate=0.03, clip_frac=0.5, clip_max=400, gene_frac=0.5, seed=41), 0.0, 1.0),
    # long-read typing: 17 typed loci, reads concentrated in the gene blocks; the deep one reaches >= 100 observations per allele (strand filter, HLATyper.cpp:1847)
    "long_typing": (dict(levels=40000, haps=4, genes=17, alleles=24, seed=11), dict(pairs=400, len=3000, single=1, indel_rate=0.02, clip_frac=0.3, clip_max=200, gene_frac=0.8, seed=13), 0.0, 1.0),
    "long_typing_deep": (dict(levels=30000, haps=4, genes=17, alleles=16, seed=17), dict(pairs=2600, len=2500, single=1, indel_rate=0.02, clip_frac=0.3, clip_max=200, gene_frac=0.95, seed=19), 0.0, 1.0),
    "long8k": (dict(levels=60000, haps=6, genes=2, alleles=200, seed=45), dict(pairs=70, len=8000, single=1, indel_rate=0.03, clip_frac=0.5, clip_max=400, gene_frac=0.5, seed=45), 0.0, 1.0),   # Viterbi scores beyond 4095: found a wrap-around of the packed keys
    "long_small": (dict(levels=12000, haps=4, genes=1, alleles=16, seed=43), dict(pairs=40, len=1500, single=1, indel_rate=0.03, clip_frac=0.5, clip_max=200, gene_frac=0.5, seed=43), 0.0, 1.0),
    "L250": (dict(levels=20000, haps=6, genes=2, alleles=32, seed=21), dict(pairs=400, len=250, clip_frac=0.3, indel_rate=0.002, seed=21), 250.0, 35.0),
}


@pytest.fixture(scope="session")
def dataset(tmp_path_factory):
    """Deterministic synthetic PRG directories + seed batches, generated once per session."""
    import harness as H
    cache = {}

    def get(name):
        if name not in cache:
            prg_kw, rd_kw, mu, sd = DATASETS[name]
            d = str(tmp_path_factory.mktemp("prg_" + name))
            H.synth_prg(d, **prg_kw)
            b = H.synth_reads(d, os.path.join(d, "seeds.bin"), **rd_kw)
            cache[name] = (d, b, mu, sd)
        return cache[name]
    return get
