"""Seeded genomes / sub-reads for the candidate-search tests (shared by CPU and GPU tests)."""
import numpy as np

from ngmlr_b200 import synth


def genome_contigs():
    g1 = synth.random_genome(60000, 1)
    g2 = synth.random_genome(30001, 2)          # odd length: nibble padding
    g1[5000:5010] = ord("N")
    g1[7000:7400] = ord("A")                    # homopolymer: same-bin de-duplication
    g1[0:3] = ord("N")
    g1[20:25] = ord("N")
    g1[59990:59995] = ord("N")
    g2[100:160] = np.tile(np.frombuffer(b"ACGTAC", dtype=np.uint8), 10)   # tandem repeat
    g3 = np.tile(synth.random_genome(300, 5), 50)                          # repeats: frequencies
    g4 = np.frombuffer(b"N" * 30 + b"ACGTACGTAGCTAGCTAGCATCGATCGATCAGCTACGATCAGCTACGACT" + b"N" * 5 +
                       b"ACGATCGATCGAC", dtype=np.uint8)
    return [g1, g2, g3, np.frombuffer(b"ACGTACGT", dtype=np.uint8), g4]


def subreads(n, seed, contigs=None):
    contigs = contigs or genome_contigs()
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        g = contigs[int(rng.integers(0, 3))]
        s = int(rng.integers(0, g.size - 400))
        sub, _m = synth.mutate(g[s:s + 320], rng, err=float(rng.choice([0, 0.1, 0.2])))
        sub = sub[:int(rng.choice([256, 256, 256, 100, 30, 13, 12, 1]))]
        if rng.random() < 0.3:
            sub = synth.revcomp(sub)
        if rng.random() < 0.2 and sub.size:
            sub = sub.copy()
            sub[rng.integers(0, sub.size, 3)] = ord("N")
        out.append(sub.tobytes())
    out += [b"", b"N" * 40, b"ACGT" * 64, b"A" * 256]
    return out
