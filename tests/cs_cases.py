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


def exact_windows(ref_starts, ref_lens, seed=3, n_random=300):
    """(start, sequence_len) pairs for DecodeRefSequenceExact(seq, start, sequence_len, 0): inside
    contigs (odd/even starts and lengths), running over a contig end ('x' padding), starting in the
    1000-N spacer before a contig. Starts behind a contig end but not within 1000 of the next contig
    make the reference decode a negative length (it crashes), so they are not part of the contract."""
    rng = np.random.default_rng(seed)
    out = []
    for s0, L in zip(ref_starts, ref_lens):
        end = s0 + L
        for st in (s0, s0 + 1, s0 + 2, s0 + 7, max(s0, end - 50), max(s0, end - 51), max(s0 + 1, end - 2), end - 1):
            for ln in (2, 3, 10, 11, 64, 301):
                out.append((int(st), int(ln)))
        for back in (1, 2, 17, 500, 998, 999):      # inside the spacer before the contig
            for ln in (5, 18, 700, 1301):
                out.append((int(s0 - back), int(ln)))
        for _ in range(n_random // len(ref_starts)):
            st = int(rng.integers(s0, end))
            out.append((st, int(rng.integers(2, 3000))))
    return out
