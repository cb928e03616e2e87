"""Deterministic synthetic workloads shaped like BASELINE.json's configs.

Everything here is host-side numpy: an i.i.d. ACGT genome, PacBio/ONT-like reads with
ins:del:sub errors, the 256-bp anchors ngmlr's stage 0/2 would have produced for them, and the
resulting (ref window, read, CorridorLine[]) alignment problems that `computeAlignment`
(src/AlignmentBuffer.cpp:226-465) passes to `IAlignment::SingleAlign`.
"""
from dataclasses import dataclass

import numpy as np

from . import corridor as _corr

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
_COMP[:] = ord("N")
for a, b in zip(b"ACGTN", b"TGCAN"):
    _COMP[a] = b


def random_genome(n, seed):
    rng = np.random.default_rng(seed)
    return BASES[rng.integers(0, 4, size=n, dtype=np.uint8)]


def revcomp(seq):
    return _COMP[seq[::-1]]


def mutate(ref_window, rng, err=0.15, ratio=(9, 4, 2)):
    """Apply insertion/deletion/substitution errors (ratio ins:del:sub) to a reference window.
    Returns (read uint8[], read_pos_of_ref int64[len+1]) where read_pos_of_ref[i] is the read
    coordinate at which reference base i starts."""
    L = ref_window.size
    tot = float(sum(ratio))
    p_ins, p_del, p_sub = (err * r / tot for r in ratio)
    u = rng.random(L)
    dele = u < p_del
    sub = (u >= p_del) & (u < p_del + p_sub)
    n_ins = (rng.random(L) < p_ins).astype(np.int64)
    # geometric tail: one more inserted base with probability 0.3, twice
    n_ins += ((n_ins > 0) & (rng.random(L) < 0.3)).astype(np.int64)
    n_ins += ((n_ins > 1) & (rng.random(L) < 0.3)).astype(np.int64)
    base = ref_window.copy()
    if sub.any():
        shift = rng.integers(1, 4, size=int(sub.sum()))
        code = np.searchsorted(BASES_SORTED, base[sub])
        base[sub] = BASES_SORTED[(code + shift) % 4]
    emit = n_ins + (~dele).astype(np.int64)
    starts = np.concatenate(([0], np.cumsum(emit)))
    total = int(starts[-1])
    read = BASES[rng.integers(0, 4, size=total, dtype=np.uint8)]  # inserted bases by default
    keep_idx = np.nonzero(~dele)[0]
    read[starts[keep_idx] + n_ins[keep_idx]] = base[keep_idx]
    return read, starts


BASES_SORTED = np.sort(BASES)


@dataclass
class AlignProblem:
    ref: bytes          # NUL-free reference window (refSeq of SingleAlign)
    qry: bytes          # read part (qrySeq)
    offsets: np.ndarray  # int32[H]  CorridorLine.offset
    lengths: np.ndarray  # int32[H]  CorridorLine.length
    ext_qstart: int = 0
    ext_qend: int = 0

    @property
    def cells(self):
        lo = np.maximum(self.offsets, 0)
        hi = np.minimum(self.offsets + self.lengths, len(self.ref))
        return int(np.maximum(hi - lo, 0).sum())


def read_lengths(n, rng, median=8000, sigma=0.4, lo=1000, hi=40000):
    return np.clip(rng.lognormal(np.log(median), sigma, size=n), lo, hi).astype(np.int64)


def make_problem(genome, start, ref_len, rng, err=0.15, ratio=(9, 4, 2), reverse=False,
                 multiplier=1, anchor_noise=8, slack=0):
    """One interval alignment problem the way computeAlignment builds it: reference window =
    the interval on the reference (+/- slack), corridor from the 256-bp anchors."""
    window = genome[start:start + ref_len]
    if reverse:
        window = revcomp(window)
    read, starts = mutate(window, rng, err, ratio)
    if read.size < 32:
        read = np.concatenate([read, window[:32]])
        starts = np.concatenate([starts, [starts[-1]]])
    qlen = int(read.size)
    # anchors: one per 256-bp sub-read, at the reference position its first base came from
    ay = np.arange(0, max(qlen - 256, 1), 256, dtype=np.int64)
    ax = np.searchsorted(starts, ay, side="right") - 1
    ax = ax + rng.integers(-anchor_noise, anchor_noise + 1, size=ax.size)
    if slack:
        pad = genome[start + ref_len:start + ref_len + slack]
        window = np.concatenate([window, pad if not reverse else revcomp(pad)])
    offs, lens = _corr.corridor_endpoints_with_anchors(qlen, int(window.size), ax, ay, multiplier)
    return AlignProblem(window.tobytes(), read.tobytes(), offs, lens)


def pacbio_problems(n, genome_len=2_000_000, seed=2, median=8000, err=0.15, ratio=(9, 4, 2),
                    genome=None):
    """Config-2-shaped batch: reads log-normal (median 8 kb, sigma 0.4, clipped 1-40 kb), uniform
    start, strand 50/50, 15 % errors split ins 9 : del 4 : sub 2."""
    rng = np.random.default_rng(seed)
    if genome is None:
        genome = random_genome(genome_len, seed + 1000)
    lens = read_lengths(n, rng, median=median)
    out = []
    for L in lens:
        L = int(min(L, genome.size - 1))
        start = int(rng.integers(0, genome.size - L))
        out.append(make_problem(genome, start, L, rng, err, ratio, reverse=bool(rng.integers(0, 2))))
    return out
