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


# ---------------------------------------------------------------------------------------------
# Interval-form workloads: reads as the sequencer reports them + the computeAlignment calls ngmlr
# would issue for them (for the resident pipeline: reads_upload -> cs_run -> compute_alignments).
# ---------------------------------------------------------------------------------------------
@dataclass
class SimInterval:
    """One AlignmentBuffer::computeAlignment call on a simulated read (coordinates of `genome`)."""
    read: int               # index into the read list
    on_read_start: int      # extractReadSeq: read->Seq + onReadStart ...
    read_len: int           # ... readSeqLen bases, reverse-complemented when `reverse`
    reverse: bool
    ref_start: int          # window genome[ref_start : ref_start + ref_len]
    ref_len: int
    ax: np.ndarray          # anchors relative to the window / the aligned read part (src/AlignmentBuffer.cpp:149-158)
    ay: np.ndarray
    full_alignment: bool = False
    corridor: int = 0       # 0: estimateCorridor(interval)

    def problem(self, genome, reads):
        """The SingleAlign problem of the FIRST attempt, as explicit text + CorridorLine rows (CPU arm)."""
        window = genome[self.ref_start:self.ref_start + self.ref_len]
        part = np.frombuffer(reads[self.read], dtype=np.uint8)[self.on_read_start:self.on_read_start + self.read_len]
        if self.reverse:
            part = revcomp_upper(part)
        if self.full_alignment:
            offs, lens = _corr.corridor_full(self.read_len, self.ref_len + 1)
        elif self.ax.size:
            offs, lens = _corr.corridor_endpoints_with_anchors(self.read_len, self.ref_len, self.ax, self.ay, 1)
        else:
            c = self.corridor or _corr.estimate_corridor(self.read_len, self.ref_len)
            offs, lens = _corr.corridor_endpoints(self.read_len, self.ref_len, min(c, 2 * (self.ref_len + 1)))
        # externalQStart / externalQEnd of alignInterval (src/AlignmentBuffer.cpp:1488-1499)
        full = len(reads[self.read])
        lo_aligned = (full - self.on_read_start - self.read_len) if self.reverse else self.on_read_start
        return AlignProblem(window.tobytes(), part.tobytes(), offs, lens, lo_aligned, full - lo_aligned - self.read_len)


def revcomp_upper(seq):
    """computeReverseSeq / cplBase (src/AlignmentBuffer.cpp:1117-1141): upper-case A C G T only."""
    return _COMP_UPPER[seq[::-1]]


_COMP_UPPER = np.arange(256, dtype=np.uint8)
for a, b in zip(b"ACGT", b"TGCA"):
    _COMP_UPPER[a] = b


def _anchors(starts, qlen, rng, noise=8):
    ay = np.arange(0, max(qlen - 256, 1), 256, dtype=np.int64)
    ax = np.searchsorted(starts, ay, side="right") - 1
    return ax + rng.integers(-noise, noise + 1, size=ax.size), ay


def simulate_reads(n, genome, contig_len, seed, median=8000, err=0.15, ratio=(9, 4, 2), sv=False, lo=1000, hi=40000):
    """-> (reads [bytes as sequenced], intervals [SimInterval]). Plain reads (configs[1] / ONT shape): one
    interval per read covering all of it. sv=True (configs[4] shape): every read carries one event drawn
    from {insertion, deletion, inversion} with a length log-uniform in 1-50 kb, and is aligned the way
    ngmlr splits such reads: indels up to 3 kb inside ONE interval (anchors on both sides of the event widen
    the corridor to about the event length), longer ones as two intervals, inversions as three intervals (the
    middle one on the other strand) plus -- for inversions up to 2 kb -- the full-matrix alignment of the
    inverted segment that the small-inversion realignment computes (getCorridorFull)."""
    rng = np.random.default_rng(seed)
    lens = read_lengths(n, rng, median=median, lo=lo, hi=hi)
    n_contigs = genome.size // contig_len
    reads, ivs = [], []
    for ridx, L in enumerate(lens):
        L = int(min(L, contig_len - 120_000 if sv else contig_len - 10))
        c = int(rng.integers(0, n_contigs))
        rev = bool(rng.integers(0, 2))
        if not sv:
            start = c * contig_len + int(rng.integers(0, contig_len - L))
            window = genome[start:start + L]
            read, starts = mutate(window, rng, err, ratio)
            if read.size < 32:
                read = np.concatenate([read, window[:32]])
                starts = np.concatenate([starts, [starts[-1]]])
            ax, ay = _anchors(starts, int(read.size), rng)
            reads.append((revcomp_upper(read) if rev else read).tobytes())
            ivs.append(SimInterval(ridx, 0, int(read.size), rev, start, L, ax, ay))
            continue
        # ---- one structural event in the middle third of the read ----
        kind = int(rng.integers(0, 3))                       # 0 insertion, 1 deletion, 2 inversion
        ev = int(np.exp(rng.uniform(np.log(1000), np.log(50000))))
        left = int(rng.integers(L // 3, 2 * L // 3))
        span = L + (ev if kind == 1 else 0)                  # reference bases under the read
        start = c * contig_len + int(rng.integers(0, contig_len - span - 10))
        a_src = genome[start:start + left]
        if kind == 0:       # insertion in the read: novel sequence between the two flanks
            b_ref0 = start + left
            mid_src = random_genome(ev, int(rng.integers(1 << 30)))
            b_src = genome[b_ref0:b_ref0 + (L - left)]
        elif kind == 1:     # deletion in the read: ev reference bases are skipped
            b_ref0 = start + left + ev
            mid_src = np.zeros(0, np.uint8)
            b_src = genome[b_ref0:b_ref0 + (L - left)]
        else:               # inversion: ev bases of the middle come from the other strand
            ev = min(ev, max(300, L - left - 300))
            b_ref0 = start + left + ev
            mid_src = revcomp_upper(genome[start + left:start + left + ev])
            b_src = genome[b_ref0:b_ref0 + max(L - left - ev, 0)]
        a_read, a_st = mutate(a_src, rng, err, ratio)
        m_read, m_st = mutate(mid_src, rng, err, ratio) if mid_src.size else (np.zeros(0, np.uint8), np.zeros(1, np.int64))
        b_read, b_st = mutate(b_src, rng, err, ratio) if b_src.size else (np.zeros(0, np.uint8), np.zeros(1, np.int64))
        read = np.concatenate([a_read, m_read, b_read])
        full = int(read.size)
        seq = revcomp_upper(read) if rev else read
        reads.append(seq.tobytes())

        def part(lo_r, n_r):   # interval on the aligned orientation [lo_r, lo_r + n_r) -> stored-read coordinates
            return (full - lo_r - n_r, n_r, True) if rev else (lo_r, n_r, False)

        if kind in (0, 1) and ev <= 3000 and a_read.size > 300 and b_read.size > 300:
            # one interval across the event: window = both flanks (+ the deleted bases), anchors from both
            ref_len = left + (ev if kind == 1 else 0) + int(b_src.size)
            ay_a = np.arange(0, max(a_read.size - 256, 1), 256, dtype=np.int64)
            ax_a = np.searchsorted(a_st, ay_a, side="right") - 1
            ay_b = np.arange(0, max(b_read.size - 256, 1), 256, dtype=np.int64)
            ax_b = np.searchsorted(b_st, ay_b, side="right") - 1 + (b_ref0 - start)
            ax = np.concatenate([ax_a, ax_b]) + rng.integers(-8, 9, size=ax_a.size + ax_b.size)
            ay = np.concatenate([ay_a, ay_b + a_read.size + m_read.size])
            p0, pn, prev = part(0, full)
            ivs.append(SimInterval(ridx, p0, pn, prev, start, ref_len, ax, ay))
            continue
        if a_read.size >= 64:
            ax, ay = _anchors(a_st, int(a_read.size), rng)
            p0, pn, prev = part(0, int(a_read.size))
            ivs.append(SimInterval(ridx, p0, pn, prev, start, left, ax, ay))
        if b_read.size >= 64:
            ax, ay = _anchors(b_st, int(b_read.size), rng)
            p0, pn, prev = part(int(a_read.size + m_read.size), int(b_read.size))
            ivs.append(SimInterval(ridx, p0, pn, prev, b_ref0, int(b_src.size), ax, ay))
        if kind == 2 and m_read.size >= 64:
            # the inverted segment aligns on the other strand: the stored-read part is complemented iff the read is not
            lo_r, n_r = int(a_read.size), int(m_read.size)
            p0 = (full - lo_r - n_r) if rev else lo_r
            # aligned sequence = revcomp(m_read) against the forward window
            rq = int(m_read.size)
            st2 = (rq - m_st)[::-1]                   # read_pos_of_ref of the forward window in revcomp(m_read)
            ay2 = np.arange(0, max(rq - 256, 1), 256, dtype=np.int64)
            ax2 = np.searchsorted(st2, ay2, side="right") - 1
            ax2 = np.clip(ax2, 0, max(ev - 1, 0)) + rng.integers(-8, 9, size=ax2.size)
            full_mat = ev <= 2000
            ivs.append(SimInterval(ridx, p0, n_r, not rev, start + left, ev, ax2 if not full_mat else np.zeros(0, np.int64),
                                   ay2 if not full_mat else np.zeros(0, np.int64), full_alignment=full_mat))
    return reads, ivs


def interval_tasks(ivs, reads, genome_to_concat, by_index=True):
    """SimIntervals -> ngmlr_b200.intervals.IntervalTask (resident read parts when by_index, else text).
    genome_to_concat(pos) maps a position of the flat genome to the concatenated, spacer-padded coordinate
    system of the encoded reference."""
    from .intervals import IntervalTask
    out = []
    for iv in ivs:
        on_start = genome_to_concat(iv.ref_start)
        full = len(reads[iv.read])
        # QStart / QEnd of alignInterval (src/AlignmentBuffer.cpp:1488-1499), in the aligned orientation
        lo_aligned = (full - iv.on_read_start - iv.read_len) if iv.reverse else iv.on_read_start
        ext_qs, ext_qe = lo_aligned, full - lo_aligned - iv.read_len
        anchors = []
        for x, y in zip(iv.ax, iv.ay):
            y_full = int(y) + ext_qs
            on_read = (full - y_full - 256) if iv.reverse else y_full
            anchors.append((on_read, on_start + int(x), int(iv.reverse)))
        corridor = iv.corridor or _corr.estimate_corridor(iv.read_len, iv.ref_len)
        if by_index:
            t = IntervalTask(on_ref_start=on_start, on_ref_stop=on_start + iv.ref_len, read_seq=None, corridor=corridor,
                             ext_qstart=ext_qs, ext_qend=ext_qe, full_read_length=full, anchors=anchors,
                             full_alignment=iv.full_alignment, read_index=iv.read, on_read_start=iv.on_read_start,
                             read_seq_len=iv.read_len, reverse=iv.reverse)
        else:
            part = np.frombuffer(reads[iv.read], dtype=np.uint8)[iv.on_read_start:iv.on_read_start + iv.read_len]
            if iv.reverse:
                part = revcomp_upper(part)
            t = IntervalTask(on_ref_start=on_start, on_ref_stop=on_start + iv.ref_len, read_seq=part.tobytes(),
                             corridor=corridor, ext_qstart=ext_qs, ext_qend=ext_qe, full_read_length=full,
                             anchors=anchors, full_alignment=iv.full_alignment)
        out.append(t)
    return out
