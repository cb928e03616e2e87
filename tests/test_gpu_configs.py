"""GPU parity tests (B200) on the shapes of the other BASELINE.json configs, at sizes the CPU oracle
finishes in seconds: ONT-shaped reads (configs[3]: longer reads, 12 % errors 1:1:1) and the SV stress
shape (configs[4]: reads with 1-50 kb insertions / deletions / inversions -> anchor-widened corridors of
several thousand columns, split intervals on both strands, full-matrix alignments of inverted segments,
retries with wider corridors). The whole hot path per read: stage 0/2 on the resident reads against the
candidate-search oracle, then ngmlr_b200_compute_alignments against the Python mirror of
AlignmentBuffer::computeAlignment over the CPU oracle (tests/test_cs_oracle.py pins that mirror to the
reference)."""
import numpy as np
import pytest

from ngmlr_b200 import B200Aligner, refindex, synth
from ngmlr_b200.intervals import compute_alignments
from oracle_lib import CsOracle, same_alignment
from test_cs_oracle import _OracleBackend

pytestmark = pytest.mark.gpu

CONTIG = 400_000


@pytest.fixture(scope="module")
def world():
    genome = synth.random_genome(2 * CONTIG, 4242)
    contigs = [genome[:CONTIG], genome[CONTIG:]]
    enc = refindex.encode_reference(contigs)
    idx = refindex.build_index(enc)
    orc = CsOracle([c.tobytes() for c in contigs])
    al = B200Aligner(0)
    al.set_index(idx)
    al.set_reference(enc)
    yield genome, enc, orc, al
    al.close()
    orc.close()


def _g2c(enc):
    return lambda pos: enc.ref_start[pos // CONTIG] + pos % CONTIG


def _check(world, oracle, reads, ivs, min_valid, min_retry=0, max_stage02_reads=6):
    genome, enc, orc, al = world
    # ---- stage 0/2 on the resident reads vs the candidate-search oracle ----
    n_sub = al.reads_upload(reads)
    m, _ms = al.cs_run()
    cstart, sc, lo, rv, sw, mx = al.cs_fetch()
    assert len(cstart) == n_sub + 1 and cstart[-1] == m
    s = 0
    checked = 0
    for r, read in enumerate(reads):
        parts = max(1, len(read) // 256)
        if r < max_stage02_reads:
            for k in range(parts):
                sub = read[k * 256:(k + 1) * 256] if len(read) >= 256 else read
                want, want_mx = orc.search(sub)
                a, b = int(cstart[s + k]), int(cstart[s + k + 1])
                got = [(float(sc[j]), int(lo[j]), int(rv[j])) for j in range(a, b)]
                assert got == [(float(x[0]), int(x[1]), int(x[2])) for x in want], (r, k)
                for j, (_s, loc, rev) in zip(range(a, b), want):
                    w = orc.decode((loc - 20) % (1 << 64), 308) or b"N" * 308
                    q = synth.revcomp_upper(np.frombuffer(sub, dtype=np.uint8)).tobytes() if rev else sub
                    assert float(sw[j]) == oracle.ssw_score(w, q), (r, k, j)
                checked += 1
        s += parts
    assert checked >= 20
    # ---- stage 4: C++ computeAlignment mirror on the device vs the Python mirror over the CPU oracle ----
    tasks_idx = synth.interval_tasks(ivs, reads, _g2c(enc), by_index=True)
    tasks_txt = synth.interval_tasks(ivs, reads, _g2c(enc), by_index=False)
    got, calls = al.compute_alignments(tasks_idx)
    got = list(got)
    want, want_calls = compute_alignments(_OracleBackend(orc, oracle), tasks_txt)
    assert list(calls) == want_calls
    n_valid = 0
    for i, (w, g) in enumerate(zip(want, got)):
        assert (w is None) == (g.ret < 0), i
        if w is not None:
            n_valid += 1
            d = same_alignment(w, dict(g.as_dict(), nm_positions=w["nm_positions"]))   # nmPerPosition not requested
            assert d == [], (i, d)
            assert g.nmCount == len(w["nm_positions"])
    assert n_valid >= min_valid and sum(c > 1 for c in calls) >= min_retry, (n_valid, list(calls))
    return got, calls


def test_ont_shaped_reads(world, oracle):
    genome = world[0]
    reads, ivs = synth.simulate_reads(20, genome, CONTIG, 31, median=6000, err=0.12, ratio=(1, 1, 1), hi=16000)
    _check(world, oracle, reads, ivs, min_valid=19)


def test_sv_stress_shape(world, oracle):
    genome = world[0]
    reads, ivs = synth.simulate_reads(40, genome, CONTIG, 32, median=4000, err=0.15, sv=True, lo=2000, hi=9000)
    kinds = dict(full=sum(iv.full_alignment for iv in ivs), rev=sum(iv.reverse for iv in ivs),
                 wide=sum(1 for iv in ivs if iv.ax.size and iv.problem(genome, reads).lengths[0] > 1500))
    assert kinds["full"] >= 2 and kinds["rev"] >= 10 and kinds["wide"] >= 3, kinds
    got, calls = _check(world, oracle, reads, ivs, min_valid=int(0.8 * len(ivs)), min_retry=0)
    assert len(ivs) > len(reads)            # split reads: more intervals than reads
