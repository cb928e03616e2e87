"""GPU parity tests (B200) of the device text stage and the resident pipeline:
  * convex_text.cu (CIGAR / MD / NM / identity / positions / nmPerPosition / low-identity regions on the
    device) == the host text stage (cigar_text.cpp, itself pinned to the reference's convertCigar through
    the oracle and the golden vectors), and both == a plain restatement of the peak scan in this file;
  * ngmlr_b200_compute_alignments (C++: windows by position, corridors in closed form generated on the
    device, read parts gathered from the resident read set, retries as further batches) == the Python
    mirror with explicit CorridorLine rows == (CPU test) AlignmentBuffer::computeAlignment;
  * stage 0/2 on the sub-reads of the resident read set == the explicit sub-read upload.
"""
import numpy as np
import pytest

import cases
import cs_cases
from ngmlr_b200 import B200Aligner, PackedBatch, PackedReads, split_read, synth
from oracle_lib import CsOracle, same_alignment

pytestmark = pytest.mark.gpu


@pytest.fixture()
def fresh():
    a = B200Aligner(0)
    yield a
    a.close()


def _junk_problems(n, seed):
    """Config-2-shaped problems whose reads carry stretches of unrelated sequence: the alignment runs
    through them with an error density that detectMisalignment's peak scan reports."""
    rng = np.random.default_rng(seed)
    g = synth.random_genome(400_000, seed + 1)
    out = []
    for k in range(n):
        L = int(rng.integers(1500, 6000))
        p = synth.make_problem(g, int(rng.integers(0, g.size - L - 10)), L, rng, err=0.12,
                               reverse=bool(rng.integers(0, 2)))
        q = bytearray(p.qry)
        for _ in range(int(rng.integers(1, 4))):
            w = int(rng.integers(30, 140))
            a = int(rng.integers(200, len(q) - 200 - w))
            q[a:a + w] = synth.random_genome(w, int(rng.integers(1 << 30))).tobytes()
        p.qry = bytes(q)
        p.ext_qstart = int(rng.integers(0, 40)) if k % 3 == 0 else 0
        p.ext_qend = int(rng.integers(0, 40)) if k % 4 == 0 else 0
        out.append(p)
    return out


def _scan_regions(nm, alignment_length):
    """AlignmentBuffer::detectMisalignment's peak scan (src/AlignmentBuffer.cpp:1319-1388), restated."""
    regions = []
    start = None
    stop = None
    dist = 20
    for i in range(alignment_length):
        pr, pq, v = (int(x) for x in nm[i]) if i < len(nm) else (0, 0, 0)
        ident = (32 - v) / 32.0
        peak = 0.0 < ident < 0.75
        if start is None:
            if peak:
                start = stop = (pr, pq)
        elif peak:
            stop = (pr, pq)
            dist = 20
        elif dist == 0:
            regions.append((start[0], stop[0], start[1], stop[1]))
            start = stop = None
            dist = 20
        else:
            dist -= 1
    return regions


def _run(al, batch, on_device, want_nm):
    al.set_text_stage(on_device, want_nm)
    al.upload(batch)
    al.run()
    return list(al.fetch()), al.stats()


def _same_text(a, b, with_nm=True):
    bad = same_alignment(a.as_dict(), b.as_dict()) if with_nm else \
        [k for k in ("ret", "score_bits", "position_offset", "qstart", "qend", "nm", "alignment_length",
                     "cigar_op_count", "sv_type", "first_ref", "first_read", "last_ref", "last_read", "nm_count",
                     "cigar", "md", "identity_bits") if a.as_dict()[k] != b.as_dict()[k]]
    if a.as_dict()["identity_bits"] != b.as_dict()["identity_bits"]:
        bad.append("identity_bits")
    if a.threw != b.threw:
        bad.append("threw")
    if a.nSvRegions != b.nSvRegions or not np.array_equal(a.svRegions, b.svRegions):
        bad.append("sv_regions")
    return bad


def test_device_text_stage_equals_host_text_stage(fresh, oracle):
    probs = (cases.random_problems(64, 303, max_len=2500) + cases.edge_problems() + _junk_problems(48, 17)
             + synth.pacbio_problems(24, genome_len=300_000, seed=9, median=3000))
    batch = PackedBatch.from_problems(probs)
    host, _ = _run(fresh, batch, False, False)
    dev, st = _run(fresh, batch, True, True)
    assert st["text_launches"] >= 1 and st["text_bytes"] > 0
    bad = [(i, d) for i, (h, g) in enumerate(zip(host, dev)) if (d := _same_text(h, g))]
    assert not bad, f"{len(bad)} of {len(probs)} differ: {bad[:6]}"
    # the host stage against the oracle (the reference's convertCigar) on a sample, and both region scans
    # against the restatement above
    for i in range(0, len(probs), 7):
        p = probs[i]
        want = oracle.single_align(p.ref, p.qry, p.offsets, p.lengths, p.ext_qstart, p.ext_qend)
        assert same_alignment(want, host[i].as_dict()) == [], i
    n_regions = 0
    for h in host:
        if h.ret < 0:
            continue
        want = _scan_regions(h.nmPerPosition, h.alignmentLength)
        assert h.nSvRegions == len(want)
        assert [tuple(int(x) for x in r) for r in h.svRegions] == want[:32]
        n_regions += len(want)
    assert n_regions >= 20, n_regions          # the scan is exercised, not vacuous
    assert sum(h.ret >= 0 for h in host) >= 100
    # without the nmPerPosition stream: everything else unchanged
    lean, _ = _run(fresh, batch, True, False)
    bad = [(i, d) for i, (h, g) in enumerate(zip(host, lean)) if (d := _same_text(h, g, with_nm=False))]
    assert not bad, bad[:6]
    assert all(len(g.nmPerPosition) == 0 for g in lean)


def test_device_text_long_runs_and_clips(fresh):
    """Long deletions (cooperative copy of the MD bases), long insertions, reads that are clipped at both
    ends, and alignments whose first recorded columns lie inside the 16-base margin."""
    rng = np.random.default_rng(5)
    g = synth.random_genome(200_000, 77)
    probs = []
    for k in range(24):
        L = int(rng.integers(1200, 4000))
        s0 = int(rng.integers(0, g.size - L - 2000))
        src = g[s0:s0 + L]
        cut = int(rng.choice([70, 130, 300, 700]))
        mid = L // 2
        if k % 2 == 0:
            src_read = np.concatenate([src[:mid], src[mid + cut:]])                    # deletion in the read
        else:
            src_read = np.concatenate([src[:mid], synth.random_genome(cut, k + 1), src[mid:]])   # insertion
        read, _ = synth.mutate(src_read, rng, err=0.08)
        junk = synth.random_genome(int(rng.integers(0, 60)), 1000 + k)
        read = np.concatenate([junk, read, junk[::-1]])                                # forces soft clips
        from ngmlr_b200 import corridor
        o, l = corridor.corridor_full(len(read), L)
        probs.append(synth.AlignProblem(src.tobytes(), read.tobytes(), o, l, int(k % 5), int(k % 3)))
    batch = PackedBatch.from_problems(probs)
    host, _ = _run(fresh, batch, False, False)
    dev, _ = _run(fresh, batch, True, True)
    bad = [(i, d) for i, (h, g) in enumerate(zip(host, dev)) if (d := _same_text(h, g))]
    assert not bad, bad[:6]
    assert sum("D" in h.pBuffer1 and max(int(x[:-1]) for x in __import__("re").findall(r"\d+D", h.pBuffer1)) >= 64
               for h in host if h.ret >= 0) >= 6


def test_text_arena_overflow_recovery(fresh):
    probs = cases.random_problems(32, 404, max_len=1500)
    batch = PackedBatch.from_problems(probs)
    want, _ = _run(fresh, batch, True, True)
    fresh.debug_set_arena_words(64)          # direction arena AND text arena start far too small
    got, st = _run(fresh, batch, True, True)
    fresh.debug_set_arena_words(-1)
    assert st["text_launches"] > 1 and st["fill_launches"] > 1
    assert all(_same_text(a, b) == [] for a, b in zip(want, got))


def _setup_reference(al):
    from ngmlr_b200 import refindex
    contigs = cs_cases.genome_contigs()
    enc = refindex.encode_reference(contigs)
    al.set_reference(enc)
    return contigs, enc


def test_compute_alignments_equals_python_mirror(fresh):
    """C++ computeAlignment mirror (closed-form corridors generated on the device) == the Python mirror that
    ships CorridorLine rows built with corridor.py (pinned to the reference's builders)."""
    from ngmlr_b200.intervals import B200Backend, compute_alignments
    from test_cs_oracle import _interval_tasks
    contigs, enc = _setup_reference(fresh)
    tasks = _interval_tasks(contigs, enc)
    fresh.set_text_stage(False, False)
    want, want_calls = compute_alignments(B200Backend(fresh), _interval_tasks(contigs, enc))
    got, calls = fresh.compute_alignments(tasks)
    got = list(got)
    assert list(calls) == want_calls
    n_valid = 0
    for w, g in zip(want, got):
        assert (w is None) == (g.ret < 0)
        if w is not None:
            n_valid += 1
            assert _same_text(w, g, with_nm=False) == []
    assert n_valid >= 24 and max(calls) >= 2


def test_compute_alignments_from_the_resident_read_set(fresh):
    """The same intervals with their read parts named by index into reads that were uploaded once
    (forward parts and reverse-complemented ones, at an offset inside a longer read) == read parts as text;
    intervals without a decodable window or with an empty reference span come back as 'no alignment'."""
    from ngmlr_b200.intervals import IntervalTask
    from test_cs_oracle import _interval_tasks
    contigs, enc = _setup_reference(fresh)
    tasks = _interval_tasks(contigs, enc)
    want, want_calls = fresh.compute_alignments(tasks)
    want = list(want)
    rng = np.random.default_rng(3)
    reads, rtasks = [], []
    for i, t in enumerate(tasks):
        part = np.frombuffer(t.read_seq, dtype=np.uint8)
        rev = i % 2 == 1
        left = synth.random_genome(int(rng.integers(0, 300)), 50 + i)
        right = synth.random_genome(int(rng.integers(0, 300)), 90 + i)
        body = synth.revcomp(part) if rev else part
        reads.append(np.concatenate([left, body, right]).tobytes())
        rtasks.append(IntervalTask(on_ref_start=t.on_ref_start, on_ref_stop=t.on_ref_stop, read_seq=None,
                                   corridor=t.corridor, ext_qstart=t.ext_qstart, ext_qend=t.ext_qend,
                                   full_read_length=t.full_read_length, anchors=t.anchors, realign=t.realign,
                                   full_alignment=t.full_alignment, short_read=t.short_read, read_index=i,
                                   on_read_start=len(left), read_seq_len=len(part), reverse=rev))
    # two intervals the reference answers with 0: empty reference span, window behind the last contig
    bad_a = IntervalTask(on_ref_start=5000, on_ref_stop=5000, read_seq=None, corridor=100, read_index=0,
                         read_seq_len=50, full_read_length=50)
    bad_b = IntervalTask(on_ref_start=enc.ref_start[-1] + enc.ref_len[-1] + 400, on_ref_stop=enc.concat_len + 99,
                         read_seq=None, corridor=100, read_index=0, read_seq_len=50, full_read_length=50)
    fresh.reads_upload(reads)
    got, calls = fresh.compute_alignments(rtasks + [bad_a, bad_b])
    got = list(got)
    assert list(calls[:len(tasks)]) == list(want_calls) and list(calls[-2:]) == [0, 0]
    assert got[-1].ret < 0 and got[-2].ret < 0
    for w, g in zip(want, got):
        assert _same_text(w, g, with_nm=False) == []
    st = fresh.compute_alignments_stats()
    assert st["h2d_bytes"] < 400 * len(rtasks) * max(calls)      # descriptors only: no sequence, no corridor rows


def test_stage02_on_resident_reads_equals_subread_upload(fresh):
    from ngmlr_b200 import refindex
    contigs = cs_cases.genome_contigs()
    enc = refindex.encode_reference(contigs)
    fresh.set_index(refindex.build_index(enc))
    fresh.set_reference(enc)
    rng = np.random.default_rng(21)
    reads = []
    for k in range(40):
        c = int(rng.integers(0, len(contigs)))
        L = min(int(rng.choice([90, 255, 256, 257, 1000, 2600, 5000])), contigs[c].size - 1)
        s0 = int(rng.integers(0, contigs[c].size - L))
        r, _ = synth.mutate(contigs[c][s0:s0 + L], rng, err=0.1)
        if k % 7 == 0:
            r[rng.integers(0, r.size, 5)] = ord("N")
        reads.append((synth.revcomp(r) if k % 2 else r).tobytes())
    subs = [s for r in reads for s in split_read(r)]
    fresh.cs_upload(PackedReads(subs))
    m0, _ = fresh.cs_run()
    want = fresh.cs_fetch()
    n_sub = fresh.reads_upload(reads)
    assert n_sub == len(subs)
    m1, _ = fresh.cs_run()
    got = fresh.cs_fetch()
    assert m0 == m1 and m0 > 50
    for a, b in zip(want, got):
        assert np.array_equal(a, b)
