"""GPU parity tests of the candidate-search kernel through the C ABI against the oracle."""
import numpy as np
import pytest

import cs_cases
from oracle_lib import CsOracle

pytestmark = pytest.mark.gpu


def test_cs_search_matches_oracle(aligner):
    from ngmlr_b200 import refindex
    contigs = cs_cases.genome_contigs()
    orc = CsOracle([c.tobytes() for c in contigs])
    idx = refindex.build_index(refindex.encode_reference(contigs))
    aligner.set_index(idx)
    subs = cs_cases.subreads(400, 23, contigs)
    got, mx = aligner.cs_search(subs)
    for i, s in enumerate(subs):
        want, wmx = orc.search(s)
        assert got[i] == want, (i, len(s), got[i][:3], want[:3])
        assert mx[i] == np.float32(wmx)
    # a second batch on the same context, different sensitivity
    got2, _ = aligner.cs_search(subs[:50], sensitivity=0.5)
    for i in range(50):
        assert got2[i] == orc.search(subs[i], sensitivity=0.5)[0]
    orc.close()


def _cpl(b):
    t = {ord("A"): ord("T"), ord("T"): ord("A"), ord("C"): ord("G"), ord("G"): ord("C")}
    return bytes(t.get(c, c) for c in b)


def test_fused_search_and_candidate_scoring_matches_oracle(aligner, oracle):
    """CS candidates + device-side DecodeRefSequence + StrippedSW score == oracle composition of
    or_cs_search -> or_cs_decode(loc - 20, 308) -> or_ssw_score (what ScoreBuffer::DoRun computes)."""
    from ngmlr_b200 import refindex
    contigs = cs_cases.genome_contigs()
    orc = CsOracle([c.tobytes() for c in contigs])
    ref = refindex.encode_reference(contigs)
    aligner.set_index(refindex.build_index(ref))
    aligner.set_reference(ref)
    subs = [s for s in cs_cases.subreads(300, 29, contigs)]
    got, _ = aligner.cs_score(subs)
    n_scored = 0
    for i, s in enumerate(subs):
        want, _ = orc.search(s)
        assert [g[:3] for g in got[i]] == want
        for (cs_sc, loc, rev, sw) in got[i]:
            pos = (loc - 20) % (1 << 64)
            window = orc.decode(pos, 308)
            if window is None:
                window = b"N" * 308
            q = _cpl(s)[::-1] if rev else s
            assert sw == oracle.ssw_score(window, q), (i, loc, rev)
            n_scored += 1
    assert n_scored > 250
    # windows at the very end of the genome ('x' padding) and odd/even positions
    tail = contigs[4].tobytes()[-13 - 40:]
    got, _ = aligner.cs_score([tail, contigs[0][1001:1257].tobytes(), contigs[0][1002:1258].tobytes()])
    for i, s in enumerate([tail, contigs[0][1001:1257].tobytes(), contigs[0][1002:1258].tobytes()]):
        for (cs_sc, loc, rev, sw) in got[i]:
            window = orc.decode((loc - 20) % (1 << 64), 308) or b"N" * 308
            q = _cpl(s)[::-1] if rev else s
            assert sw == oracle.ssw_score(window, q)
    orc.close()


def test_resident_pipeline_equals_one_shot_call(aligner):
    """cs_upload / cs_run / cs_fetch (device-resident: prefix sums, compaction and scoring without
    host round trips) must return exactly what cs_score_batch returns."""
    from ngmlr_b200 import refindex
    contigs = cs_cases.genome_contigs()
    ref = refindex.encode_reference(contigs)
    aligner.set_index(refindex.build_index(ref))
    aligner.set_reference(ref)
    subs = cs_cases.subreads(500, 31, contigs)
    want, wmx = aligner.cs_score(subs)
    aligner.cs_upload(subs)
    for _ in range(2):   # repeatable on one upload
        m, ms = aligner.cs_run()
        start, sc, lo, rv, sw, mx = aligner.cs_fetch()
        assert m == start[-1] == sum(len(w) for w in want) and ms > 0
        for i, w in enumerate(want):
            got = [(sc[j], int(lo[j]), int(rv[j]), sw[j]) for j in range(start[i], start[i + 1])]
            assert got == w, i
        assert np.array_equal(mx, wmx)


def test_decode_windows_matches_oracle_and_golden(aligner):
    """DecodeRefSequenceExact on the device: contig interiors, contig ends ('x' padding after the
    byte-pair overshoot), starts inside a spacer, odd/even positions and lengths."""
    import golden_util as gu
    from ngmlr_b200 import refindex
    contigs = cs_cases.genome_contigs()
    orc = CsOracle([c.tobytes() for c in contigs])
    enc = refindex.encode_reference(contigs)
    aligner.set_reference(enc)
    wins = cs_cases.exact_windows(enc.ref_start, enc.ref_len)
    got = aligner.decode_windows([w[0] for w in wins], [w[1] for w in wins])
    g = gu.load("decode_exact_golden.json")
    assert len(got) == g["n"]
    for (st, ln), text, sha in zip(wins, got, g["text_sha"]):
        assert text == orc.decode_exact(st, ln), (st, ln)
        assert gu.digest(text) == sha
    # outside the contract: the call refuses instead of guessing
    with pytest.raises(RuntimeError):
        aligner.decode_windows([enc.concat_len + 5], [10])
    with pytest.raises(RuntimeError):
        aligner.decode_windows([enc.ref_start[-1] + enc.ref_len[-1] + 300], [10])   # behind the last contig
    orc.close()


def test_align_from_reference_windows_equals_text_upload(aligner):
    """convex_upload_windows (reference decoded on the device) == convex_upload with the same windows
    decoded by the oracle: identical alignments incl. CIGAR/MD text."""
    from ngmlr_b200 import PackedBatch, corridor, refindex, synth
    contigs = cs_cases.genome_contigs()
    orc = CsOracle([c.tobytes() for c in contigs])
    enc = refindex.encode_reference(contigs)
    aligner.set_reference(enc)
    rng = np.random.default_rng(11)
    starts, stops, refs, qrys, offs, lens = [], [], [], [], [], []
    for k in range(24):
        c = int(rng.integers(0, 2))
        L = int(rng.integers(300, 2500))
        s0 = int(rng.integers(10, contigs[c].size - L - 10)) if k % 6 else contigs[c].size - L   # some end at the contig end
        read, _ = synth.mutate(contigs[c][s0:s0 + L], rng, err=0.12)
        on_start = enc.ref_start[c] + s0 - (37 if s0 > 40 else 0)
        on_stop = enc.ref_start[c] + s0 + L + (53 if k % 6 else 9)   # k % 6 == 0: runs past the contig end
        text = orc.decode_exact(on_start, on_stop - on_start + 1)
        assert len(text) == on_stop - on_start
        o, l = corridor.corridor_endpoints(len(read), len(text), corridor.estimate_corridor(len(read), len(text)), realign=True)
        starts.append(on_start); stops.append(on_stop); refs.append(text); qrys.append(read.tobytes())
        offs.append(o); lens.append(l)
    batch = PackedBatch(refs, qrys, offs, lens)
    aligner.upload(batch)
    aligner.run()
    want = [r.as_dict() for r in aligner.fetch()]
    aligner.upload_windows(batch, starts, stops)
    aligner.run()
    got = [r.as_dict() for r in aligner.fetch()]
    from oracle_lib import same_alignment
    for a, b in zip(want, got):
        assert same_alignment(a, b) == []
    assert sum(1 for r in want if r["ret"] > 0) >= 20
    orc.close()


def test_batched_compute_alignment_on_the_device(aligner, oracle):
    """ngmlr_b200.intervals.compute_alignments through the GPU backend (decode_windows +
    convex_upload_windows + run + fetch, retries as further batches) == the same host logic over the
    CPU oracle (which tests/test_cs_oracle.py pins against AlignmentBuffer::computeAlignment)."""
    from ngmlr_b200 import refindex
    from ngmlr_b200.intervals import B200Backend, compute_alignments
    from oracle_lib import same_alignment
    from test_cs_oracle import _OracleBackend, _interval_tasks
    contigs = cs_cases.genome_contigs()
    orc = CsOracle([c.tobytes() for c in contigs])
    enc = refindex.encode_reference(contigs)
    aligner.set_reference(enc)
    want, want_calls = compute_alignments(_OracleBackend(orc, oracle), _interval_tasks(contigs, enc))
    got, got_calls = compute_alignments(B200Backend(aligner), _interval_tasks(contigs, enc))
    assert got_calls == want_calls
    for w, g in zip(want, got):
        assert (w is None) == (g is None)
        if w is not None:
            assert same_alignment(w, g.as_dict()) == []
    orc.close()
