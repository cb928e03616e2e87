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
