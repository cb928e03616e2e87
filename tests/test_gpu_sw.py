"""GPU parity tests for the sub-read scorer (StrippedSW replacement) through the C ABI."""
import numpy as np
import pytest

import cases
import golden_util as gu

pytestmark = pytest.mark.gpu


def test_sw_matches_oracle_and_golden(aligner, oracle):
    refs, qrys = cases.sw_pairs(256, 31)
    got = aligner.BatchScore(refs, qrys)
    want = np.array([oracle.ssw_score(r, q) for r, q in zip(refs, qrys)], dtype=np.float32)
    assert np.array_equal(got, want)
    g = gu.load("sw_golden.json")
    assert np.array_equal(got, np.array(g["scores"], dtype=np.float32))


def test_sw_full_score_batch(aligner, oracle):
    """swBatchSize = 1024 pairs, the unit ScoreBuffer hands to BatchScore (StrippedSW.h:53-55)."""
    refs, qrys = cases.sw_pairs(1024, 77)
    got = aligner.BatchScore(refs, qrys)
    want = np.array([oracle.ssw_score(r, q) for r, q in zip(refs, qrys)], dtype=np.float32)
    assert np.array_equal(got, want)
    assert got.max() >= 250  # error-free sub-reads score ~256


def test_sw_long_inputs_multi_pass_and_limits(aligner, oracle):
    g = gu.load("sw_golden.json")
    long_ref = b"ACGT" * 700
    extra = [(long_ref, long_ref[:2000]), (long_ref, long_ref[:1000] + b"G" + long_ref[1000:2000]),
             (b"A" * 100001, b"A" * 10), (b"", b""), (b"A", b"A"), (b"ACGT", b"")]
    got = aligner.BatchScore([r for r, _ in extra], [q for _, q in extra])
    assert list(got) == g["extra"]
    rng = np.random.default_rng(3)
    from ngmlr_b200 import synth
    gen = synth.random_genome(20000, 1)
    refs, qrys = [], []
    for _ in range(12):
        s = int(rng.integers(0, 8000))
        L = int(rng.integers(300, 5000))
        q, _m = synth.mutate(gen[s:s + L], rng, err=0.05)
        refs.append(gen[s:s + L + 200].tobytes())
        qrys.append(q.tobytes())
    got = aligner.BatchScore(refs, qrys)
    want = [oracle.ssw_score(r, q) for r, q in zip(refs, qrys)]
    assert list(got) == want


def test_single_score(aligner, oracle):
    refs, qrys = cases.sw_pairs(4, 5)
    for r, q in zip(refs, qrys):
        assert aligner.SingleScore(r, q) == oracle.ssw_score(r, q)
