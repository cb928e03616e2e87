"""GPU parity tests (B200): the CUDA convex path, called through the C ABI, against the oracle on
the same seeded inputs and against the committed golden vectors from the unmodified reference."""
import numpy as np
import pytest

import cases
import golden_util as gu
from ngmlr_b200 import PackedBatch, corridor, synth
from oracle_lib import DEFAULT_SCORING, same_alignment

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=[1, 0], ids=["small-batch-teams", "regular-kernels"])
def _small_batch_mode(request, aligner):
    """Every test of this module runs twice: with the default (batches of <= one problem per SM are filled by
    16-warp teams) and with the regular kernels (4-warp teams / one warp per problem) for every batch size."""
    aligner.set_small_batch_teams(request.param)
    yield
    aligner.set_small_batch_teams(1)


def _compare_batch(aligner, oracle, probs, scoring=DEFAULT_SCORING, rule=0, check_dirs=False):
    batch = PackedBatch.from_problems(probs)
    res = aligner.BatchAlign(batch)
    assert len(res) == len(probs)
    bad = []
    for i, (p, r) in enumerate(zip(probs, res)):
        want = oracle.single_align(p.ref, p.qry, p.offsets, p.lengths, p.ext_qstart, p.ext_qend,
                                   scoring=scoring, rule=rule)
        got = r.as_dict()
        d = same_alignment(want, got)
        if want["status"] != got["status"]:
            d.append("status")
        assert r.cells == p.cells
        if check_dirs and not (np.asarray(p.lengths) < 0).any():
            total = int(np.maximum(p.lengths, 0).sum())
            dg, bs, bx, by = aligner.debug_directions(i, total)
            do, os_, ox, oy = oracle.fill(p.ref, p.qry, p.offsets, p.lengths, scoring, rule)
            if not np.array_equal(dg, do):
                w = np.nonzero(dg != do)[0]
                d.append(f"dirs({len(w)} cells, first at {int(w[0])}: gpu {dg[w[0]]} oracle {do[w[0]]})")
            if (np.float32(bs).view(np.uint32), bx, by) != (np.float32(os_).view(np.uint32), ox, oy):
                d.append(f"best gpu {(bs, bx, by)} oracle {(os_, ox, oy)}")
        if d:
            bad.append((i, d, want["ret"], got["ret"]))
    assert not bad, f"{len(bad)} of {len(probs)} problems differ: {bad[:5]}"
    return res


def test_small_random_with_direction_matrix(aligner, oracle):
    _compare_batch(aligner, oracle, cases.random_problems(24, 101, max_len=500), check_dirs=True)


def test_edge_cases(aligner, oracle):
    _compare_batch(aligner, oracle, cases.edge_problems(), check_dirs=True)


def test_random_default_scoring(aligner, oracle):
    res = _compare_batch(aligner, oracle, cases.random_problems(96, 202, max_len=2500))
    assert sum(r.ret >= 0 for r in res) > 40


def test_single_align_matches_batch(aligner, oracle):
    p = cases.random_problems(1, 5, min_len=800, max_len=900, modes=(0,))[0]
    r = aligner.SingleAlign(p.ref, p.qry, p.offsets, p.lengths, 7, 9)
    want = oracle.single_align(p.ref, p.qry, p.offsets, p.lengths, 7, 9)
    assert same_alignment(want, r.as_dict()) == []
    assert r.ret == len(p.qry) + 7 + 9


def test_golden_vectors_from_reference(aligner):
    from ngmlr_b200 import B200Aligner
    for name, sc, probs, recs in gu.golden_sets():
        a = aligner if name == "default" else B200Aligner(0, scoring=sc)
        res = a.BatchAlign(PackedBatch.from_problems(probs))
        for i, (r, rec) in enumerate(zip(res, recs)):
            gu.check_against_record(r.as_dict(), rec, f"{name}[{i}]")
        if a is not aligner:
            a.close()


@pytest.mark.parametrize("sc", [cases.WEIRD_SCORING, cases.MILD_SCORING])
def test_non_default_scoring_uses_as_coded_sse_semantics(oracle, sc):
    from ngmlr_b200 import B200Aligner
    a = B200Aligner(0, scoring=sc)
    try:
        _compare_batch(a, oracle, cases.random_problems(32, 303, max_len=900), scoring=sc, rule=0,
                       check_dirs=True)
    finally:
        a.close()


def test_raw_kernel_equals_scalar_kernel_on_default_scoring(aligner, oracle):
    probs = cases.random_problems(32, 404, max_len=1200)
    aligner.force_raw(1)
    try:
        _compare_batch(aligner, oracle, probs, check_dirs=True)
    finally:
        aligner.force_raw(-1)


def test_pacbio_shaped_reads(aligner, oracle):
    """Config-2-shaped problems (8 kb reads, 15 % error, anchored corridor) at a size the oracle
    finishes in seconds."""
    probs = synth.pacbio_problems(6, genome_len=400_000, seed=2, median=6000)
    res = _compare_batch(aligner, oracle, probs)
    assert all(r.ret == len(p.qry) for r, p in zip(res, probs))


def test_properties_at_scale(aligner):
    """Size-independent properties on a batch the oracle would need minutes for: every CIGAR
    covers the whole read, NM/identity are consistent with the CIGAR+MD, the score is reproducible
    and independent of batch composition/order."""
    import re
    probs = synth.pacbio_problems(192, genome_len=2_000_000, seed=9, median=8000)
    batch = PackedBatch.from_problems(probs)
    res = list(aligner.BatchAlign(batch))  # materialise: the next batch call reuses the context's buffers
    for p, r in zip(probs, res):
        assert r.ret == len(p.qry), "valid alignment must cover the full read"
        ops = re.findall(r"(\d+)([MIDS])", r.pBuffer1)
        assert "".join(f"{n}{o}" for n, o in ops) == r.pBuffer1
        read_bases = sum(int(n) for n, o in ops if o in "MIS")
        assert read_bases == len(p.qry)
        ref_span = sum(int(n) for n, o in ops if o in "MD")
        assert r.lastPosition[0] == ref_span and r.PositionOffset + ref_span <= len(p.ref)
        md_mism = len(re.findall(r"[A-Z]", re.sub(r"\^[A-Z]+", "", r.pBuffer2)))
        indel = sum(int(n) for n, o in ops if o in "ID")
        assert r.NM == md_mism + indel
        assert r.cigarOpCount == len(ops)
        aln_cols = sum(int(n) for n, o in ops if o in "MID")
        assert r.alignmentLength == aln_cols
        assert np.float32(r.Identity) == np.float32(np.float32(aln_cols - r.NM) / np.float32(aln_cols))
        assert 0.6 < r.Identity <= 1.0 and r.Score > 0
    # order / batch-composition independence, bit-exact
    perm = np.random.default_rng(0).permutation(len(probs))[:64]
    res2 = aligner.BatchAlign(PackedBatch.from_problems([probs[i] for i in perm]))
    for j, i in enumerate(perm):
        assert same_alignment(res[i].as_dict(), res2[j].as_dict()) == []


def test_empty_batch(aligner):
    assert aligner.BatchAlign(PackedBatch([], [], [], [])) == []


@pytest.mark.parametrize("team", [0, 1])
def test_both_fill_schedules_are_bit_exact(aligner, oracle, team):
    """One warp per problem and the 4-warp team pipeline must give identical matrices, including
    on ragged / non-monotone / tiny corridors where team members wait on each other's strip."""
    aligner.force_team(team)
    try:
        _compare_batch(aligner, oracle, cases.edge_problems(), check_dirs=True)
        _compare_batch(aligner, oracle, cases.random_problems(40, 606, max_len=1500), check_dirs=True)
        probs = synth.pacbio_problems(4, genome_len=300_000, seed=12, median=5000)
        _compare_batch(aligner, oracle, probs)
    finally:
        aligner.force_team(-1)


def test_big_team_fill_is_bit_exact(aligner, oracle):
    """The few huge matrices of a batch are filled by 16-warp teams in a launch of their own, beside the
    ordinary one (thresholds lowered here so that ordinary test problems qualify): same matrices, same
    alignments -- in a mixed batch, and in a batch that consists of 'huge' problems only."""
    probs = (cases.random_problems(60, 717, max_len=1800, modes=(0, 2, 3))
             + synth.pacbio_problems(6, genome_len=300_000, seed=13, median=4000))
    try:
        aligner.debug_set_big_team(150_000, 100)
        _compare_batch(aligner, oracle, probs, check_dirs=True)
        big_only = [p for p in probs if p.cells >= 150_000 and p.lengths[0] >= 100][:12]
        assert len(big_only) >= 6
        _compare_batch(aligner, oracle, big_only, check_dirs=True)
    finally:
        aligner.debug_set_big_team(8 << 20, 768)


@pytest.mark.parametrize("team", [0, 1])
def test_fill_grid_cap_does_not_change_results(aligner, oracle, team):
    """ngmlr_b200_set_fill_ctas_per_sm: a persistent grid of 1 CTA per SM (every CTA walks many
    problems) gives the same matrices as full occupancy."""
    aligner.force_team(team)
    aligner.set_fill_ctas_per_sm(1)
    try:
        _compare_batch(aligner, oracle, cases.random_problems(300, 909, max_len=400), check_dirs=True)
    finally:
        aligner.set_fill_ctas_per_sm(0)
        aligner.force_team(-1)


def test_direction_arena_overflow_is_recovered(aligner, oracle):
    """The fill kernel bump-allocates its direction words; a too-small arena must be detected,
    grown and the batch re-run -- results unchanged."""
    probs = cases.random_problems(24, 707, max_len=1200)
    aligner.debug_set_arena_words(5000)
    try:
        _compare_batch(aligner, oracle, probs, check_dirs=True)
        assert aligner.stats()["fill_launches"] >= 2
    finally:
        aligner.debug_set_arena_words(-1)


def test_rows_wider_than_int16_use_the_as_coded_kernel(aligner, oracle):
    """Rows wider than 32767 cells: indelRun is a C `short` in the reference; such batches are routed
    to the RAW kernel, which keeps the 16-bit wrap. Full-matrix corridors over a 40 kb window."""
    rng = np.random.default_rng(5)
    g = synth.random_genome(41000, 77)
    probs = []
    for h in (40, 70):
        q = g[20000:20000 + h].copy()
        q[rng.integers(0, h, 3)] = ord("A")
        o, l = corridor.corridor_full(h, 40000)
        probs.append(synth.AlignProblem(g[:40000].tobytes(), q.tobytes(), o, l))
    # a homopolymer reference: one zero-score deletion run along the whole row
    o, l = corridor.corridor_full(34, 36000)
    probs.append(synth.AlignProblem(b"A" * 36000, b"C" * 34, o, l))
    _compare_batch(aligner, oracle, probs, check_dirs=True)
