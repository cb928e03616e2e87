"""CPU tests of host-side logic: corridor geometry mirrors and batch packing."""
import numpy as np

from ngmlr_b200 import PackedBatch, corridor, synth


def test_corridor_with_anchors_matches_c_float_semantics():
    # hand-computed with float32 arithmetic as in src/AlignmentBuffer.cpp:129-197
    offs, lens = corridor.corridor_endpoints_with_anchors(1000, 1100, [0, 300, 600], [0, 256, 512], 1)
    k = np.float32(1000) / np.float32(1100)
    diffs = [np.float32(y) / k - np.float32(x) for x, y in ((0, 0), (300, 256), (600, 512))]
    right = max([d for d in diffs if d > 0] + [np.float32(0)])
    left = max([-d for d in diffs if not d > 0] + [np.float32(0)])
    left = np.float32(left + 128)
    right = np.float32(right + 128)
    left = np.float32(left + np.float32(np.float32(left + right) * np.float32(0.1)))
    right = np.float32(right + np.float32(np.float32(left + right) * np.float32(0.1)))
    assert lens[0] == int(np.float32(left + right))
    assert offs[0] == int(np.float32(np.float32(0) / k) - right)
    assert offs[999] == int(np.float32(np.float32(999) / k) - right)
    assert (np.diff(offs) >= 0).all() and (lens == lens[0]).all()


def test_other_corridors():
    o, l = corridor.corridor_linear(10, 40)
    assert list(o[:3]) == [-20, -19, -18] and (l == 40).all()
    o, l = corridor.corridor_full(5, 1000)
    assert (o == -200).all() and (l == 1200).all()
    o, l = corridor.corridor_endpoints(100, 120, 400)
    assert (l == 100).all() and o[0] == int((np.float32(0) - np.float32(50)) / (np.float32(100) / np.float32(120)))
    assert corridor.estimate_corridor(8000, 8100) == max(int(np.float32(100) * np.float32(2.1)), 1600)


def test_packed_batch_layout():
    probs = synth.pacbio_problems(3, genome_len=100_000, seed=4, median=1500)
    b = PackedBatch.from_problems(probs)
    assert b.n == 3 and b.row_start[-1] == sum(len(p.qry) for p in probs)
    assert b.offsets.dtype == np.int32 and b.lengths.size == b.row_start[-1]
    assert b.read_bases == sum(len(p.qry) for p in probs)


def test_synthetic_reads_have_requested_error_profile():
    rng = np.random.default_rng(1)
    g = synth.random_genome(50_000, 1)
    read, starts = synth.mutate(g[:20000], rng, err=0.15, ratio=(9, 4, 2))
    # more insertions than deletions -> read longer than the reference window
    assert 20000 * 1.02 < read.size < 20000 * 1.12
    assert starts[-1] == read.size and (np.diff(starts) >= 0).all()


def test_select_candidates_matches_oracle_and_golden(oracle):
    """ngmlr_b200_select_candidates (host glue, no GPU) == ScoreBuffer::topNSE/computeMQ."""
    import golden_util as gu
    from ngmlr_b200 import select_candidates
    from oracle_lib import score_select_cases
    gold = gu.load("score_select_golden.json")
    scs = score_select_cases(77, 400)
    start = np.zeros(len(scs) + 1, dtype=np.int64)
    start[1:] = np.cumsum([s.size for s in scs])
    order, kept, mq = select_candidates(start, np.concatenate(scs))
    for i, (sc, g) in enumerate(zip(scs, gold)):
        local = order[start[i]:start[i + 1]] - start[i]
        assert (int(kept[i]), int(mq[i])) == (g["kept"], g["mq"])
        assert gu.digest(local.astype(np.int32)) == g["order_sha"]
        o = oracle.score_select(sc)
        assert np.array_equal(local, o[0]) and (int(kept[i]), int(mq[i])) == o[1:]
    # empty batch
    o2, k2, m2 = select_candidates(np.zeros(1, np.int64), np.zeros(0, np.float32))
    assert o2.size == 0 and k2.size == 0 and m2.size == 0
