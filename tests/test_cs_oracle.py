"""CPU tests of the candidate-search row: the C restatement (oracle/cs_oracle.c) and the numpy
builder of the reference-format index (ngmlr_b200/refindex.py) against the WHOLE unmodified
reference (oracle/_ref/libngmlr_full.so: SequenceProvider, CompactPrefixTable, CS) when present,
and against each other always."""
import numpy as np
import pytest

import cs_cases
from oracle_lib import CsOracle, CsReference


@pytest.fixture(scope="module")
def contigs():
    return cs_cases.genome_contigs()


@pytest.fixture(scope="module")
def cs_oracle(contigs):
    o = CsOracle([c.tobytes() for c in contigs])
    yield o
    o.close()


def test_numpy_index_builder_equals_oracle(contigs, cs_oracle):
    from ngmlr_b200 import refindex
    ref = refindex.encode_reference(contigs)
    idx = refindex.build_index(ref)
    assert ref.concat_len == cs_oracle.concat_len and ref.ref_start == cs_oracle.ref_starts()
    assert np.array_equal(ref.enc, cs_oracle.encoded())
    tab, rci, pos = cs_oracle.index()
    assert np.array_equal(tab, idx.tab) and np.array_equal(rci, idx.rci) and np.array_equal(pos, idx.pos)
    packed = idx.packed_index().reshape(-1, 5)
    assert packed.shape[0] == 4 ** 13 + 1
    assert np.array_equal(packed[:, :4].copy().view(np.uint32).reshape(-1), tab)


def test_oracle_search_basic_properties(contigs, cs_oracle):
    g1 = contigs[0]
    res, mx = cs_oracle.search(g1[20000:20256].tobytes())
    assert res[0][1] // 16 == (1000 + 20000) // 16 and res[0][2] == 0 and mx == res[0][0] and 80 <= mx <= 82
    from ngmlr_b200 import synth
    res, mx = cs_oracle.search(synth.revcomp(g1[30000:30256]).tobytes())
    assert res[0][2] == 1 and abs(res[0][1] - (1000 + 30000)) <= 16
    assert cs_oracle.search(b"")[0] == [] and cs_oracle.search(b"ACGTACGTACGT")[0] == []


@pytest.mark.skipif(not CsReference.available(), reason="oracle/_ref/libngmlr_full.so not built")
def test_oracle_equals_whole_reference(contigs, cs_oracle, tmp_path_factory):
    path = str(tmp_path_factory.mktemp("cs") / "ref.fa")
    with open(path, "w") as f:
        for i, c in enumerate(contigs):
            f.write(f">c{i}\n{c.tobytes().decode()}\n")
    ref = CsReference(path)
    assert ref.concat_len == cs_oracle.concat_len and ref.ref_starts() == cs_oracle.ref_starts()
    a, b = ref.index(), cs_oracle.index()
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    for sub in cs_cases.subreads(250, 17, contigs):
        if len(sub) == 0:
            continue
        for bits in (16, 8):   # table size must not matter (overflow -> retry with a larger table)
            assert ref.search(sub, table_bits=bits) == cs_oracle.search(sub)
    for p in (0, 1, 999, 1000, 1001, 5555, 61000, ref.concat_len - 100, ref.concat_len - 1, ref.concat_len):
        assert ref.decode(p) == cs_oracle.decode(p)
    # DecodeRefSequenceExact (alignment windows, corridor 0 as in extractReferenceSequenceForAlignment)
    from ngmlr_b200 import refindex
    enc = refindex.encode_reference(contigs)
    wins = cs_cases.exact_windows(enc.ref_start, enc.ref_len)
    assert len(wins) > 500
    for st, ln in wins:
        assert ref.decode_exact(st, ln) == cs_oracle.decode_exact(st, ln), (st, ln)
    assert ref.decode_exact(ref.concat_len, 10) is None and cs_oracle.decode_exact(ref.concat_len, 10) is None
    for st, ln, c in ((5000, 100, 12), (5001, 101, 12), (61000, 64, 40)):   # non-zero corridor as well
        assert ref.decode_exact(st, ln, c) == cs_oracle.decode_exact(st, ln, c)


def test_oracle_decode_exact_matches_golden(contigs, cs_oracle):
    import golden_util as gu
    from ngmlr_b200 import refindex
    g = gu.load("decode_exact_golden.json")
    enc = refindex.encode_reference(contigs)
    wins = cs_cases.exact_windows(enc.ref_start, enc.ref_len)
    assert len(wins) == g["n"] and gu.digest(np.array(wins, dtype=np.int64)) == g["windows_sha"], "generator drifted"
    for (st, ln), sha in zip(wins, g["text_sha"]):
        assert gu.digest(cs_oracle.decode_exact(st, ln)) == sha, (st, ln)
    for st, ln, text in g["samples"]:
        assert cs_oracle.decode_exact(st, ln) == text.encode()
