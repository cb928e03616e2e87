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


@pytest.fixture(scope="module")
def built(contigs):
    """(EncodedReference, KmerIndex) of the test genome from the numpy builder, built once."""
    from ngmlr_b200 import refindex
    ref = refindex.encode_reference(contigs)
    return ref, refindex.build_index(ref)


def test_numpy_index_builder_equals_oracle(contigs, cs_oracle, built):
    ref, idx = built
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


def test_cache_files_round_trip(contigs, tmp_path, built):
    """ngmfiles: what is written is read back identically (always runnable)."""
    from ngmlr_b200 import ngmfiles
    ref, idx = built
    ngmfiles.write_encoded_reference(str(tmp_path / "r-enc.2.ngm"), ref, skipped_lens=[8])
    ngmfiles.write_index(str(tmp_path / "r-ht-13-2.2.ngm"), idx)
    ref2, names = ngmfiles.read_encoded_reference(str(tmp_path / "r-enc.2.ngm"))
    assert np.array_equal(ref2.enc, ref.enc) and ref2.concat_len == ref.concat_len
    assert ref2.ref_start == ref.ref_start and ref2.ref_len == ref.ref_len and names[0] == "c0"
    idx2, skip, off = ngmfiles.read_index(str(tmp_path / "r-ht-13-2.2.ngm"))
    assert (idx2.k, skip, off) == (13, 2, 0)
    assert np.array_equal(idx2.tab, idx.tab) and np.array_equal(idx2.rci, idx.rci) and np.array_equal(idx2.pos, idx.pos)
    with open(tmp_path / "bad.ngm", "wb") as f:
        f.write(b"\0" * 64)
    with pytest.raises(ValueError):
        ngmfiles.read_index(str(tmp_path / "bad.ngm"))


@pytest.mark.skipif(not CsReference.available(), reason="oracle/_ref/libngmlr_full.so not built")
def test_cache_files_are_byte_compatible_with_the_reference(contigs, tmp_path, built):
    """The UNMODIFIED reference writes <fasta>-enc.2.ngm and <fasta>-ht-13-2.2.ngm (in a subprocess:
    its singletons initialise once per process); ngmfiles reads them into the arrays refindex builds,
    and writes the same bytes (the index file entirely; the encoded reference up to the end of the used
    part of binRef -- the reference writes its uninitialised allocation tail)."""
    import subprocess
    import sys
    from ngmlr_b200 import ngmfiles, refindex
    fasta = str(tmp_path / "ref.fa")
    with open(fasta, "w") as f:
        for i, c in enumerate(contigs):
            f.write(f">contig{i} some description\n{c.tobytes().decode()}\n")
    code = ("import ctypes, sys; lib = ctypes.CDLL(sys.argv[1]); "
            "sys.exit(lib.ref_cs_init_save(sys.argv[2].encode()))")
    run = subprocess.run([sys.executable, "-c", code, CsReference.PATH, fasta], capture_output=True, text=True,
                         timeout=600)
    assert run.returncode == 0, run.stderr[-2000:]
    enc_path, idx_path = fasta + "-enc.2.ngm", fasta + "-ht-13-2.2.ngm"
    ref, idx = built
    got_ref, names = ngmfiles.read_encoded_reference(enc_path)
    assert np.array_equal(got_ref.enc, ref.enc) and got_ref.concat_len == ref.concat_len
    assert got_ref.ref_start == ref.ref_start and got_ref.ref_len == ref.ref_len
    assert names == [f"contig{i}" for i in range(len(contigs)) if contigs[i].size > 10]
    got_idx, skip, off = ngmfiles.read_index(idx_path)
    assert (got_idx.k, skip, off) == (13, 2, 0)
    assert np.array_equal(got_idx.tab, idx.tab) and np.array_equal(got_idx.rci, idx.rci)
    assert np.array_equal(got_idx.pos, idx.pos)
    # writers: same bytes
    ngmfiles.write_index(str(tmp_path / "mine-ht.ngm"), idx)
    assert open(tmp_path / "mine-ht.ngm", "rb").read() == open(idx_path, "rb").read()
    skipped = [int(c.size) for c in contigs if not c.size > 10]
    ngmfiles.write_encoded_reference(str(tmp_path / "mine-enc.ngm"), ref, names, skipped_lens=skipped)
    mine, theirs = open(tmp_path / "mine-enc.ngm", "rb").read(), open(enc_path, "rb").read()
    defined = 24 + 128 * len(ref.ref_start) + int(ref.enc.size)
    assert len(mine) == len(theirs) and mine[:defined] == theirs[:defined]
