"""GPU parity tests (B200) of the k-mer index built on the device (cs_index_build.cu =
CompactPrefixTable::CreateTable): Index records and Location lists bit-identical with the C oracle (pinned
to the unmodified reference) and the numpy builder, on the edge-case genome (N runs, homopolymers, tandem
repeats, contigs that start / end with N, a final run of exactly k characters, an odd-length contig) and on
a genome with planted repeats where the frequency cutoff (maxPrefixFreq = 1000, src/PrefixTable.cpp:28,298)
fires and where slots are allocated but never used (990 < frequency < 1000); candidate search on the
device-built index equals the oracle; the cache file written from it is byte-identical."""
import numpy as np
import pytest

import cs_cases
from ngmlr_b200 import B200Aligner, ngmfiles, refindex, synth
from oracle_lib import CsOracle

pytestmark = pytest.mark.gpu


def _repeat_genome():
    """Planted 40-bp motifs on a grid whose spacing is a multiple of kmerSkip + 1 = 3, so that every copy is
    sampled at the same k-mers: 1300 copies (frequency >= 1000: no slots), 995 copies (slots allocated,
    m_RevCompIndex truncates to 0: never used), 400 copies (used)."""
    g = synth.random_genome(1_600_000, 99)
    at = 3000
    for motif_seed, copies in ((1, 1300), (2, 995), (3, 400)):
        motif = synth.random_genome(40, 1000 + motif_seed)
        for _ in range(copies):
            g[at:at + 40] = motif
            at += 300
    g[900_000:900_050] = ord("N")
    return [g[:1_000_000], g[1_000_000:]]


@pytest.mark.parametrize("which", ["edge", "repeats"])
def test_device_index_equals_oracle_and_numpy(which):
    contigs = cs_cases.genome_contigs() if which == "edge" else _repeat_genome()
    enc = refindex.encode_reference(contigs)
    want = refindex.build_index(enc)
    orc = CsOracle([c.tobytes() for c in contigs])
    al = B200Aligner(0)
    try:
        tab, rci, pos = orc.index()
        assert np.array_equal(tab, want.tab) and np.array_equal(rci, want.rci) and np.array_equal(pos, want.pos)
        al.set_reference(enc)
        got = al.build_index(enc, fetch=True)
        assert np.array_equal(got.tab, want.tab)
        assert np.array_equal(got.rci, want.rci)
        assert got.pos.size == want.pos.size and np.array_equal(got.pos, want.pos)
        if which == "repeats":
            freq = np.diff(want.tab.astype(np.int64))
            assert (want.rci[:-1] == 0)[freq > 0].any(), "no allocated-but-unused k-mer in the test genome"
        # the candidate search runs on the device-built index
        subs = cs_cases.subreads(60, 5, contigs) if which == "edge" else \
            [contigs[0][s:s + 256].tobytes() for s in range(5000, 900_000, 30_011)]
        res, mx = al.cs_search(subs)
        for sub, r, m in zip(subs, res, mx):
            w, wm = orc.search(sub)
            assert [(float(a), int(b), int(c)) for a, b, c in r] == [(float(a), int(b), int(c)) for a, b, c in w]
            assert float(m) == float(wm)
    finally:
        al.close()
        orc.close()


def test_cache_file_from_device_index_is_byte_identical(tmp_path):
    contigs = cs_cases.genome_contigs()
    enc = refindex.encode_reference(contigs)
    want = refindex.build_index(enc)
    al = B200Aligner(0)
    try:
        al.set_reference(enc)
        got = al.build_index(enc, fetch=True)
    finally:
        al.close()
    a, b = str(tmp_path / "a-ht-13-2.2.ngm"), str(tmp_path / "b-ht-13-2.2.ngm")
    ngmfiles.write_index(a, want, skip=2)
    ngmfiles.write_index(b, got, skip=2)
    assert open(a, "rb").read() == open(b, "rb").read()


def test_device_reference_encoding_equals_numpy_and_feeds_the_index(tmp_path):
    """_SequenceProvider::Init's encoding on the device (ngmlr_b200_cs_encode_reference): same bytes, contig table
    and concat length as refindex.encode_reference (itself held equal to the unmodified reference's binRef,
    tests/test_cs_oracle.py) on contigs with lower case, IUPAC codes, N runs, odd lengths, a contig of <= 10
    characters (skipped) and one of 11; the index built from it and the cache files written through the C writers
    are the numpy builder's."""
    contigs = [bytes(c) for c in cs_cases.genome_contigs()]
    rng = np.random.default_rng(3)
    messy = bytearray(rng.choice(np.frombuffer(b"ACGTacgtNnRYKMSWrykm-*", dtype=np.uint8), 4097).tobytes())
    contigs = [contigs[0], b"ACGTACGTAC", bytes(messy), b"ACGTACGTACG"] + contigs[1:]
    want = refindex.encode_reference([np.frombuffer(c, dtype=np.uint8) for c in contigs])
    al = B200Aligner(0)
    try:
        got = al.encode_reference(contigs)
        assert got.ref_start == want.ref_start and got.ref_len == want.ref_len and got.concat_len == want.concat_len
        assert np.array_equal(got.enc, want.enc)
        idx = al.build_index(got, fetch=True)
        ref_idx = refindex.build_index(want)
        assert np.array_equal(idx.tab, ref_idx.tab) and np.array_equal(idx.rci, ref_idx.rci)
        assert np.array_equal(idx.pos, ref_idx.pos)
        # windows decoded from the device-encoded genome (refStartPos was installed with it)
        starts = [want.ref_start[1] + 5, want.ref_start[2] - 20]
        assert al.decode_windows(starts, [200, 64]) == _numpy_windows(want, starts, [200, 64], al)
        ngmfiles.c_write_encoded_reference(str(tmp_path / "a-enc.2.ngm"), got, skipped_lens=[10])
        ngmfiles.write_encoded_reference(str(tmp_path / "b-enc.2.ngm"), want, skipped_lens=[10])
        assert open(tmp_path / "a-enc.2.ngm", "rb").read() == open(tmp_path / "b-enc.2.ngm", "rb").read()
        ngmfiles.c_write_index(str(tmp_path / "a-ht.ngm"), idx)
        ngmfiles.write_index(str(tmp_path / "b-ht.ngm"), ref_idx)
        assert open(tmp_path / "a-ht.ngm", "rb").read() == open(tmp_path / "b-ht.ngm", "rb").read()
    finally:
        al.close()


def _numpy_windows(enc, starts, lens, al):
    """The same windows from a context whose reference was uploaded the usual way."""
    other = B200Aligner(0)
    try:
        other.set_reference(enc)
        return other.decode_windows(starts, lens)
    finally:
        other.close()
