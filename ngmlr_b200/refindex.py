"""Host-side reference layout and k-mer index in the reference's in-memory format (numpy).

The CS kernel consumes exactly the arrays ngmlr holds in memory, so that a maintainer can hand
them over unchanged:
  * the encoded genome `binRef` (2 bases/byte, A0 T1 G2 C3 N4, 1000-N spacers before, between and
    after contigs)                          <- _SequenceProvider::Init, src/SequenceProvider.cpp:292-400
  * `Index{uint m_TabIndex; char m_RevCompIndex}` x (4^k + 1) and `Location{uint}` lists
                                            <- CompactPrefixTable, src/PrefixTable.{h,cpp}
This module builds the same arrays for synthetic genomes (tests, bench.py) without the reference;
tests/test_cs_oracle.py checks it against the oracle and the unmodified reference. Index
construction itself is SURVEY section 8(f).3 ("next") -- this is input preparation, not a kernel.
"""
from dataclasses import dataclass

import numpy as np

ENC4 = np.full(256, 4, dtype=np.uint8)
for _c, _v in zip(b"ATGCatgc", (0, 1, 2, 3, 0, 1, 2, 3)):
    ENC4[_c] = _v
DEC4 = np.frombuffer(b"ATGCN", dtype=np.uint8)


@dataclass
class EncodedReference:
    enc: np.ndarray          # uint8 binRef
    concat_len: int          # GetConcatRefLen()
    ref_start: list          # concat position of each kept contig
    ref_len: list

    def base_codes(self, start, n):
        """enc4 codes of concat positions [start, start+n)."""
        p = np.arange(start, start + n, dtype=np.int64)
        b = self.enc[p >> 1]
        return np.where(p & 1, b & 0xF, b >> 4).astype(np.uint8)


def encode_reference(contigs):
    """contigs: list of uint8 arrays / bytes. Contigs of <= 10 bases are skipped (minRefSeqLen)."""
    spacer = np.full(500, (4 << 4) | 4, dtype=np.uint8)
    parts = [spacer]
    starts, lens = [], []
    nbytes = 500
    for c in contigs:
        c = np.frombuffer(bytes(c), dtype=np.uint8) if not isinstance(c, np.ndarray) else c
        L = int(c.size)
        if not L > 10:
            continue
        starts.append(nbytes * 2)
        lens.append(L)
        codes = ENC4[c]
        if L & 1:
            codes = np.concatenate([codes, [4]]).astype(np.uint8)
        parts.append(((codes[0::2] << 4) | codes[1::2]).astype(np.uint8))
        parts.append(spacer)
        nbytes += (L + 1) // 2 + 500
    enc = np.concatenate(parts)
    return EncodedReference(enc, int(enc.size) * 2 - 1, starts, lens)


def decode_contigs(ref: EncodedReference):
    """The contigs of an encoded reference as characters (A C G T N), e.g. to simulate reads from a
    reference that arrived packed (parallel.broadcast_reference)."""
    return [DEC4[np.minimum(ref.base_codes(s, n), 4)] for s, n in zip(ref.ref_start, ref.ref_len)]


def _revcomp_codes(prefix, k):
    """revComp of src/PrefixTable.cpp:70-88 for the CS k-mer code (A0 C1 T2 G3)."""
    mask = (1 << (2 * k)) - 1
    c = (prefix.astype(np.uint64) ^ np.uint64(0xAAAAAAAAAAAAAAAA)) & np.uint64(mask)
    r = np.zeros_like(c)
    for _ in range(k):
        r = (r << np.uint64(2)) | (c & np.uint64(3))
        c >>= np.uint64(2)
    return r


def _kmer_callbacks(seq, offset, k, skip):
    """(prefix, pos) for every callback CS::PrefixIteration (src/CSstatic.cpp:23-73) makes on `seq`
    (uint8 chars) with prefixskip=skip, offset=offset. Vectorised per N-free run."""
    L = int(seq.size)
    code = ((seq >> 1) & 3).astype(np.uint64)
    is_n = seq == ord("N")
    out_p, out_x = [], []
    # emulate the tail recursion: (cursor, length) segments
    cur, length, off = 0, L, offset
    nidx = np.flatnonzero(is_n)
    while True:
        if length < k:
            break
        if is_n[cur]:
            # leading N run
            j = cur
            while j < L and is_n[j]:
                j += 1
            n_skip = j - cur
            if n_skip >= length - k:
                break
            cur, length, off = j, length - n_skip, off + n_skip
        # next N at or after cur within this segment
        ii = np.searchsorted(nidx, cur)
        stop = int(nidx[ii]) if ii < nidx.size and nidx[ii] < cur + length else cur + length
        run = stop - cur  # N-free run length starting at cur
        if run >= k:
            cw = code[cur:stop]
            # rolling k-mer codes for positions 0..run-k
            pref = np.zeros(run - k + 1, dtype=np.uint64)
            for t in range(k):
                pref = (pref << np.uint64(2)) | cw[t:t + run - k + 1]
            sel = np.arange(0, run - k + 1, skip + 1)
            out_p.append(pref[sel])
            out_x.append((off + sel).astype(np.uint64))
        if stop >= cur + length:
            break
        # restart after the N at `stop`: PrefixIteration(sequence+i+1, length-i-1, offset+i+1)
        i = stop - cur
        cur, length, off = stop + 1, length - i - 1, off + i + 1
    if not out_p:
        return np.zeros(0, np.uint64), np.zeros(0, np.uint64)
    return np.concatenate(out_p), np.concatenate(out_x)


@dataclass
class KmerIndex:
    k: int
    bin_shift: int
    tab: np.ndarray   # uint32[4^k + 1]  Index::m_TabIndex
    rci: np.ndarray   # int8[4^k + 1]    Index::m_RevCompIndex (used() <=> != 0)
    pos: np.ndarray   # uint32[n]        Location::m_Location (unit offset 0)

    def packed_index(self):
        """The reference's packed 5-byte `Index` records (#pragma pack(1), src/PrefixTable.h:17-35)."""
        raw = np.zeros((self.tab.size, 5), dtype=np.uint8)
        raw[:, :4] = self.tab.view(np.uint8).reshape(-1, 4)
        raw[:, 4] = self.rci.view(np.uint8)
        return raw.reshape(-1)


def build_index(ref: EncodedReference, k=13, skip=2, bin_shift=4, max_freq=1000):
    """CompactPrefixTable construction (CountKmer / createRefTableIndex / BuildPrefixTable,
    src/PrefixTable.cpp:269-321, 372-437) for one table unit (< 4 G positions)."""
    n_idx = (1 << (2 * k)) + 1
    prefs, poss, keeps = [], [], []
    for start, L in zip(ref.ref_start, ref.ref_len):
        chars = DEC4[ref.base_codes(start, L)].copy()
        # Generate() decodes with a buffer length that loses the last 2 characters ('x'/NUL -> code 0)
        if L >= 2:
            chars[L - 2:] = 0
        p, x = _kmer_callbacks(chars, start, k, skip)
        if p.size == 0:
            continue
        same = np.zeros(p.size, dtype=bool)
        same[1:] = p[1:] == p[:-1]
        b = (x >> np.uint64(bin_shift)).astype(np.int64)
        # position within a streak of equal consecutive prefixes
        idx = np.arange(p.size)
        streak_start = np.maximum.accumulate(np.where(~same, idx, 0))
        spos = idx - streak_start
        prev_b = np.empty_like(b)
        prev_b[0] = -1
        prev_b[1:] = b[:-1]
        keep = (spos < 2) | (b != prev_b)
        prefs.append(p)
        poss.append(x)
        keeps.append(keep)
    if prefs:
        p = np.concatenate(prefs)
        x = np.concatenate(poss)
        keep = np.concatenate(keeps)
    else:
        p = np.zeros(0, np.uint64)
        x = np.zeros(0, np.uint64)
        keep = np.zeros(0, bool)
    pk, xk = p[keep].astype(np.int64), x[keep]
    freq = np.bincount(pk, minlength=n_idx - 1).astype(np.int64)
    allp = np.arange(n_idx - 1, dtype=np.uint64)
    total = freq + freq[_revcomp_codes(allp, k).astype(np.int64)]
    alloc = (freq > 0) & (total < max_freq)
    rci = np.zeros(n_idx, dtype=np.int8)
    val = (np.float32(max_freq) - total[alloc].astype(np.float32)) * np.float32(100.0) / np.float32(max_freq)
    rci[:-1][alloc] = val.astype(np.int32).astype(np.int8)  # float -> char truncation
    tab = np.ones(n_idx, dtype=np.uint32)
    csum = np.cumsum(np.where(alloc, freq, 0))
    tab[1:] = (csum + 1).astype(np.uint32)
    npos = int(csum[-1]) if csum.size else 0
    pos = np.zeros(npos, dtype=np.uint32)
    used = rci[:-1] != 0
    sel = used[pk]
    pk2, xk2 = pk[sel], xk[sel]
    order = np.argsort(pk2, kind="stable")
    pk2, xk2 = pk2[order], xk2[order]
    # slot = tab[prefix]-1 + rank within prefix
    first = np.searchsorted(pk2, pk2, side="left")
    slot = tab[pk2].astype(np.int64) - 1 + (np.arange(pk2.size) - first)
    pos[slot] = xk2.astype(np.uint32)
    return KmerIndex(k, bin_shift, tab, rci, pos)
