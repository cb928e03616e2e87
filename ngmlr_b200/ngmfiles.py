"""Readers / writers for ngmlr's on-disk caches, byte-compatible with the reference, so that the
arrays ngmlr already has on disk can be handed to the device pipeline unchanged (and vice versa):

  <ref>-enc.2.ngm        encoded reference  _SequenceProvider::writeEncRefToFile / readEncRefFromFile
                         (src/SequenceProvider.cpp:207-272): uint cookie 0x74656, uint refCount,
                         uloc binRefIndex (= 2 * used bytes), uloc encRefSize (allocated bytes),
                         RefIdx[refCount] (128 bytes each, src/SequenceProvider.h:56-63), binRef bytes.
  <ref>-ht-<k>-<skip>.2.ngm  k-mer index  CompactPrefixTable::saveToFile / readFromFile
                         (src/PrefixTable.cpp:534-630): uint cookie 0x1701E, prefix length, ref skip,
                         unit count, index size; per unit: uint cRefTableLen, Index[index size]
                         (5 bytes each), Location[cRefTableLen], uloc Offset; uint signature.

This module is the tests' independent restatement of the formats (numpy); the product's own writers / readers are
the C entry points ngmlr_b200_ngm_* (csrc/ngm_files.cpp), bound below as c_write_index / c_read_index /
c_write_encoded_reference / c_read_encoded_reference and held equal to these byte for byte.
"""
import struct

import numpy as np

from .refindex import EncodedReference, KmerIndex

REF_ENC_COOKIE = 0x74656
REF_TAB_COOKIE = 0x1701E
REFIDX = np.dtype([("SeqId", "<u4"), ("Flags", "<u4"), ("SeqStart", "<u8"), ("SeqLen", "<u4"),
                   ("NameLen", "<u4"), ("name", "S100"), ("pad", "V4")])
assert REFIDX.itemsize == 128


def read_encoded_reference(path):
    """-> (EncodedReference, [contig names]). Only the used part of binRef is returned (the reference
    writes its whole allocation; bytes past binRefIndex / 2 are uninitialised there)."""
    with open(path, "rb") as f:
        cookie, ref_count, bin_ref_index, enc_size = struct.unpack("<IIQQ", f.read(24))
        if cookie != REF_ENC_COOKIE:
            raise ValueError(f"{path}: not an encoded reference (cookie {cookie:#x})")
        idx = np.frombuffer(f.read(REFIDX.itemsize * ref_count), dtype=REFIDX)
        enc = np.frombuffer(f.read(enc_size), dtype=np.uint8)
    if enc.size != enc_size or bin_ref_index % 2 or bin_ref_index // 2 > enc_size:
        raise ValueError(f"{path}: truncated or inconsistent")
    used = bin_ref_index // 2
    names = [bytes(r["name"])[:int(r["NameLen"])].decode("ascii", "replace") for r in idx]
    ref = EncodedReference(enc[:used].copy(), int(bin_ref_index) - 1, [int(v) for v in idx["SeqStart"]],
                           [int(v) for v in idx["SeqLen"]])
    return ref, names


def reference_alloc_bytes(ref_lens, skipped_lens=()):
    """binRefSize the reference allocates (and writes): ((size / 2) | 1) + 1 with getSize() = 1000 +
    sum((len | 1) + 1 + 1000) over ALL sequences of the FASTA file, kept or skipped
    (src/SequenceProvider.cpp:274-290, 318)."""
    size = 1000
    for n in list(ref_lens) + list(skipped_lens):
        size += (int(n) | 1) + 1 + 1000
    return ((size // 2) | 1) + 1


def write_encoded_reference(path, ref, names=None, skipped_lens=()):
    """Writes what the reference writes for the same FASTA file; the unused tail of binRef (garbage in
    the reference's file) is written as zeros."""
    n = len(ref.ref_start)
    names = names or [f"c{i}" for i in range(n)]
    idx = np.zeros(n, dtype=REFIDX)
    idx["SeqId"] = np.arange(n)
    idx["SeqStart"] = ref.ref_start
    idx["SeqLen"] = ref.ref_len
    for i, nm in enumerate(names):
        b = nm.encode()[:100]
        idx["name"][i] = b
        idx["NameLen"][i] = len(b)
    alloc = max(reference_alloc_bytes(ref.ref_len, skipped_lens), int(ref.enc.size))
    with open(path, "wb") as f:
        f.write(struct.pack("<IIQQ", REF_ENC_COOKIE, n, 2 * int(ref.enc.size), alloc))
        f.write(idx.tobytes())
        f.write(np.ascontiguousarray(ref.enc, dtype=np.uint8).tobytes())
        f.write(b"\0" * (alloc - int(ref.enc.size)))


def read_index(path, bin_shift=4):
    """-> (KmerIndex of unit 0, ref_skip, unit_offset). Multi-unit tables (genomes beyond 4 G positions)
    are refused: the device search takes one unit."""
    with open(path, "rb") as f:
        data = f.read()
    cookie, k, skip, units, index_size = struct.unpack_from("<IIIII", data, 0)
    if cookie != REF_TAB_COOKIE:
        raise ValueError(f"{path}: not a reference index (cookie {cookie:#x})")
    (sig,) = struct.unpack_from("<I", data, len(data) - 4)
    if sig != (cookie + k + skip + units + index_size) & 0xFFFFFFFF:
        raise ValueError(f"{path}: signature mismatch (the reference would rebuild)")
    if units != 1:
        raise ValueError(f"{path}: {units} table units; only single-unit tables are supported")
    at = 20
    (n_pos,) = struct.unpack_from("<I", data, at)
    at += 4
    raw = np.frombuffer(data, dtype=np.uint8, count=index_size * 5, offset=at).reshape(-1, 5)
    at += index_size * 5
    tab = raw[:, :4].copy().view("<u4").reshape(-1)
    rci = raw[:, 4].copy().view(np.int8)
    pos = np.frombuffer(data, dtype="<u4", count=n_pos, offset=at).copy()
    at += n_pos * 4
    (unit_offset,) = struct.unpack_from("<Q", data, at)
    return KmerIndex(int(k), bin_shift, tab, rci, pos), int(skip), int(unit_offset)


def write_index(path, idx, skip=2, unit_offset=0):
    index_size = int(idx.tab.size)
    with open(path, "wb") as f:
        f.write(struct.pack("<IIIII", REF_TAB_COOKIE, idx.k, skip, 1, index_size))
        f.write(struct.pack("<I", int(idx.pos.size)))
        f.write(idx.packed_index().tobytes())
        f.write(np.ascontiguousarray(idx.pos, dtype="<u4").tobytes())
        f.write(struct.pack("<Q", unit_offset))
        f.write(struct.pack("<I", (REF_TAB_COOKIE + idx.k + skip + 1 + index_size) & 0xFFFFFFFF))


# ---- the library's C writers / readers (csrc/ngm_files.cpp) -------------------------------------------------------
def _c():
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    u32p, u64p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    lib.ngmlr_b200_ngm_write_index.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p,
                                               C.c_uint32, C.c_uint64]
    lib.ngmlr_b200_ngm_read_index.argtypes = [C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), u32p, u32p, u64p,
                                              C.c_void_p, C.c_void_p]
    lib.ngmlr_b200_ngm_write_reference.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p,
                                                   C.c_void_p, C.POINTER(C.c_char_p)]
    lib.ngmlr_b200_ngm_read_reference.argtypes = [C.c_char_p, u32p, u64p, u64p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                  C.c_void_p]
    return C, lib


def c_write_index(path, idx, skip=2, unit_offset=0):
    C, lib = _c()
    packed = np.ascontiguousarray(idx.packed_index())
    pos = np.ascontiguousarray(idx.pos, dtype="<u4")
    rc = lib.ngmlr_b200_ngm_write_index(str(path).encode(), int(idx.k), int(skip), packed.ctypes.data, int(idx.tab.size),
                                        pos.ctypes.data, int(pos.size), int(unit_offset))
    if rc:
        raise OSError(f"ngmlr_b200_ngm_write_index({path}) = {rc}")


def c_read_index(path, bin_shift=4):
    C, lib = _c()
    k, skip, n_idx, n_pos, off = C.c_int32(), C.c_int32(), C.c_uint32(), C.c_uint32(), C.c_uint64()
    args = (str(path).encode(), C.byref(k), C.byref(skip), C.byref(n_idx), C.byref(n_pos), C.byref(off))
    rc = lib.ngmlr_b200_ngm_read_index(*args, None, None)
    if rc:
        raise ValueError(f"ngmlr_b200_ngm_read_index({path}) = {rc}")
    raw = np.zeros((n_idx.value, 5), np.uint8)
    pos = np.zeros(n_pos.value, "<u4")
    rc = lib.ngmlr_b200_ngm_read_index(*args, raw.ctypes.data, pos.ctypes.data)
    if rc:
        raise ValueError(f"ngmlr_b200_ngm_read_index({path}) = {rc}")
    tab = raw[:, :4].copy().view("<u4").reshape(-1)
    rci = raw[:, 4].copy().view(np.int8)
    return KmerIndex(int(k.value), bin_shift, tab, rci, pos), int(skip.value), int(off.value)


def c_write_encoded_reference(path, ref, names=None, skipped_lens=()):
    C, lib = _c()
    n = len(ref.ref_start)
    enc = np.ascontiguousarray(ref.enc, dtype=np.uint8)
    starts = np.asarray(ref.ref_start, dtype="<u8")
    lens = np.asarray(ref.ref_len, dtype="<u4")
    alloc = max(reference_alloc_bytes(ref.ref_len, skipped_lens), int(enc.size))
    nm = (C.c_char_p * max(n, 1))(*[x.encode() for x in names]) if names else None
    rc = lib.ngmlr_b200_ngm_write_reference(str(path).encode(), enc.ctypes.data, int(enc.size), int(alloc), n,
                                            starts.ctypes.data, lens.ctypes.data, nm)
    if rc:
        raise OSError(f"ngmlr_b200_ngm_write_reference({path}) = {rc}")


def c_read_encoded_reference(path):
    C, lib = _c()
    n, used, alloc = C.c_uint32(), C.c_uint64(), C.c_uint64()
    args = (str(path).encode(), C.byref(n), C.byref(used), C.byref(alloc))
    rc = lib.ngmlr_b200_ngm_read_reference(*args, None, None, None, None)
    if rc:
        raise ValueError(f"ngmlr_b200_ngm_read_reference({path}) = {rc}")
    starts = np.zeros(n.value, "<u8")
    lens = np.zeros(n.value, "<u4")
    names = np.zeros((n.value, 101), np.uint8)
    enc = np.zeros(used.value, np.uint8)
    rc = lib.ngmlr_b200_ngm_read_reference(*args, starts.ctypes.data, lens.ctypes.data, names.ctypes.data,
                                           enc.ctypes.data)
    if rc:
        raise ValueError(f"ngmlr_b200_ngm_read_reference({path}) = {rc}")
    ref = EncodedReference(enc, 2 * int(used.value) - 1, [int(v) for v in starts], [int(v) for v in lens])
    return ref, [bytes(r).split(b"\0")[0].decode("ascii", "replace") for r in names]
