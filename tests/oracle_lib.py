"""ctypes bindings for the TEST-ONLY oracles: oracle/liboracle.so (C restatement) and, when it
was built in this container, oracle/_ref/libngmlr_ref.so (the unmodified reference)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

DEFAULT_SCORING = (2.0, -5.0, -5.0, -5.0, -1.0, 0.15)  # match mismatch gapopen gapext(max) gapext_min decay


class AlignOut(C.Structure):
    _fields_ = [("ret", C.c_int), ("score", C.c_float), ("position_offset", C.c_int),
                ("qstart", C.c_int), ("qend", C.c_int), ("nm", C.c_int),
                ("alignment_length", C.c_int), ("cigar_op_count", C.c_int), ("sv_type", C.c_int),
                ("identity", C.c_float), ("first_ref", C.c_int), ("first_read", C.c_int),
                ("last_ref", C.c_int), ("last_read", C.c_int), ("nm_count", C.c_int)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["score_bits"] = int(np.float32(d["score"]).view(np.uint32))
        d["identity_bits"] = int(np.float32(d["identity"]).view(np.uint32))
        return d


class Scoring(C.Structure):
    _fields_ = [(k, C.c_float) for k in ("mat", "mis", "gap_open_read", "gap_open_ref", "gap_ext",
                                          "gap_ext_min", "gap_decay")]


def scoring_struct(sc=DEFAULT_SCORING):
    mat, mis, go, ge, gemin, decay = sc
    return Scoring(mat, mis, go, go, ge, gemin, decay)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a, t=C.c_int):
    return a.ctypes.data_as(C.POINTER(t))


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "liboracle.so"], check=True)


def build_ref():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "ref"], check=True)


class Oracle:
    """The C restatement (always available)."""

    def __init__(self):
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        self.lib = C.CDLL(path)
        self.lib.or_ssw_score.restype = C.c_float
        self.lib.or_ssw_score_striped.restype = C.c_float
        self.lib.or_convex_cells.restype = C.c_int64

    def single_align(self, ref, qry, offsets, lengths, ext_qstart=0, ext_qend=0,
                     scoring=DEFAULT_SCORING, rule=0):
        offsets, lengths = _i32(offsets), _i32(lengths)
        H = len(offsets)
        out = AlignOut()
        cap = 8 * (len(qry) + len(ref)) + 64
        cig = C.create_string_buffer(cap)
        md = C.create_string_buffer(cap)
        nm = np.zeros(3 * (2 * (len(qry) + 1) + len(ref)), dtype=np.int32)
        sc = scoring_struct(scoring)
        st = self.lib.or_convex_single_align(C.byref(sc), ref, qry, _p(offsets), _p(lengths), H,
                                             ext_qstart, ext_qend, rule, C.byref(out), cig, cap, md,
                                             cap, _p(nm), len(nm) // 3)
        d = out.as_dict()
        d.update(status=st, cigar=cig.value.decode(), md=md.value.decode(),
                 nm_positions=nm[:3 * d["nm_count"]].reshape(-1, 3).copy())
        return d

    def fill(self, ref, qry, offsets, lengths, scoring=DEFAULT_SCORING, rule=0):
        offsets, lengths = _i32(offsets), _i32(lengths)
        dirs = np.zeros(int(np.maximum(lengths, 0).sum()) + 1, dtype=np.uint8)
        bs, bx, by = C.c_float(), C.c_int(), C.c_int()
        sc = scoring_struct(scoring)
        self.lib.or_convex_fill(C.byref(sc), ref, len(ref), qry, len(offsets), _p(offsets), _p(lengths),
                                rule, _p(dirs, C.c_ubyte), C.byref(bs), C.byref(bx), C.byref(by))
        return dirs[:-1], bs.value, bx.value, by.value

    def cells(self, ref_len, offsets, lengths):
        offsets, lengths = _i32(offsets), _i32(lengths)
        return int(self.lib.or_convex_cells(ref_len, len(offsets), _p(offsets), _p(lengths)))

    def ssw_score(self, ref, qry, striped=False):
        f = self.lib.or_ssw_score_striped if striped else self.lib.or_ssw_score
        return float(f(ref, qry))


class Reference:
    """The unmodified reference (only where oracle/_ref was built, i.e. /root/reference exists)."""

    PATH = os.path.join(ORACLE_DIR, "_ref", "libngmlr_ref.so")

    @classmethod
    def available(cls):
        return os.path.exists(cls.PATH)

    def __init__(self, scoring=DEFAULT_SCORING):
        self.lib = C.CDLL(self.PATH)
        self.lib.ref_convex_create.restype = C.c_void_p
        self.lib.ref_convex_create.argtypes = [C.c_float] * 6
        self.lib.ref_ssw_create.restype = C.c_void_p
        self.h = C.c_void_p(self.lib.ref_convex_create(*scoring))
        self.ssw = C.c_void_p(self.lib.ref_ssw_create())

    def close(self):
        if self.h:
            self.lib.ref_convex_destroy(self.h)
            self.lib.ref_ssw_destroy(self.ssw)
            self.h = None

    def single_align(self, ref, qry, offsets, lengths, ext_qstart=0, ext_qend=0):
        offsets, lengths = _i32(offsets), _i32(lengths)
        out = AlignOut()
        cap = 8 * (len(qry) + len(ref)) + 64
        cig = C.create_string_buffer(cap)
        md = C.create_string_buffer(cap)
        nm = np.zeros(3 * (2 * (len(qry) + 1) + len(ref)), dtype=np.int32)
        threw = self.lib.ref_convex_single_align(self.h, ref, qry, _p(offsets), _p(lengths),
                                                 len(offsets), ext_qstart, ext_qend, C.byref(out),
                                                 cig, cap, md, cap, _p(nm), len(nm) // 3)
        d = out.as_dict()
        d.update(status=threw, cigar=cig.value.decode(), md=md.value.decode(),
                 nm_positions=nm[:3 * d["nm_count"]].reshape(-1, 3).copy())
        return d

    def fill(self, ref, qry, offsets, lengths, which=0):
        offsets, lengths = _i32(offsets), _i32(lengths)
        dirs = np.zeros(int(np.maximum(lengths, 0).sum()) + 1, dtype=np.uint8)
        bs, bx, by = C.c_float(), C.c_int(), C.c_int()
        self.lib.ref_convex_fill(self.h, ref, qry, _p(offsets), _p(lengths), len(offsets), which,
                                 _p(dirs, C.c_ubyte), C.byref(bs), C.byref(bx), C.byref(by))
        return dirs[:-1], bs.value, bx.value, by.value

    def ssw_score(self, ref, qry):
        r = C.c_float(0)
        self.lib.ref_ssw_single_score(self.ssw, ref, qry, C.byref(r))
        return r.value


ALIGN_KEYS = ("ret", "score_bits", "position_offset", "qstart", "qend", "nm", "alignment_length",
              "cigar_op_count", "sv_type", "identity_bits", "first_ref", "first_read", "last_ref",
              "last_read", "nm_count", "cigar", "md")


def same_alignment(a, b):
    """Bit-exact comparison of two single_align results; returns list of differing keys."""
    bad = []
    if a["ret"] < 0 and b["ret"] < 0:
        return [] if a["score_bits"] == b["score_bits"] else ["score_bits"]
    for k in ALIGN_KEYS:
        if a[k] != b[k]:
            bad.append(k)
    if not bad and not np.array_equal(a["nm_positions"], b["nm_positions"]):
        bad.append("nm_positions")
    return bad
