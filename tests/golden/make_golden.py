"""Generates tests/golden/*.json from the UNMODIFIED reference (oracle/_ref/libngmlr_ref.so, built
by oracle/Makefile from /root/reference). Run in the build container only:

    python tests/golden/make_golden.py

The vectors pin: (a) ConvexAlignFast::SingleAlign outputs (return value, score bits, CIGAR, MD,
NM, positions, nmPerPosition checksum, direction-matrix checksum, best cell) for seeded problems
under three scorings, (b) StrippedSW scores, (c) ScoreBuffer::topNSE candidate order / kept / MQ,
(d) DecodeRefSequenceExact windows, (e) corridor builders. Inputs are regenerated from the seeds by
tests/cases.py, so only outputs are stored (plus an input checksum to detect generator drift)."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import numpy as np  # noqa: E402

import cases  # noqa: E402
from oracle_lib import DEFAULT_SCORING, CsReference, Reference, build_ref, score_select_cases  # noqa: E402


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes() if not isinstance(a, (bytes, str)) else
                 (a if isinstance(a, bytes) else a.encode()))
    return h.hexdigest()[:16]


def problem_digest(p):
    return digest(p.ref, p.qry, p.offsets, p.lengths)


def record(ref, p):
    r = ref.single_align(p.ref, p.qry, p.offsets, p.lengths, p.ext_qstart, p.ext_qend)
    dirs, bs, bx, by = ref.fill(p.ref, p.qry, p.offsets, p.lengths, 0)
    keep = {k: r[k] for k in ("ret", "status", "score_bits", "position_offset", "qstart", "qend", "nm",
                              "alignment_length", "cigar_op_count", "sv_type", "identity_bits",
                              "first_ref", "first_read", "last_ref", "last_read", "nm_count")}
    keep["cigar"] = r["cigar"] if len(r["cigar"]) < 400 else None
    keep["cigar_sha"] = digest(r["cigar"])
    keep["md_sha"] = digest(r["md"])
    keep["nm_sha"] = digest(r["nm_positions"])
    keep["dirs_sha"] = digest(dirs)
    keep["best"] = [int(np.float32(bs).view(np.uint32)), bx, by]
    keep["input_sha"] = problem_digest(p)
    return keep


def main():
    build_ref()
    out = {}
    sets = {
        "default": (DEFAULT_SCORING, cases.random_problems(48, 1234) + cases.edge_problems()),
        "weird": (cases.WEIRD_SCORING, cases.random_problems(16, 4321)),
        "mild": (cases.MILD_SCORING, cases.random_problems(16, 777)),
    }
    for name, (sc, probs) in sets.items():
        ref = Reference(sc)
        out[name] = {"scoring": list(sc), "records": [record(ref, p) for p in probs]}
        ref.close()
    with open(os.path.join(HERE, "convex_golden.json"), "w") as f:
        json.dump(out, f, indent=0)
    ref = Reference()
    refs, qrys = cases.sw_pairs(256, 31)
    sw = {"scores": [ref.ssw_score(r, q) for r, q in zip(refs, qrys)],
          "input_sha": digest(b"".join(refs), b"".join(qrys))}
    # long inputs: gap-255 semantics, saturation guard, over-long -> -1
    long_ref = (b"ACGT" * 700)
    extra = [(long_ref, long_ref[:2000]), (long_ref, long_ref[:1000] + b"G" + long_ref[1000:2000]),
             (b"A" * 100001, b"A" * 10), (b"", b""), (b"A", b"A"), (b"ACGT", b"")]
    sw["extra"] = [ref.ssw_score(r, q) for r, q in extra]
    ref.close()
    with open(os.path.join(HERE, "sw_golden.json"), "w") as f:
        json.dump(sw, f)
    # (c) ScoreBuffer::topNSE / computeMQ: order produced by the reference's std::sort, kept count, MQ
    sel = []
    for sc in score_select_cases(77, 400):
        order, kept, mq = CsReference.score_select(sc)
        sel.append({"n": int(sc.size), "input_sha": digest(sc), "order_sha": digest(order), "kept": kept, "mq": mq,
                    "order": [int(v) for v in order] if sc.size <= 24 else None})
    with open(os.path.join(HERE, "score_select_golden.json"), "w") as f:
        json.dump(sel, f)
    # (d) DecodeRefSequenceExact windows (alignment reference windows) on the cs_cases genome
    import tempfile
    import cs_cases
    from ngmlr_b200 import refindex
    contigs = cs_cases.genome_contigs()
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "ref.fa")
        with open(path, "w") as f:
            for i, c in enumerate(contigs):
                f.write(f">c{i}\n{c.tobytes().decode()}\n")
        csref = CsReference(path)
        enc = refindex.encode_reference(contigs)
        wins = cs_cases.exact_windows(enc.ref_start, enc.ref_len)
        dec = [csref.decode_exact(st, ln) for st, ln in wins]
    with open(os.path.join(HERE, "decode_exact_golden.json"), "w") as f:
        json.dump({"windows_sha": digest(np.array(wins, dtype=np.int64)), "n": len(wins),
                   "text_sha": [digest(d) for d in dec],
                   "samples": [[wins[i][0], wins[i][1], dec[i].decode()] for i in range(0, len(wins), 37) if wins[i][1] <= 64]}, f)
    # (e) corridor builders of the caller (getCorridorLinear/Full/Endpoints/EndpointsWithAnchors)
    import ctypes as C
    lib = C.CDLL(CsReference.PATH)

    def ref_corridor(kind, c, arg, realign=0):
        an = c["anchors"] if kind == 3 else []
        n = len(an)
        a0 = (C.c_int * max(n, 1))(*[a[0] for a in an])
        a1 = (C.c_ulonglong * max(n, 1))(*[a[1] for a in an])
        a2 = (C.c_int * max(n, 1))(*[a[2] for a in an])
        off = np.zeros(c["q"], dtype=np.int32)
        ln = np.zeros(c["q"], dtype=np.int32)
        h = lib.ref_corridor(kind, c["q"], c["r"], arg, realign, n, a0, a1, a2, C.c_ulonglong(c["on_ref_start"]),
                             c["ext_qstart"], 256, c["full_len"], off.ctypes.data_as(C.c_void_p),
                             ln.ctypes.data_as(C.c_void_p))
        assert h == c["q"]
        return digest(off, ln)

    cor = []
    for c in cases.corridor_cases():
        cor.append({"linear": ref_corridor(0, c, c["corridor"]), "full": ref_corridor(1, c, c["r"]),
                    "endpoints": ref_corridor(2, c, c["corridor"], c["realign"]),
                    "anchors": ref_corridor(3, c, c["multiplier"])})
    with open(os.path.join(HERE, "corridor_golden.json"), "w") as f:
        json.dump(cor, f)
    print("wrote golden vectors:", {k: len(v["records"]) for k, v in out.items()}, len(sw["scores"]))


if __name__ == "__main__":
    main()
