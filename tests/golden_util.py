import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes() if not isinstance(a, (bytes, str)) else
                 (a if isinstance(a, bytes) else a.encode()))
    return h.hexdigest()[:16]


def load(name):
    with open(os.path.join(HERE, "golden", name)) as f:
        return json.load(f)


def golden_sets():
    """[(set name, scoring tuple, problems, records)] with inputs regenerated from the seeds."""
    import cases
    g = load("convex_golden.json")
    probs = {
        "default": cases.random_problems(48, 1234) + cases.edge_problems(),
        "weird": cases.random_problems(16, 4321),
        "mild": cases.random_problems(16, 777),
    }
    out = []
    for name, blob in g.items():
        ps = probs[name]
        assert len(ps) == len(blob["records"])
        for p, r in zip(ps, blob["records"]):
            assert digest(p.ref, p.qry, p.offsets, p.lengths) == r["input_sha"], "generator drifted"
        out.append((name, tuple(blob["scoring"]), ps, blob["records"]))
    return out


SCALAR_KEYS = ("ret", "score_bits", "position_offset", "qstart", "qend", "nm", "alignment_length",
               "cigar_op_count", "sv_type", "identity_bits", "first_ref", "first_read", "last_ref",
               "last_read", "nm_count")


def check_against_record(res, rec, what=""):
    """res: dict in oracle_lib.single_align format. Failed alignments compare ret/score only."""
    assert res["status"] == rec["status"], f"{what}: threw mismatch"
    if rec["ret"] < 0:
        assert res["ret"] < 0, f"{what}: expected failure"
        return
    for k in SCALAR_KEYS:
        assert res[k] == rec[k], f"{what}: {k}: {res[k]} != {rec[k]}"
    assert digest(res["cigar"]) == rec["cigar_sha"], f"{what}: CIGAR differs"
    assert digest(res["md"]) == rec["md_sha"], f"{what}: MD differs"
    assert digest(np.asarray(res["nm_positions"], dtype=np.int32)) == rec["nm_sha"], f"{what}: nmPerPosition differs"
