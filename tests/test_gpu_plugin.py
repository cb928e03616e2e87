"""GPU test of the C++ plugin boundary: a driver built against include/ngmlr_b200_ialignment.h
dlopens libngmlr_b200.so and calls it through the IAlignment vtable exactly as ngmlr would
(CreateAlignment, SingleAlign with CorridorLine[], BatchAlign, BatchScore, SingleScore)."""
import os
import re
import subprocess

import numpy as np
import pytest

import cases
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_driver(tmp_path):
    exe = str(tmp_path / "plugin_driver")
    subprocess.run(["g++", "-O1", "-std=c++11", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "plugin_driver.cpp"), "-o", exe, "-ldl"], check=True)
    return exe


def test_plugin_through_ialignment_vtable(tmp_path):
    probs = cases.random_problems(10, 515, max_len=700, modes=(0, 1, 3))
    path = tmp_path / "problems.txt"
    with open(path, "w") as f:
        f.write(f"{len(probs)}\n")
        for p in probs:
            f.write(f"{len(p.ref)} {len(p.qry)} {p.ext_qstart} {p.ext_qend}\n{p.ref.decode()}\n{p.qry.decode()}\n")
            f.write(" ".join(f"{o} {l}" for o, l in zip(p.offsets, p.lengths)) + "\n")
    exe = _build_driver(tmp_path)
    out = subprocess.run([exe, os.path.join(ROOT, "ngmlr_b200", "libngmlr_b200.so"), str(path)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.splitlines()
    assert lines[0] == "batchsize 1024 1024"
    orc = Oracle()
    kv = lambda l: dict(x.split("=", 1) for x in l.split()[2:])
    singles = {int(l.split()[1]): kv(l) for l in lines if l.startswith("single ")}
    batches = {int(l.split()[1]): kv(l) for l in lines if l.startswith("batch ")}
    assert all("ok" in l for l in lines if l.startswith("offsetInMatrix"))
    assert "batchalign rc=%d" % len(probs) in lines
    for i, p in enumerate(probs):
        want = orc.single_align(p.ref, p.qry, p.offsets, p.lengths, p.ext_qstart, p.ext_qend)
        got = singles[i]
        assert int(got["ret"]) == want["ret"], i
        assert int(got["score_bits"]) == want["score_bits"], i
        if want["ret"] >= 0:
            assert got["cigar"] == want["cigar"] and got["md"] == want["md"]
            assert int(got["identity_bits"]) == want["identity_bits"]
            for k, w in (("pos", "position_offset"), ("qstart", "qstart"), ("qend", "qend"), ("nm", "nm"),
                         ("alen", "alignment_length"), ("ops", "cigar_op_count"), ("sv", "sv_type")):
                assert int(got[k]) == want[w], (i, k)
            assert got["first"] == f"{want['first_ref']},{want['first_read']}"
            assert got["last"] == f"{want['last_ref']},{want['last_read']}"
            b = batches[i]
            for k in ("score_bits", "cigar", "md", "pos", "nm", "qstart", "qend"):
                assert b[k] == got[k], (i, k)
        else:
            assert int(batches[i]["ret"]) == -1
    scores = [l.split() for l in lines if l.startswith("score ")]
    assert "batchscore rc=%d" % len(probs) in lines
    for i, p in enumerate(probs):
        w = orc.ssw_score(p.ref[:306], p.qry[:256])
        assert float(scores[i][2]) == w and float(scores[i][3]) == w and int(scores[i][4]) == 1
    assert "single_int_corridor_throws 1" in lines


def test_oversized_matrices_are_refused_like_prepare(tmp_path):
    """AlignmentMatrixFast::prepare refuses a matrix of >= --max-matrix-size MB and SingleAlign then returns -1 with
    Score -1 (src/AlignmentMatrixFast.cpp:45-58, src/ConvexAlignFast.cpp:466-468). With the limit at 1 MB
    (NGMLR_B200_MAX_MATRIX_MB, the plugin's Config.getMaxMatrixSizeMB()): full matrices of >= 10^6 cells fail -- alone
    and inside a batch, without disturbing their neighbours --, the smaller problems equal the oracle."""
    from ngmlr_b200 import corridor
    probs = cases.random_problems(6, 616, min_len=300, max_len=700, modes=(0, 3))
    for i in (1, 4):                           # two full matrices: 1300+ x 1300+ cells
        big = cases.random_problems(1, 700 + i, min_len=1300, max_len=1400, modes=(0,))[0]
        big.offsets, big.lengths = corridor.corridor_full(len(big.qry), len(big.ref))
        assert int(np.sum(big.lengths)) >= 1_000_000
        probs[i] = big
    path = tmp_path / "problems.txt"
    with open(path, "w") as f:
        f.write(f"{len(probs)}\n")
        for p in probs:
            f.write(f"{len(p.ref)} {len(p.qry)} {p.ext_qstart} {p.ext_qend}\n{p.ref.decode()}\n{p.qry.decode()}\n")
            f.write(" ".join(f"{o} {l}" for o, l in zip(p.offsets, p.lengths)) + "\n")
    exe = _build_driver(tmp_path)
    env = dict(os.environ, NGMLR_B200_MAX_MATRIX_MB="1")
    out = subprocess.run([exe, os.path.join(ROOT, "ngmlr_b200", "libngmlr_b200.so"), str(path)],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.splitlines()
    kv = lambda l: dict(x.split("=", 1) for x in l.split()[2:])
    singles = {int(l.split()[1]): kv(l) for l in lines if l.startswith("single ")}
    batches = {int(l.split()[1]): kv(l) for l in lines if l.startswith("batch ")}
    assert all("ok" in l for l in lines if l.startswith("offsetInMatrix"))   # prepare() still publishes the offsets
    orc = Oracle()
    minus_one = int(np.float32(-1.0).view(np.uint32))
    for i, p in enumerate(probs):
        if i in (1, 4):
            assert int(singles[i]["ret"]) == -1 and int(singles[i]["score_bits"]) == minus_one
            assert int(batches[i]["ret"]) == -1 and int(batches[i]["score_bits"]) == minus_one
            continue
        want = orc.single_align(p.ref, p.qry, p.offsets, p.lengths, p.ext_qstart, p.ext_qend)
        for got in (singles[i], batches[i]):
            assert int(got["score_bits"]) == want["score_bits"], i
            if want["ret"] >= 0:
                assert got["cigar"] == want["cigar"] and got["md"] == want["md"]
