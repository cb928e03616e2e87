"""The real call stream of plain ngmlr -- every IAlignment::SingleAlign / BatchScore / SingleScore call the
unmodified reference makes on a FASTQ, recorded by the decorators of oracle/record_aligners.cpp
(oracle/_ref/ngmlr_rec) -- replayed call by call:
  * CPU: through the oracle (pins the C restatement on retries x5, realignments, full matrices,
    corridors several thousand columns wide -- inputs the seeded generators only approximate);
  * GPU (-m gpu): through the CUDA library (batched convex path with the device text stage, sub-read
    scorer), bit for bit against what the reference returned."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REC = os.path.join(ROOT, "oracle", "_ref", "ngmlr_rec")
sys.path.insert(0, os.path.join(ROOT, "scripts"))

needs_rec = pytest.mark.skipif(not os.path.exists(REC), reason="oracle/_ref/ngmlr_rec not built")


@pytest.fixture(scope="module")
def stream(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("rec"))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "record_workload.py"), "--reads", "28",
                        "--genome-mb", "1", "--contigs", "2", "--threads", "8", "--median", "5000", "--sv",
                        "--seed", "7", "--out", d], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    from replay_workload import parse
    aligns, pairs = parse(os.path.join(d, "calls.bin"))
    side = json.load(open(os.path.join(d, "calls.json")))
    assert len(aligns) >= 40 and len(pairs) >= 1000
    return aligns, pairs, side


def _same(a, r):
    if a["threw"] or a["ret"] < 0:
        return r["ret"] < 0
    return (r["ret"] == a["ret"] and r["score_bits"] == a["score_bits"] and r["cigar"] == a["cigar"]
            and r["md"] == a["md"] and r["nm"] == a["nm"] and r["position_offset"] == a["po"]
            and r["qstart"] == a["qstart"] and r["qend"] == a["qend"])


@needs_rec
def test_oracle_equals_the_reference_on_its_real_call_stream(stream, oracle):
    aligns, pairs, _side = stream
    bad = [i for i, a in enumerate(aligns)
           if not _same(a, oracle.single_align(a["ref"], a["qry"], a["off"], a["len"], a["qs"], a["qe"]))]
    assert not bad, bad[:5]
    W = np.array([int(a["len"][0]) for a in aligns])
    # the stream really contains what the seeded generators lack: wide corridors and repeated / failed calls
    assert W.max() >= 1500 and len(aligns) > 28
    for r, q, sb in pairs[:1500]:
        assert int(np.float32(oracle.ssw_score(r, q)).view(np.uint32)) == sb


@needs_rec
@pytest.mark.gpu
def test_cuda_library_equals_the_reference_on_its_real_call_stream(stream):
    from ngmlr_b200 import B200Aligner, PackedBatch
    aligns, pairs, _side = stream
    al = B200Aligner(0)
    try:
        for on_device in (True, False):
            al.set_text_stage(on_device, False)
            batch = PackedBatch([a["ref"] for a in aligns], [a["qry"] for a in aligns], [a["off"] for a in aligns],
                                [a["len"] for a in aligns], [a["qs"] for a in aligns], [a["qe"] for a in aligns])
            res = al.BatchAlign(batch)
            bad = [i for i, (a, g) in enumerate(zip(aligns, res)) if not _same(a, g.as_dict())]
            assert not bad, (on_device, bad[:5])
        got = al.BatchScore([p[0] for p in pairs], [p[1] for p in pairs])
        want = np.array([p[2] for p in pairs], dtype=np.uint32)
        assert np.array_equal(got.view(np.uint32), want)
    finally:
        al.close()
