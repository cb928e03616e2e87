"""Whole-pipeline drop-in test (BASELINE.json configs[0] in spirit: tiny reference + reads, PacBio
preset, -t 1): the UNMODIFIED ngmlr binary vs the same objects linked with the CUDA plugin in place
of ConvexAlignFast.cpp / StrippedSW.cpp (oracle/swap_aligners.cpp, built by oracle/Makefile here;
the binaries travel to the GPU box as build artefacts). Every SAM record -- flag, position, MAPQ,
CIGAR, AS/NM/MD/XI/QS/QE/SA... -- must be identical, which exercises SingleAlign retries with wider
corridors, realignment around SVs, SingleScore inversion checks and BatchScore through ngmlr's own
ScoreBuffer / AlignmentBuffer code."""
import os
import subprocess

import pytest

import e2e_data

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLAIN = os.path.join(ROOT, "oracle", "_ref", "ngmlr")
SWAPPED = os.path.join(ROOT, "oracle", "_ref", "ngmlr_b200")


@pytest.mark.skipif(not (os.path.exists(PLAIN) and os.path.exists(SWAPPED)),
                    reason="oracle/_ref/ngmlr{,_b200} not built (needs /root/reference at build time)")
def test_sam_identical_with_plugin(tmp_path):
    ref, fq = e2e_data.write_dataset(str(tmp_path))
    outs = {}
    for name, exe in (("cpu", PLAIN), ("b200", SWAPPED)):
        sam = str(tmp_path / f"{name}.sam")
        env = dict(os.environ, NGMLR_B200_LIB=os.path.join(ROOT, "ngmlr_b200", "libngmlr_b200.so"))
        r = subprocess.run([exe, "-r", ref, "-q", fq, "-o", sam, "-t", "1", "--skip-write", "--no-progress"],
                           capture_output=True, text=True, env=env, timeout=900)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[name] = e2e_data.sam_records(sam)
    assert len(outs["cpu"]) >= 24
    assert len(outs["cpu"]) == len(outs["b200"])
    for a, b in zip(outs["cpu"], outs["b200"]):
        assert a == b, f"SAM record differs:\nCPU : {a[:300]}\nB200: {b[:300]}"


@pytest.mark.skipif(not (os.path.exists(PLAIN) and os.path.exists(SWAPPED)),
                    reason="oracle/_ref/ngmlr{,_b200} not built (needs /root/reference at build time)")
def test_sam_identical_with_cross_thread_batcher(tmp_path):
    """8 ngmlr worker threads, every blocking SingleAlign parked in the plugin's cross-thread batcher
    (NGMLR_B200_BATCH_WINDOW_US): the same records as the single-threaded CPU run (record order in the
    file depends on thread timing, so compare sorted)."""
    ref, fq = e2e_data.write_dataset(str(tmp_path))
    outs = {}
    for name, exe, threads, extra in (("cpu", PLAIN, "1", {}),
                                      ("b200", SWAPPED, "8", {"NGMLR_B200_BATCH_WINDOW_US": "300"})):
        sam = str(tmp_path / f"{name}.sam")
        env = dict(os.environ, NGMLR_B200_LIB=os.path.join(ROOT, "ngmlr_b200", "libngmlr_b200.so"), **extra)
        r = subprocess.run([exe, "-r", ref, "-q", fq, "-o", sam, "-t", threads, "--skip-write", "--no-progress"],
                           capture_output=True, text=True, env=env, timeout=900)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[name] = sorted(e2e_data.sam_records(sam))
    assert len(outs["cpu"]) >= 24 and outs["cpu"] == outs["b200"]
