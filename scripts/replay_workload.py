#!/usr/bin/env python
"""Replay a recorded call stream of plain ngmlr (scripts/record_workload.py) through the CUDA library:
every SingleAlign call through the batched convex path (CorridorLine rows exactly as recorded, device text
stage), every BatchScore / SingleScore pair through the sub-read scorer -- outputs compared bit for bit
with what the reference returned (ret, score bits, NM, PositionOffset, QStart, QEnd, CIGAR, MD; scores).
Prints one JSON line: shape of the real workload (cells per read base, H / W percentiles, retries) and the
GPU time for it."""
import argparse
import json
import os
import struct
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def parse(path):
    buf = open(path, "rb").read()
    mv = memoryview(buf)
    at = 0
    aligns, pairs = [], []
    n = len(buf)
    while at + 4 <= n:
        kind = struct.unpack_from("<i", buf, at)[0]
        if kind == 1:
            h = struct.unpack_from("<15i", buf, at)
            at += 60
            rl, ql, hh, qs, qe, ret, threw, sbits, nm, po, qst, qen, cl, ml = h[1:]
            ref = bytes(mv[at:at + rl]); at += rl
            qry = bytes(mv[at:at + ql]); at += ql
            off = np.frombuffer(buf, dtype=np.int32, count=hh, offset=at); at += 4 * hh
            ln = np.frombuffer(buf, dtype=np.int32, count=hh, offset=at); at += 4 * hh
            cig = bytes(mv[at:at + cl]).decode(); at += cl
            md = bytes(mv[at:at + ml]).decode(); at += ml
            aligns.append(dict(ref=ref, qry=qry, off=off, len=ln, h=hh, qs=qs, qe=qe, ret=ret, threw=threw,
                               score_bits=sbits & 0xffffffff, nm=nm, po=po, qstart=qst, qend=qen, cigar=cig, md=md))
        elif kind == 2:
            cnt = struct.unpack_from("<i", buf, at + 4)[0]
            at += 8
            for _ in range(cnt):
                rl, ql, sb = struct.unpack_from("<3i", buf, at)
                at += 12
                pairs.append((bytes(mv[at:at + rl]), bytes(mv[at + rl:at + rl + ql]), sb & 0xffffffff))
                at += rl + ql
        elif kind == 3:
            _k, rl, ql, _rc, sb = struct.unpack_from("<5i", buf, at)
            at += 20
            pairs.append((bytes(mv[at:at + rl]), bytes(mv[at + rl:at + rl + ql]), sb & 0xffffffff))
            at += rl + ql
        else:
            raise SystemExit(f"bad record kind {kind} at {at}")
    return aligns, pairs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dir", default="/tmp/ngmlr_calls")
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--gpu", type=int, default=0)
    args = ap.parse_args()
    from ngmlr_b200 import B200Aligner, PackedBatch
    side = json.load(open(os.path.join(args.dir, "calls.json")))
    aligns, pairs = parse(os.path.join(args.dir, "calls.bin"))
    al = B200Aligner(args.gpu)
    al.set_text_stage(True, False)
    # ---- SingleAlign stream ----
    cells = 0
    bad = []
    gpu_ms = dict(fill=0.0, traceback=0.0, text=0.0)
    wall = 0.0
    valid_heights = []
    for b0 in range(0, len(aligns), args.batch):
        chunk = [a for a in aligns[b0:b0 + args.batch] if a["h"] == len(a["qry"])]
        skipped = len(aligns[b0:b0 + args.batch]) - len(chunk)
        assert skipped == 0, "recorded call with corridorHeight != read length"
        batch = PackedBatch([a["ref"] for a in chunk], [a["qry"] for a in chunk], [a["off"] for a in chunk],
                            [a["len"] for a in chunk], [a["qs"] for a in chunk], [a["qe"] for a in chunk])
        t0 = time.perf_counter()
        al.upload(batch)
        al.run()
        res = al.fetch()
        wall += time.perf_counter() - t0
        st = al.stats()
        cells += st["cells"]
        for k in gpu_ms:
            gpu_ms[k] += st[k + "_ms"]
        for j, a in enumerate(chunk):
            g = res[j]
            if a["threw"] or a["ret"] < 0:
                ok = (g.ret < 0 or g.threw) and (a["threw"] or np.float32(g.Score).view(np.uint32) == a["score_bits"])
            else:
                ok = (g.ret == a["ret"] and int(np.float32(g.Score).view(np.uint32)) == a["score_bits"]
                      and g.NM == a["nm"] and g.PositionOffset == a["po"] and g.QStart == a["qstart"]
                      and g.QEnd == a["qend"] and g.pBuffer1 == a["cigar"] and g.pBuffer2 == a["md"])
            if not ok:
                bad.append(b0 + j)
    # ---- BatchScore / SingleScore stream ----
    sw_bad = 0
    sw_ms = 0.0
    for b0 in range(0, len(pairs), 65536):
        ch = pairs[b0:b0 + 65536]
        t0 = time.perf_counter()
        got = al.BatchScore([p[0] for p in ch], [p[1] for p in ch])
        wall += time.perf_counter() - t0
        sw_ms += al.sw_kernel_ms()
        want = np.array([p[2] for p in ch], dtype=np.uint32)
        sw_bad += int((got.view(np.uint32) != want).sum())
    H = np.array([a["h"] for a in aligns])
    W = np.array([int(a["len"][0]) if a["h"] else 0 for a in aligns])
    mat = np.array([int(np.maximum(a["len"], 0).sum()) for a in aligns], dtype=np.int64)
    pct = lambda v, q: int(np.percentile(v, q)) if len(v) else 0
    line = {
        "what": "recorded call stream of plain ngmlr replayed through the CUDA library",
        "recipe": side.get("recipe"), "reads": side["reads"], "read_bases": side["read_bases"],
        "singlealign_calls": len(aligns), "calls_per_read": len(aligns) / max(side["reads"], 1),
        "invalid_or_failed_calls": int(sum(1 for a in aligns if a["threw"] or a["ret"] < 0)),
        "dp_cells": int(cells), "cells_per_read_base": cells / max(side["read_bases"], 1),
        "H_p50_p99_max": [pct(H, 50), pct(H, 99), int(H.max()) if len(H) else 0],
        "W_p50_p90_p99_max": [pct(W, 50), pct(W, 90), pct(W, 99), int(W.max()) if len(W) else 0],
        "largest_matrix_cells": int(mat.max()) if len(mat) else 0,
        "subread_scorings": len(pairs),
        "parity": {"singlealign_mismatches": len(bad), "first": bad[:5], "score_mismatches": sw_bad},
        "gpu_kernel_ms": dict(gpu_ms, sw_score=sw_ms),
        "gpu_wall_s_host_buffers": wall,
        "gbp_per_s_host_buffers": side["read_bases"] / wall / 1e9,
        "gcells_per_s_fill": cells / max(gpu_ms["fill"], 1e-9) / 1e6,
        "cpu_reference": {"wall_s_incl_index_build": side["wall_s_incl_index_build"], "threads": side["threads"],
                          "done_line": side["ngmlr_done_line"]},
    }
    print(json.dumps(line))
    al.close()
    if bad or sw_bad:
        raise SystemExit(f"replay mismatch: {len(bad)} alignments, {sw_bad} scores")


if __name__ == "__main__":
    main()
