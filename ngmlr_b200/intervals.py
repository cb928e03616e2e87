"""Batched mirror of AlignmentBuffer::computeAlignment (src/AlignmentBuffer.cpp:226-465): what ngmlr
does around one SingleAlign call -- extract the reference window of the interval, pick the corridor
builder for the attempt, align, and retry with a wider corridor while the alignment does not cover
the read -- restated for a whole batch of intervals: attempt k of every interval that is still
invalid forms ONE device batch.

    while not valid and corridor * multiplier <= 2 * refSeqLen and attempts left (5; 1 if full):
        lines = full   -> getCorridorFull(refSeqLen)
                short  -> getCorridorLinear(corridor * multiplier)
                multiplier < 3 and not realign and anchors -> getCorridorEndpointsWithAnchors(multiplier)
                else   -> getCorridorEndpoints(corridor * multiplier, realign)
        cigarLength = SingleAlign(...);  valid = cigarLength == fullReadLength;  multiplier += 1
    a thrown SingleAlign ends the interval (computeAlignment's catch(...) returns 0)

The backend supplies the two device operations: `decode(starts, seq_lens) -> [bytes]`
(DecodeRefSequenceExact) and `align(tasks, refs, offsets, lengths) -> [Align-like]` (SingleAlign).
`B200Backend` runs them on the GPU through the C ABI; the CPU tests plug in the oracle.
"""
from dataclasses import dataclass, field

import numpy as np

from . import corridor as _cor


@dataclass
class IntervalTask:
    on_ref_start: int
    on_ref_stop: int
    read_seq: bytes
    corridor: int
    ext_qstart: int = 0
    ext_qend: int = 0
    full_read_length: int = 0          # 0: len(read_seq) + clips
    anchors: list = field(default_factory=list)   # [(onRead, onRef, isReverse)]
    realign: bool = False
    full_alignment: bool = False
    short_read: bool = False
    read_part_length: int = 256
    # the read part named by index into the resident read set (B200Aligner.reads_upload) instead of as text
    read_index: int = -1
    on_read_start: int = 0
    read_seq_len: int = 0
    reverse: bool = False

    def __post_init__(self):
        if self.read_seq is not None and not self.read_seq_len:
            self.read_seq_len = len(self.read_seq)
        if not self.full_read_length:
            self.full_read_length = self.read_seq_len + self.ext_qstart + self.ext_qend


def _corridor_for(task, ref_len, multiplier):
    q = len(task.read_seq)
    if task.full_alignment:
        return _cor.corridor_full(q, task.ref_seq_len)   # getCorridorFull(refSeqLen): the buffer length, NUL included
    if task.short_read:
        return _cor.corridor_linear(q, task.corridor_eff * multiplier)
    if multiplier < 3 and not task.realign and task.anchors:
        ax = [a[1] - task.on_ref_start for a in task.anchors]
        ay = [task.full_read_length - a[0] - task.read_part_length - task.ext_qstart if a[2]
              else a[0] - task.ext_qstart for a in task.anchors]
        return _cor.corridor_endpoints_with_anchors(q, ref_len, ax, ay, multiplier)
    return _cor.corridor_endpoints(q, ref_len, task.corridor_eff * multiplier, realign=task.realign)


def compute_alignments(backend, tasks):
    """-> one result per task: the backend's Align-like object of the first valid attempt, or None
    (computeAlignment returned 0). Also returns the number of SingleAlign calls made, per task."""
    n = len(tasks)
    results = [None] * n
    calls = [0] * n
    live = []
    # extractReferenceSequenceForAlignment: onRefStart >= onRefStop -> no sequence -> 0
    cand = [i for i, t in enumerate(tasks) if t.read_seq is not None and t.on_ref_start < t.on_ref_stop]
    refs = {}
    if cand:
        texts = backend.decode([tasks[i].on_ref_start for i in cand],
                               [tasks[i].on_ref_stop - tasks[i].on_ref_start + 1 for i in cand])
        for i, tx in zip(cand, texts):
            if tx is None:
                continue
            refs[i] = tx
            t = tasks[i]
            t.ref_seq_len = t.on_ref_stop - t.on_ref_start + 1          # refSeqLength incl. the NUL
            t.corridor_eff = min(t.corridor, t.ref_seq_len * 2)           # (:266-267)
            t.retry = 1 if t.full_alignment else 5
            t.mult = 1
            live.append(i)
    while live:
        batch = []
        for i in live:
            t = tasks[i]
            if t.corridor_eff * t.mult <= t.ref_seq_len * 2 and t.retry > 0:
                t.retry -= 1
                batch.append(i)
        if not batch:
            break
        offs, lens = [], []
        for i in batch:
            o, l = _corridor_for(tasks[i], len(refs[i]), tasks[i].mult)
            offs.append(o)
            lens.append(l)
        out = backend.align([tasks[i] for i in batch], [refs[i] for i in batch], offs, lens)
        nxt = []
        for i, r in zip(batch, out):
            calls[i] += 1
            t = tasks[i]
            if getattr(r, "threw", False) or (isinstance(r, dict) and r.get("status")):
                continue                                 # catch(...) -> 0
            ret = r["ret"] if isinstance(r, dict) else r.ret
            if ret == t.full_read_length:
                results[i] = r
            else:
                t.mult += 1
                nxt.append(i)
        live = nxt
    return results, calls


class B200Backend:
    """The two device operations through a B200Aligner whose reference is set (set_reference)."""

    def __init__(self, aligner):
        self.al = aligner

    def decode(self, starts, seq_lens):
        return self.al.decode_windows(starts, seq_lens)

    def align(self, tasks, refs, offsets, lengths):
        from .aligner import PackedBatch
        batch = PackedBatch(refs, [t.read_seq for t in tasks], offsets, lengths,
                            [t.ext_qstart for t in tasks], [t.ext_qend for t in tasks])
        self.al.upload_windows(batch, [t.on_ref_start for t in tasks], [t.on_ref_stop for t in tasks])
        self.al.run()
        return list(self.al.fetch())
