"""Writes tests/golden/sam_golden.json: the output of the UNMODIFIED reference writer (SAMWriter through
oracle/_ref/libngmlr_full.so, shim ref_sam_write in oracle/ref_cs_shim.cpp) for the seeded cases of
tests/sam_cases.py -- per configuration the sha256 of the whole text, a 16-hex digest per line and the header.
Run in this container (needs /root/reference compiled by `make -C oracle ref`):  python tests/golden/make_sam_golden.py
Also usable as a module: reference_text(config) returns the live text (each call in a fresh process is the
caller's business: the reference keeps singletons)."""
import ctypes as C
import hashlib
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

import sam_cases  # noqa: E402

FULL = os.path.join(ROOT, "oracle", "_ref", "libngmlr_full.so")
RG_FIELDS = [b"smp", None, b"PACBIO", b"a description", None, None, None, b"ngmlr", b"centre", None, None]
CMDLINE = b"ngmlr -r ref.fa -q reads.fq -t 4"


class RefSamAln(C.Structure):
    _fields_ = [("ref_pos", C.c_ulonglong), ("ref_id", C.c_int), ("reverse", C.c_int), ("score", C.c_float),
                ("mq", C.c_int), ("nm", C.c_int), ("identity", C.c_float), ("qstart", C.c_int), ("qend", C.c_int),
                ("sv_type", C.c_int), ("primary", C.c_int), ("skip", C.c_int), ("cigar_ops", C.c_int),
                ("cigar", C.c_char_p), ("md", C.c_char_p)]


class RefSamRead(C.Structure):
    _fields_ = [("name", C.c_char_p), ("seq", C.c_char_p), ("qual", C.c_char_p), ("length", C.c_int),
                ("n_aln", C.c_int), ("first_aln", C.c_longlong), ("mapped", C.c_int), ("empty", C.c_int)]


class Reference:
    def __init__(self):
        self.lib = C.CDLL(FULL)
        self.tmp = tempfile.mkdtemp(prefix="samgold_")
        fa = os.path.join(self.tmp, "ref.fa")
        open(fa, "wb").write(sam_cases.fasta_bytes())
        self.lib.ref_cs_init(fa.encode())
        self.lib.ref_sam_write.restype = C.c_longlong
        self.lib.ref_cs_ref_len.restype = C.c_ulonglong
        n = self.lib.ref_cs_ref_count()
        self.names, self.lens = [], []
        for i in range(0, n, 2):        # every contig is listed twice (forward / reverse slot), src/SAMWriter.cpp:30-35
            buf = C.create_string_buffer(1024)
            k = self.lib.ref_cs_ref_name(i, buf, 1024)
            self.names.append(buf.raw[:k])
            self.lens.append(int(self.lib.ref_cs_ref_len(i)))

    def _call(self, what, reads, opts, cmdline=CMDLINE, rg_fields=None):
        n_aln = sum(len(r.alignments) for r in reads)
        rr = (RefSamRead * max(len(reads), 1))()
        aa = (RefSamAln * max(n_aln, 1))()
        k = 0
        for i, r in enumerate(reads):
            rr[i] = RefSamRead(r.name, r.seq, r.qual, len(r.seq), len(r.alignments), k, int(r.mapped), int(r.empty))
            for a in r.alignments:
                # the reference's ref ids count forward / reverse slots: contig j is id 2 * j
                aa[k] = RefSamAln(a.ref_pos, 2 * a.ref_id, int(a.reverse), a.score, a.mq, a.nm, a.identity, a.qstart,
                                  a.qend, a.sv_type, int(a.primary), int(a.skip), a.cigar_ops, a.cigar, a.md)
                k += 1
        fields = (C.c_char_p * 11)(*rg_fields) if rg_fields else None
        args = (what, rr, len(reads), aa, int(opts["bam_cigar_fix"]), int(opts["write_unmapped"]), opts["rg_id"],
                fields, cmdline)
        need = self.lib.ref_sam_write(*args, None, 0)
        buf = C.create_string_buffer(need + 1)
        got = self.lib.ref_sam_write(*args, buf, need)
        assert got == need
        return buf.raw[:need]

    def records(self, reads, opts):
        return self._call(1, reads, opts)

    def header(self, opts, rg_fields=None):
        return self._call(0, [], opts, rg_fields=rg_fields)


def line_digests(text):
    return [hashlib.sha256(l).hexdigest()[:16] for l in text.split(b"\n")[:-1]]


def reference_blob():
    ref = Reference()
    blob = {"ref_names": [n.decode() for n in ref.names], "ref_lens": ref.lens, "configs": {}}
    for name, opts in sam_cases.CONFIGS:
        reads = sam_cases.config_reads(name)
        text = ref.records(reads, opts)
        header = ref.header(opts, RG_FIELDS if opts["rg_id"] else None)
        blob["configs"][name] = {"sha256": hashlib.sha256(text).hexdigest(), "bytes": len(text),
                                 "lines": line_digests(text), "header": header.decode(),
                                 "first_lines": [l.decode("latin1") for l in text.split(b"\n")[:3]]}
    return blob


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--dump":   # live text of one configuration on stdout (tests)
        ref = Reference()
        opts = dict(sam_cases.CONFIGS)[sys.argv[2]]
        sys.stdout.buffer.write(ref.records(sam_cases.config_reads(sys.argv[2]), opts))
        sys.exit(0)
    blob = reference_blob()
    with open(os.path.join(HERE, "sam_golden.json"), "w") as f:
        json.dump(blob, f, indent=0)
    print({k: (v["bytes"], len(v["lines"])) for k, v in blob["configs"].items()})
