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


class Reference:
    def __init__(self):
        import oracle_lib
        lib = C.CDLL(FULL)
        self.tmp = tempfile.mkdtemp(prefix="samgold_")
        fa = os.path.join(self.tmp, "ref.fa")
        open(fa, "wb").write(sam_cases.fasta_bytes())
        lib.ref_cs_init(fa.encode())
        self.w = oracle_lib.SamReference(lib)
        self.names, self.lens = self.w.names, self.w.lens

    def records(self, reads, opts):
        return self.w.records(reads, cmdline=CMDLINE, **opts)

    def header(self, opts, rg_fields=None):
        return self.w.header(cmdline=CMDLINE, rg_fields=rg_fields, **opts)


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
