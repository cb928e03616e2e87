// oracle/swap_samwriter.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Link-time substitution like oracle/swap_aligners.cpp, for SURVEY 8(f)4: the UNMODIFIED reference is linked with
// THIS file instead of src/SAMWriter.cpp. It defines the member functions of the reference's own `SAMWriter` class
// (declaration from src/SAMWriter.h) on top of libngmlr_b200.so's ngmlr_b200_sam_header / ngmlr_b200_sam_format:
// GenericReadWriter::WriteRead (inline, src/GenericReadWriter.h:78-108) keeps calling DoWriteRead once per
// non-skipped alignment; the first such call of a read formats ALL its records through the library, the later ones
// find them written. oracle/_ref/ngmlr_sam (plain aligners + this writer) must write the file the plain binary
// writes (tests/test_sam_text.py); oracle/_ref/ngmlr_b200 carries both substitutions.
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "SAMWriter.h"
#include "SequenceProvider.h"
#include "Version.h"
#include "../include/ngmlr_b200.h"

namespace {

typedef size_t (*pfHeader)(int, const char* const*, const uint64_t*, const char*, const char*,
                           const ngmlr_b200_sam_options*, const char* const*, char*, size_t);
typedef int (*pfFormat)(const ngmlr_b200_sam_options*, int64_t, const ngmlr_b200_sam_read*, const ngmlr_b200_sam_aln*,
                        int, const char* const*, const int32_t*, char*, size_t, size_t*);

void* lib() {
  static void* handle = 0;
  if (!handle) {
    const char* path = getenv("NGMLR_B200_LIB");
    handle = dlopen(path ? path : "libngmlr_b200.so", RTLD_NOW);
    if (!handle) {
      fprintf(stderr, "ngmlr_b200: cannot load the library for the SAM text: %s\n", dlerror());
      abort();
    }
  }
  return handle;
}

// contig names as SequenceProvider lists them (every contig twice: ids count both strands)
struct Names {
  std::vector<const char*> ptr;
  std::vector<int32_t> len;
  Names() {
    const int n = SequenceProvider.GetRefCount();
    for (int i = 0; i < n; ++i) {
      int l = 0;
      ptr.push_back(SequenceProvider.GetRefName(i, l));
      len.push_back(l);
    }
  }
};
const Names& names() {
  static Names n;
  return n;
}

ngmlr_b200_sam_options options(const char* rg) {
  ngmlr_b200_sam_options o;
  o.write_unmapped = Config.getWriteUnampped();
  o.bam_cigar_fix = Config.getBamCigarFix();
  o.fix_quality_orientation = 0;  // as coded
  o.threads = 1;                  // one read at a time here: ngmlr's own worker threads are the parallelism
  o.rg_id = rg;
  return o;
}

}  // namespace

void SAMWriter::DoWriteProlog() {
  std::vector<std::string> nm;
  std::vector<const char*> np;
  std::vector<uint64_t> ln;
  for (int i = 0; i < SequenceProvider.GetRefCount(); i += 2) {  // (src/SAMWriter.cpp:30-35)
    int l = 0;
    const char* p = SequenceProvider.GetRefName(i, l);
    nm.push_back(std::string(p, (size_t)l));
    ln.push_back(SequenceProvider.GetRefLen(i));
  }
  for (size_t i = 0; i < nm.size(); ++i) np.push_back(nm[i].c_str());
  const std::string version = std::string(VERSION_MAJOR) + "." + VERSION_MINOR + "." + VERSION_BUILD;
  const char* rg[11] = {Config.getRgSm(), Config.getRgLb(), Config.getRgPl(), Config.getRgDs(), Config.getRgDt(),
                        Config.getRgPu(), Config.getRgPi(), Config.getRgPg(), Config.getRgCn(), Config.getRgFo(),
                        Config.getRgKs()};
  const ngmlr_b200_sam_options o = options(rgId);
  pfHeader header = (pfHeader)dlsym(lib(), "ngmlr_b200_sam_header");
  const size_t n = header((int)np.size(), np.data(), ln.data(), version.c_str(), Config.getFullCommandLineCall(), &o, rg,
                          writeBuffer + bufferPosition, (size_t)(BUFFER_SIZE - bufferPosition));
  bufferPosition += (int)n;
  m_Writer->Flush(bufferPosition, BUFFER_LIMIT, writeBuffer, true);
}

// all records of `read` (mapped != 0) or its unmapped record
static void write_read(MappedRead const* read, bool mapped, const char* rgId, char* writeBuffer, int& bufferPosition,
                       int buffer_size) {
  std::vector<ngmlr_b200_sam_aln> as;
  const int n_aln = read->Calculated > 0 ? read->Calculated : 0;
  for (int i = 0; i < n_aln; ++i) {
    Align const& a = read->Alignments[i];
    LocationScore const& s = read->Scores[i];
    ngmlr_b200_sam_aln x;
    x.ref_pos = s.Location.m_Location;
    x.ref_id = s.Location.getrefId();
    x.reverse = s.Location.isReverse();
    x.score = s.Score.f;
    x.mq = a.MQ;
    x.nm = a.NM;
    x.identity = a.Identity;
    x.qstart = a.QStart;
    x.qend = a.QEnd;
    x.sv_type = a.svType;
    x.primary = a.primary;
    x.skip = a.skip;
    x.cigar_ops = a.cigarOpCount;
    x.cigar = a.pBuffer1 ? a.pBuffer1 : "";
    x.md = a.pBuffer2 ? a.pBuffer2 : "";
    as.push_back(x);
  }
  ngmlr_b200_sam_read r;
  r.name = read->name;
  r.seq = read->Seq;
  r.qual = read->qlty;
  r.length = read->length;
  r.n_aln = n_aln;
  r.first_aln = 0;
  r.mapped = mapped;
  r.empty = read->HasFlag(NGMNames::Empty);
  const ngmlr_b200_sam_options o = options(rgId);
  const Names& nm = names();
  size_t n = 0;
  pfFormat format = (pfFormat)dlsym(lib(), "ngmlr_b200_sam_format");
  const int rc = format(&o, 1, &r, as.data(), (int)nm.ptr.size(), nm.ptr.data(), nm.len.data(),
                        writeBuffer + bufferPosition, (size_t)(buffer_size - bufferPosition), &n);
  if (rc != 0) {
    fprintf(stderr, "ngmlr_b200_sam_format failed (%d, %zu bytes needed)\n", rc, n);
    abort();
  }
  bufferPosition += (int)n;
}

void SAMWriter::DoWriteRead(MappedRead const* const read, int const scoreID) {
  for (int j = 0; j < scoreID; ++j)
    if (!read->Alignments[j].skip) {  // an earlier record of this read: everything is already written
      NGM.AddWrittenRead(read->ReadId);
      return;
    }
  NGM.AddWrittenRead(read->ReadId);
  write_read(read, true, rgId, writeBuffer, bufferPosition, BUFFER_SIZE);
  m_Writer->Flush(bufferPosition, BUFFER_LIMIT, writeBuffer);
}

void SAMWriter::DoWriteUnmappedRead(MappedRead const* const read, int flags) {
  NGM.AddUnmappedRead(read, 0);
  if (writeUnmapped) NGM.AddWrittenRead(read->ReadId);
  write_read(read, false, rgId, writeBuffer, bufferPosition, BUFFER_SIZE);
  m_Writer->Flush(bufferPosition, BUFFER_LIMIT, writeBuffer);
}

void SAMWriter::DoWriteReadGeneric(MappedRead const* const, int const, char const*, int const, int const, int const, int) {
  fprintf(stderr, "ngmlr_b200 SAM writer: DoWriteReadGeneric is not reached through WriteRead\n");
  abort();
}

void SAMWriter::DoWriteUnmappedReadGeneric(MappedRead const* const, int const, char const, int const, int const, int const,
                                           int const, int) {
  fprintf(stderr, "ngmlr_b200 SAM writer: DoWriteUnmappedReadGeneric is not reached through WriteRead\n");
  abort();
}

void SAMWriter::DoWritePair(MappedRead const* const, int const, MappedRead const* const, int const) {
  fprintf(stderr, "ngmlr_b200 SAM writer: paired-end records are not built (ngmlr never writes them)\n");
  abort();
}

void SAMWriter::DoWriteEpilog() {}
