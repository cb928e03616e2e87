// ngmlr_b200/csrc/sam_text.cpp -- SAM records of a batch of reads, formatted by host threads.
//
// Replaces SAMWriter::DoWriteProlog (src/SAMWriter.cpp:22-85), DoWriteRead -> DoWriteReadGeneric (:87-224),
// DoWriteUnmappedRead -> DoWriteUnmappedReadGeneric (:301-363) and the loop of GenericReadWriter::WriteRead
// (src/GenericReadWriter.h:78-108). The reference appends every field with vsprintf into one 100 MB buffer
// under the output mutex' eventual flush; here a record is emitted by one routine into a "sink" that either
// counts or copies, so a batch is sized (pass 1, parallel over reads), given offsets (one prefix sum) and
// written in place (pass 2, parallel) -- records come out in read order whatever the thread count.
//
// As-coded details that are kept (the tests compare bytes with the unmodified writer):
//   * positions: `m_Location + 1` is a 64-bit value printed with %u (POS) and %d (SA:Z) -> low 32 bits;
//   * AS / XE are (int) of the float score, XS is the constant 0, XI is round(identity * 1e4) / 1e4 with %g,
//     CV is (length - clipped) * 100.0f / length with %f;
//   * a reverse-strand record reverses the read's quality string IN PLACE (:104-108), so a later reverse
//     record of the same read finds it forward again (opts->fix_quality_orientation = 1 turns that off);
//   * --bam-fix: >= 0x10000 CIGAR operations -> "<length>S" in the CIGAR column and the operations as
//     CG:B:I,(len << 4 | op) printed with %d (:123-130, :199-218);
//   * the unmapped record ends with the quality column (no trailing tab) and ORs 0x4 into the flags.
// Deliberate difference: a "*" quality (FASTA input) is never reversed; the reference reverses `length`
// bytes of a 2-byte buffer there (undefined behaviour, SURVEY 8(f)4 "FASTA reverse-strand overflow").
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/ngmlr_b200.h"

namespace nb {
int host_threads();  // capi.cu
}

namespace {

struct CountSink {
  size_t n = 0;
  void put(const char*, size_t len) { n += len; }
  void ch(char) { ++n; }
  // sequence columns: only the length matters
  void seq_fwd(const char*, size_t len) { n += len; }
  void seq_revcomp(const char*, size_t len) { n += len; }
  void reversed(const char*, size_t len) { n += len; }
};

struct WriteSink {
  char* p;
  void put(const char* s, size_t len) {
    memcpy(p, s, len);
    p += len;
  }
  void ch(char c) { *p++ = c; }
  void seq_fwd(const char* s, size_t len) { put(s, len); }
  // MappedRead::computeReverseSeq (src/MappedRead.cpp:37-69): A<->T, C<->G, anything else unchanged
  void seq_revcomp(const char* s, size_t len) {
    static const struct Table {
      unsigned char t[256];
      Table() {
        for (int i = 0; i < 256; ++i) t[i] = (unsigned char)i;
        t['A'] = 'T'; t['T'] = 'A'; t['C'] = 'G'; t['G'] = 'C';
      }
    } tab;
    const unsigned char* src = (const unsigned char*)s + len;
    for (size_t i = 0; i < len; ++i) p[i] = (char)tab.t[*--src];
    p += len;
  }
  void reversed(const char* s, size_t len) {
    std::reverse_copy(s, s + len, p);
    p += len;
  }
};

template <class Sink>
inline void put_str(Sink& o, const char* s) { o.put(s, strlen(s)); }

template <class Sink>
inline void put_lit(Sink& o, const char* s, size_t n) { o.put(s, n); }
#define LIT(o, s) put_lit(o, s, sizeof(s) - 1)

template <class Sink>
inline void put_u64(Sink& o, uint64_t v) {
  char b[24];
  int n = 0;
  do { b[n++] = (char)('0' + v % 10); v /= 10; } while (v);
  char r[24];
  for (int i = 0; i < n; ++i) r[i] = b[n - 1 - i];
  o.put(r, (size_t)n);
}

template <class Sink>
inline void put_i32(Sink& o, int32_t v) {  // %d
  if (v < 0) {
    o.ch('-');
    put_u64(o, (uint64_t)(-(int64_t)v));
  } else {
    put_u64(o, (uint64_t)v);
  }
}

template <class Sink>
inline void put_u32(Sink& o, uint32_t v) { put_u64(o, v); }  // %u

// length of the prefix of s that "%.*s" with precision n prints
inline size_t prefix_len(const char* s, size_t n) {
  const void* z = memchr(s, 0, n);
  return z ? (size_t)((const char*)z - s) : n;
}

struct Ctx {
  const ngmlr_b200_sam_options* opts;
  const ngmlr_b200_sam_aln* alns;
  const char* const* ref_names;
  const int32_t* ref_name_lens;
  int n_refs;
};

template <class Sink>
inline void put_ref_name(Sink& o, const Ctx& c, int id) {
  if (id >= 0 && id < c.n_refs) o.put(c.ref_names[id], (size_t)c.ref_name_lens[id]);
}

// DoWriteUnmappedReadGeneric(read, -1, '*', -1, -1, 0, 0, flags | 0x4)   (:301-363)
template <class Sink>
void unmapped_record(Sink& o, const Ctx& c, const ngmlr_b200_sam_read& r) {
  if (!c.opts->write_unmapped) return;
  put_str(o, r.name);
  LIT(o, "\t4\t*\t0\t0\t*\t*\t0\t0\t");
  o.seq_fwd(r.seq, prefix_len(r.seq, (size_t)r.length));
  o.ch('\t');
  if (r.qual) o.put(r.qual, prefix_len(r.qual, (size_t)r.length));
  else o.ch('*');
  if (c.opts->rg_id) {
    LIT(o, "\tRG:Z:");
    put_str(o, c.opts->rg_id);
  }
  o.ch('\n');
}

// DoWriteReadGeneric(read, i, "*", -1, 0, mappingQlty, 0)   (:92-224); qual_reversed = state of the read's
// quality string when the record is printed
template <class Sink>
void mapped_record(Sink& o, const Ctx& c, const ngmlr_b200_sam_read& r, int i, bool qual_reversed) {
  const ngmlr_b200_sam_aln* al = c.alns + r.first_aln;
  const ngmlr_b200_sam_aln& a = al[i];
  const size_t len = (size_t)r.length;
  int flags = 0;
  if (!a.primary) flags |= 0x800;
  if (a.reverse) flags |= 0x10;
  put_str(o, r.name);
  o.ch('\t');
  put_i32(o, flags);
  o.ch('\t');
  put_ref_name(o, c, a.ref_id);
  o.ch('\t');
  put_u32(o, (uint32_t)(a.ref_pos + 1));
  o.ch('\t');
  put_i32(o, a.mq);
  o.ch('\t');
  const bool long_cigar = c.opts->bam_cigar_fix && !a.skip && a.cigar_ops >= 0x10000;
  if (long_cigar) {
    put_i32(o, r.length);
    o.ch('S');
  } else {
    put_str(o, a.cigar);
  }
  LIT(o, "\t*\t0\t0\t");  // mate name "*", mate position -1 + 1 as %u, template length 0
  if (a.reverse) o.seq_revcomp(r.seq, len);  // RevSeq is zero-filled behind `length`: no embedded NUL possible
  else o.seq_fwd(r.seq, prefix_len(r.seq, len));
  o.ch('\t');
  if (r.qual) {
    const size_t ql = prefix_len(r.qual, len);
    if (qual_reversed && ql == len) o.reversed(r.qual, len);
    else o.put(r.qual, ql);
    o.ch('\t');
  } else {
    LIT(o, "*\t");
  }
  if (c.opts->rg_id) {
    LIT(o, "RG:Z:");
    put_str(o, c.opts->rg_id);
    o.ch('\t');
  }
  LIT(o, "AS:i:");
  put_i32(o, (int)a.score);
  LIT(o, "\tNM:i:");
  put_i32(o, a.nm);
  {
    // the reference's unqualified round() is ::round(double): double quotient, narrowed by the assignment
    const float identity = (float)(::round((double)(a.identity * 10000.0f)) / (double)10000.0f);
    char b[48];
    const int n = snprintf(b, sizeof b, "\tXI:f:%g\tXS:i:0\tXE:i:", identity);
    o.put(b, (size_t)n);
  }
  put_i32(o, (int)a.score);
  LIT(o, "\tXR:i:");
  put_i32(o, r.length - a.qstart - a.qend);
  LIT(o, "\tMD:Z:");
  put_str(o, a.md);
  o.ch('\t');
  if (a.sv_type > -1) {
    LIT(o, "SV:i:");
    put_i32(o, a.sv_type);
    o.ch('\t');
  }
  if (r.n_aln > 1) {
    bool first = true;
    for (int j = 0; j < r.n_aln; ++j) {
      if (j == i || al[j].skip) continue;
      if (first) {
        LIT(o, "SA:Z:");
        first = false;
      }
      put_ref_name(o, c, al[j].ref_id);
      o.ch(',');
      put_i32(o, (int32_t)(al[j].ref_pos + 1));
      o.ch(',');
      o.ch(al[j].reverse ? '-' : '+');
      o.ch(',');
      put_str(o, al[j].cigar);
      o.ch(',');
      put_i32(o, al[j].mq);
      o.ch(',');
      put_i32(o, al[j].nm);
      o.ch(';');
    }
    if (!first) o.ch('\t');
  }
  LIT(o, "QS:i:");
  put_i32(o, a.qstart);
  LIT(o, "\tQE:i:");
  put_i32(o, r.length - a.qend);
  {
    const int clipped = a.qstart + a.qend;
    const float covered = (r.length - clipped) * 100.0f / r.length;
    char b[64];
    const int n = snprintf(b, sizeof b, "\tCV:f:%f", covered);
    o.put(b, (size_t)n);
  }
  if (long_cigar) {
    LIT(o, "\tCG:B:I");
    char* p = const_cast<char*>(a.cigar);
    for (int k = 0; k < a.cigar_ops; ++k) {
      const long l = strtol(p, &p, 10);
      int op = 0;
      switch (*p) {
        case 'M': op = 0; break;
        case 'I': op = 1; break;
        case 'D': op = 2; break;
        case 'N': op = 3; break;
        case 'S': op = 4; break;
        case 'H': op = 5; break;
        case '=': op = 7; break;
        case 'X': op = 8; break;
        default: op = 0; break;
      }
      ++p;
      o.ch(',');
      put_i32(o, (int32_t)((unsigned int)l << 4 | (unsigned int)op));
    }
  }
  o.ch('\n');
}

// GenericReadWriter::WriteRead(read, mapped)
template <class Sink>
void read_records(Sink& o, const Ctx& c, const ngmlr_b200_sam_read& r) {
  bool mapped_once = false;
  bool qual_reversed = false;
  if (r.mapped) {
    for (int i = 0; i < r.n_aln; ++i) {
      const ngmlr_b200_sam_aln& a = c.alns[r.first_aln + i];
      if (a.skip) continue;
      mapped_once = true;
      if (a.reverse) qual_reversed = c.opts->fix_quality_orientation ? true : !qual_reversed;
      else if (c.opts->fix_quality_orientation) qual_reversed = false;
      mapped_record(o, c, r, i, qual_reversed);
    }
  }
  if (!mapped_once && !r.empty) unmapped_record(o, c, r);
}

template <typename F>
void run_parallel(int64_t n, int threads, F fn) {
  const int64_t chunk = 16;
  const int64_t pieces = (n + chunk - 1) / chunk;
  threads = (int)std::min<int64_t>(threads, pieces);
  std::atomic<int64_t> next(0);
  auto work = [&]() {
    for (;;) {
      const int64_t b = next.fetch_add(chunk);
      if (b >= n) break;
      const int64_t e = std::min(n, b + chunk);
      for (int64_t i = b; i < e; ++i) fn(i);
    }
  };
  if (threads <= 1) {
    work();
    return;
  }
  std::vector<std::thread> pool;
  pool.reserve((size_t)threads - 1);
  for (int t = 1; t < threads; ++t) pool.emplace_back(work);
  work();
  for (auto& t : pool) t.join();
}

}  // namespace

extern "C" {

size_t ngmlr_b200_sam_header(int n_refs, const char* const* ref_names, const uint64_t* ref_lens,
                             const char* version, const char* command_line, const ngmlr_b200_sam_options* opts,
                             const char* const* rg_fields, char* out, size_t cap) {
  auto emit = [&](auto& o) {
    LIT(o, "@HD\tVN:1.0\tSO:unsorted\n");
    for (int i = 0; i < n_refs; ++i) {
      LIT(o, "@SQ\tSN:");
      put_str(o, ref_names[i]);
      LIT(o, "\tLN:");
      put_u64(o, ref_lens[i]);
      o.ch('\n');
    }
    LIT(o, "@PG\tID:ngmlr\tPN:nextgenmap-lr\tVN:");
    put_str(o, version ? version : "");
    LIT(o, "\tCL:");
    put_str(o, command_line ? command_line : "(null)");  // vsprintf("%s", 0) prints "(null)"
    o.ch('\n');
    if (opts && opts->rg_id) {
      static const char* const keys[11] = {"SM", "LB", "PL", "DS", "DT", "PU", "PI", "PG", "CN", "FO", "KS"};
      LIT(o, "@RG\tID:");
      put_str(o, opts->rg_id);
      for (int k = 0; k < 11; ++k) {
        if (!rg_fields || !rg_fields[k]) continue;
        o.ch('\t');
        o.put(keys[k], 2);
        o.ch(':');
        put_str(o, rg_fields[k]);
      }
      o.ch('\n');
    }
  };
  CountSink cs;
  emit(cs);
  if (out && cs.n <= cap) {
    WriteSink ws{out};
    emit(ws);
  }
  return cs.n;
}

int ngmlr_b200_sam_format(const ngmlr_b200_sam_options* opts, int64_t n_reads, const ngmlr_b200_sam_read* reads,
                          const ngmlr_b200_sam_aln* alns, int n_refs, const char* const* ref_names,
                          const int32_t* ref_name_lens, char* out, size_t cap, size_t* written) {
  if (!opts || n_reads < 0 || (n_reads && !reads) || !written) return -1;
  for (int64_t i = 0; i < n_reads; ++i) {
    const ngmlr_b200_sam_read& r = reads[i];
    if (!r.name || !r.seq || r.length < 0 || (r.n_aln > 0 && !alns)) return -1;
  }
  const Ctx c{opts, alns, ref_names, ref_name_lens, n_refs};
  const int threads = opts->threads > 0 ? opts->threads : nb::host_threads();
  std::vector<size_t> off((size_t)n_reads + 1, 0);
  run_parallel(n_reads, threads, [&](int64_t i) {
    CountSink cs;
    read_records(cs, c, reads[i]);
    off[(size_t)i + 1] = cs.n;
  });
  for (int64_t i = 0; i < n_reads; ++i) off[(size_t)i + 1] += off[(size_t)i];
  *written = off[(size_t)n_reads];
  if (*written > cap || (!out && *written)) return -2;
  run_parallel(n_reads, threads, [&](int64_t i) {
    WriteSink ws{out + off[(size_t)i]};
    read_records(ws, c, reads[i]);
  });
  return 0;
}

}  // extern "C"
