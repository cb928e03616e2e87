// oracle/record_aligners.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The recording IAlignment decorator of SURVEY.md section 7 step 0 / BASELINE.md section 3.2: builds
// oracle/_ref/ngmlr_rec = the UNMODIFIED reference (main.cpp and every other src/*.cpp object, compiled
// where they lie) with THIS file owning the names Convex::ConvexAlignFast and StrippedSW. Every member
// function forwards to the reference's own implementation (the same two sources compiled under other
// class names, record_factory.cpp) and appends the call -- inputs and outputs -- to the file named by
// NGMLR_RECORD_FILE. The result is the real SingleAlign / BatchScore / SingleScore stream of plain ngmlr
// on a FASTQ (retries x5, realignments, full matrices, short-read paths ...), which
// scripts/replay_workload.py replays through the CUDA library and checks bit for bit.
//
// Record layout (little endian, int32 unless noted):
//   SingleAlign  : 1, refLen, qryLen, corridorHeight, extQStart, extQEnd, ret, threw, scoreBits, NM,
//                  PositionOffset, QStart, QEnd, cigarLen, mdLen, ref bytes, qry bytes,
//                  corridor offsets[height], lengths[height], cigar bytes, md bytes
//   BatchScore   : 2, n, then per pair: refLen, qryLen, scoreBits, ref bytes, qry bytes
//   SingleScore  : 3, refLen, qryLen, ret, scoreBits, ref bytes, qry bytes
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <vector>

#include "ConvexAlignFast.h"
#include "IConfig.h"
#include "StrippedSW.h"

extern "C" IAlignment* rec_make_convex(int, float, float, float, float, float, float);
extern "C" IAlignment* rec_make_ssw();

namespace {

std::mutex g_mutex;
std::map<const void*, IAlignment*> g_impl;
FILE* g_out = 0;

FILE* out() {
  if (!g_out) {
    const char* p = getenv("NGMLR_RECORD_FILE");
    g_out = fopen(p ? p : "/tmp/ngmlr_calls.bin", "wb");
    if (!g_out) abort();
  }
  return g_out;
}

struct Flush {
  ~Flush() {
    if (g_out) fclose(g_out);
  }
} g_flush;

IAlignment* impl_of(const void* self, bool convex) {
  std::lock_guard<std::mutex> lock(g_mutex);
  std::map<const void*, IAlignment*>::iterator it = g_impl.find(self);
  if (it != g_impl.end()) return it->second;
  IAlignment* a = convex ? 0 : rec_make_ssw();  // StrippedSW's inline constructor cannot register itself
  g_impl[self] = a;
  return a;
}

void drop(const void* self) {
  std::lock_guard<std::mutex> lock(g_mutex);
  std::map<const void*, IAlignment*>::iterator it = g_impl.find(self);
  if (it == g_impl.end()) return;
  delete it->second;
  g_impl.erase(it);
}

inline int fbits(float f) {
  int b;
  memcpy(&b, &f, 4);
  return b;
}

void put(std::vector<char>& b, const void* p, size_t n) { b.insert(b.end(), (const char*)p, (const char*)p + n); }
void put_i(std::vector<char>& b, int v) { put(b, &v, 4); }

}  // namespace

namespace Convex {

ConvexAlignFast::ConvexAlignFast(int const stdOutMode, float const match, float const mismatch,
                                 float const gapOpen, float const gapExtend, float const gapExtendMin,
                                 float const gapDecay)
    : defaultMaxBinaryCigarLength(200000), pacbioDebug(false), stdoutPrintAlignCorridor(stdOutMode) {
  mat = match; mis = mismatch; gap_open_read = gapOpen; gap_open_ref = gapOpen; gap_ext = gapExtend;
  gap_decay = gapDecay; gap_ext_min = gapExtendMin;
  matrix = 0;
  binaryCigar = 0;
  maxBinaryCigarLength = 0;
  alignmentId = 0;
  IAlignment* a = rec_make_convex(stdOutMode, match, mismatch, gapOpen, gapExtend, gapExtendMin, gapDecay);
  std::lock_guard<std::mutex> lock(g_mutex);
  g_impl[this] = a;
}

ConvexAlignFast::~ConvexAlignFast() { drop(this); }
int ConvexAlignFast::GetScoreBatchSize() const { return 0; }
int ConvexAlignFast::GetAlignBatchSize() const { return 0; }

int ConvexAlignFast::BatchScore(int const mode, int const n, char const* const* const r, char const* const* const q,
                                float* const res, void* ext) {
  return impl_of(this, true)->BatchScore(mode, n, r, q, res, ext);
}
int ConvexAlignFast::BatchAlign(int const mode, int const n, char const* const* const r, char const* const* const q,
                                Align* const res, void* ext) {
  return impl_of(this, true)->BatchAlign(mode, n, r, q, res, ext);
}
int ConvexAlignFast::SingleAlign(int const mode, int const corridor, char const* const r, char const* const q,
                                 Align& a, void* ext) {
  return impl_of(this, true)->SingleAlign(mode, corridor, r, q, a, ext);
}
int ConvexAlignFast::SingleAlign(int const mode, CorridorLine* c, int const h, char const* const r,
                                 char const* const q, Align& a, int const qs, int const qe, void* ext) {
  const int rl = (int)strlen(r), ql = (int)strlen(q);
  std::vector<int> off((size_t)h), len((size_t)h);
  for (int y = 0; y < h; ++y) {
    off[y] = c[y].offset;
    len[y] = c[y].length;
  }
  int ret = -1, threw = 0;
  try {
    ret = impl_of(this, true)->SingleAlign(mode, c, h, r, q, a, qs, qe, ext);
  } catch (...) {
    threw = 1;
  }
  const bool ok = !threw && ret >= 0;
  const int cl = ok && a.pBuffer1 ? (int)strlen(a.pBuffer1) : 0, ml = ok && a.pBuffer2 ? (int)strlen(a.pBuffer2) : 0;
  std::vector<char> b;
  b.reserve((size_t)rl + ql + 8 * (size_t)h + cl + ml + 64);
  const int hdr[15] = {1, rl, ql, h, qs, qe, ret, threw, fbits(a.Score), ok ? a.NM : 0, ok ? a.PositionOffset : 0,
                       ok ? a.QStart : 0, ok ? a.QEnd : 0, cl, ml};
  put(b, hdr, sizeof(hdr));
  put(b, r, rl);
  put(b, q, ql);
  put(b, off.data(), 4 * (size_t)h);
  put(b, len.data(), 4 * (size_t)h);
  if (cl) put(b, a.pBuffer1, cl);
  if (ml) put(b, a.pBuffer2, ml);
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    fwrite(b.data(), 1, b.size(), out());
  }
  if (threw) throw 1;
  return ret;
}

}  // namespace Convex

int StrippedSW::BatchScore(int const mode, int const n, char const* const* const r, char const* const* const q,
                           float* const res, void* ext) {
  const int rc = impl_of(this, false)->BatchScore(mode, n, r, q, res, ext);
  std::vector<char> b;
  put_i(b, 2);
  put_i(b, n);
  for (int i = 0; i < n; ++i) {
    const int rl = (int)strlen(r[i]), ql = (int)strlen(q[i]);
    put_i(b, rl);
    put_i(b, ql);
    put_i(b, fbits(res[i]));
    put(b, r[i], rl);
    put(b, q[i], ql);
  }
  std::lock_guard<std::mutex> lock(g_mutex);
  fwrite(b.data(), 1, b.size(), out());
  return rc;
}
int StrippedSW::SingleScore(int const mode, int const corridor, char const* const r, char const* const q,
                            float& res, void* ext) {
  const int rc = impl_of(this, false)->SingleScore(mode, corridor, r, q, res, ext);
  const int rl = (int)strlen(r), ql = (int)strlen(q);
  std::vector<char> b;
  const int hdr[5] = {3, rl, ql, rc, fbits(res)};
  put(b, hdr, sizeof(hdr));
  put(b, r, rl);
  put(b, q, ql);
  std::lock_guard<std::mutex> lock(g_mutex);
  fwrite(b.data(), 1, b.size(), out());
  return rc;
}
int StrippedSW::SingleAlign(int const, int const, char const* const, char const* const, Align&, void*) {
  throw "Not implemented";
}
int StrippedSW::BatchAlign(int const, int const, char const* const* const, char const* const* const, Align* const, void*) {
  throw "Not implemented";
}
