// oracle/swap_aligners.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Link-time substitution used to build oracle/_ref/ngmlr_b200: the UNMODIFIED reference (main.cpp
// and every other src/*.cpp object, compiled where they lie) is linked with THIS file instead of
// src/ConvexAlignFast.cpp and src/StrippedSW.cpp. It defines the member functions of the
// reference's own `Convex::ConvexAlignFast` and `StrippedSW` classes (declarations from the
// reference headers) as forwarders to the IAlignment object libngmlr_b200.so's CreateAlignment()
// returns. Every construction site of the reference (src/AlignmentBuffer.h:355-368,
// src/NGM.cpp:350-362, src/AlignmentBuffer.cpp:1217, src/ScoreBuffer.cpp:217) therefore gets the
// CUDA implementation without touching a line of ngmlr -- the integration of INTEGRATION.md, done
// by the linker. tests/test_gpu_ngmlr_e2e.py compares the SAM this binary writes with the SAM of
// the plain reference binary (oracle/_ref/ngmlr).
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>

#include <map>
#include <mutex>

#include "ConvexAlignFast.h"
#include "IConfig.h"
#include "StrippedSW.h"

namespace {

typedef void (*pfScoring)(float, float, float, float, float, float);

void* plugin() {
  static void* handle = 0;
  if (!handle) {
    const char* path = getenv("NGMLR_B200_LIB");
    handle = dlopen(path ? path : "libngmlr_b200.so", RTLD_NOW);
    if (!handle) {
      fprintf(stderr, "ngmlr_b200: cannot load plugin: %s\n", dlerror());
      abort();  // no silent CPU fallback
    }
  }
  return handle;
}

IAlignment* create_impl() {
  IAlignment* a = ((pfCreateAlignment)dlsym(plugin(), "CreateAlignment"))(0);
  if (!a) {
    fprintf(stderr, "ngmlr_b200: CreateAlignment failed (no CUDA device?)\n");
    abort();
  }
  return a;
}

std::mutex g_mutex;
std::map<const void*, IAlignment*> g_impl;  // reference object -> plugin object

IAlignment* impl_of(const void* self) {
  std::lock_guard<std::mutex> lock(g_mutex);
  std::map<const void*, IAlignment*>::iterator it = g_impl.find(self);
  if (it != g_impl.end()) return it->second;
  IAlignment* a = create_impl();  // StrippedSW's inline constructor cannot register itself
  g_impl[self] = a;
  return a;
}

void drop(const void* self) {
  std::lock_guard<std::mutex> lock(g_mutex);
  std::map<const void*, IAlignment*>::iterator it = g_impl.find(self);
  if (it == g_impl.end()) return;
  ((pfDeleteAlignment)dlsym(plugin(), "DeleteAlignment"))(it->second);
  g_impl.erase(it);
}

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
  ((pfScoring)dlsym(plugin(), "SetAlignmentScoring"))(match, mismatch, gapOpen, gapExtend, gapExtendMin, gapDecay);
  std::lock_guard<std::mutex> lock(g_mutex);
  g_impl[this] = create_impl();
}

ConvexAlignFast::~ConvexAlignFast() { drop(this); }
int ConvexAlignFast::GetScoreBatchSize() const { return 0; }
int ConvexAlignFast::GetAlignBatchSize() const { return 0; }

int ConvexAlignFast::BatchScore(int const mode, int const n, char const* const* const r, char const* const* const q,
                                float* const res, void* ext) {
  return impl_of(this)->BatchScore(mode, n, r, q, res, ext);
}
int ConvexAlignFast::BatchAlign(int const mode, int const n, char const* const* const r, char const* const* const q,
                                Align* const res, void* ext) {
  return impl_of(this)->BatchAlign(mode, n, r, q, res, ext);
}
int ConvexAlignFast::SingleAlign(int const mode, int const corridor, char const* const r, char const* const q,
                                 Align& a, void* ext) {
  return impl_of(this)->SingleAlign(mode, corridor, r, q, a, ext);
}
int ConvexAlignFast::SingleAlign(int const mode, CorridorLine* c, int const h, char const* const r,
                                 char const* const q, Align& a, int const qs, int const qe, void* ext) {
  return impl_of(this)->SingleAlign(mode, c, h, r, q, a, qs, qe, ext);
}

}  // namespace Convex

int StrippedSW::BatchScore(int const mode, int const n, char const* const* const r, char const* const* const q,
                           float* const res, void* ext) {
  return impl_of(this)->BatchScore(mode, n, r, q, res, ext);
}
int StrippedSW::SingleScore(int const mode, int const corridor, char const* const r, char const* const q,
                            float& res, void* ext) {
  return impl_of(this)->SingleScore(mode, corridor, r, q, res, ext);
}
int StrippedSW::SingleAlign(int const, int const, char const* const, char const* const, Align&, void*) {
  throw "Not implemented";
}
int StrippedSW::BatchAlign(int const, int const, char const* const* const, char const* const* const, Align* const, void*) {
  throw "Not implemented";
}
