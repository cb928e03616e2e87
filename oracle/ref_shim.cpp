// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (not part of the product path).
//
// Thin extern "C" driver around the UNMODIFIED reference sources, compiled where they lie
// under /root/reference by oracle/Makefile into oracle/_ref/libngmlr_ref.so. It exists so
// that (1) the C restatement in oracle/convex_oracle.c can be pinned against the real
// ConvexAlignFast / StrippedSW, (2) golden vectors under tests/golden/ can be generated
// (tests/golden/make_golden.py) and (3) bench.py --impl reference can time the reference's
// own CPU implementation.  Nothing under ngmlr_b200/ links or loads this.
//
// Reference entry points driven here:
//   Convex::ConvexAlignFast::SingleAlign          src/ConvexAlignFast.cpp:452-559
//   Convex::ConvexAlignFast::fwdFillMatrixSSESimple src/ConvexAlignFast.cpp:914-1287 (via SingleAlign,
//        and directly for the per-cell direction dump)
//   StrippedSW::BatchScore / SingleScore          src/StrippedSW.cpp:118-202
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

#define private public  // direction-matrix / fwdFill access for per-cell comparison only
#include "ConvexAlignFast.h"
#undef private
#include "IConfig.h"
#include "ILog.h"
#include "StrippedSW.h"

namespace {
class QuietLog : public ILog {
public:
  void _Message(int const, char const* const, char const* const, ...) const override {}
  void _Debug(int const, char const* const, char const* const, ...) const override {}
};
QuietLog quiet_log;
}  // namespace

IConfig* _config = new IConfig();
ILog const* _log = &quiet_log;

extern "C" {

struct RefAlignOut {
  int ret;          // return value of SingleAlign (cigar read length or -1)
  float score;
  int position_offset, qstart, qend, nm, alignment_length, cigar_op_count, sv_type;
  float identity;
  int first_ref, first_read, last_ref, last_read;
  int nm_count;     // number of PositionNM entries written (derived: counts slots touched)
};

void* ref_convex_create(float mat, float mis, float gap_open, float gap_ext, float gap_ext_min,
                        float gap_decay) {
  return new Convex::ConvexAlignFast(0, mat, mis, gap_open, gap_ext, gap_ext_min, gap_decay);
}

void ref_convex_destroy(void* h) { delete static_cast<Convex::ConvexAlignFast*>(h); }

// One SingleAlign call. cigar_out / md_out must hold cigar_cap / md_cap bytes; nm_out holds
// 3*nm_cap ints (refPosition, readPosition, nm). Returns 0, or 1 if the reference threw.
int ref_convex_single_align(void* h, char const* ref, char const* qry, int const* offsets,
                            int const* lengths, int height, int ext_qstart, int ext_qend,
                            RefAlignOut* out, char* cigar_out, int cigar_cap, char* md_out,
                            int md_cap, int* nm_out, int nm_cap) {
  Convex::ConvexAlignFast* aligner = static_cast<Convex::ConvexAlignFast*>(h);
  std::vector<CorridorLine> lines(height > 0 ? height : 1);
  for (int i = 0; i < height; ++i) {
    lines[i].offset = offsets[i];
    lines[i].length = lengths[i];
    lines[i].offsetInMatrix = 0;
  }
  int const qry_len = (int)strlen(qry);
  Align a;
  // Same allocation contract as the caller, src/AlignmentBuffer.cpp:271-278
  a.maxBufferLength = qry_len * 4 + 16;
  a.maxMdBufferLength = qry_len * 4 + 16;
  a.pBuffer1 = new char[a.maxBufferLength];
  a.pBuffer2 = new char[a.maxMdBufferLength];
  a.pBuffer1[0] = 0;
  a.pBuffer2[0] = 0;
  a.nmPerPostionLength = (qry_len + 1) * 2;
  a.nmPerPosition = new PositionNM[a.nmPerPostionLength];
  for (int i = 0; i < a.nmPerPostionLength; ++i) a.nmPerPosition[i].refPosition = -12345;
  int threw = 0;
  int ret = -1;
  try {
    ret = aligner->SingleAlign(0, lines.data(), height, ref, qry, a, ext_qstart, ext_qend, 0);
  } catch (...) {
    threw = 1;
  }
  out->ret = ret;
  out->score = a.Score;
  out->position_offset = a.PositionOffset;
  out->qstart = a.QStart;
  out->qend = a.QEnd;
  out->nm = a.NM;
  out->alignment_length = a.alignmentLength;
  out->cigar_op_count = a.cigarOpCount;
  out->sv_type = a.svType;
  out->identity = a.Identity;
  out->first_ref = a.firstPosition.refPosition;
  out->first_read = a.firstPosition.readPosition;
  out->last_ref = a.lastPosition.refPosition;
  out->last_read = a.lastPosition.readPosition;
  int n = 0;
  if (ret >= 0) {
    while (n < a.nmPerPostionLength && a.nmPerPosition[n].refPosition != -12345) ++n;
  }
  out->nm_count = n;
  if (cigar_out && cigar_cap > 0) {
    strncpy(cigar_out, ret >= 0 ? a.pBuffer1 : "", cigar_cap - 1);
    cigar_out[cigar_cap - 1] = 0;
  }
  if (md_out && md_cap > 0) {
    strncpy(md_out, ret >= 0 ? a.pBuffer2 : "", md_cap - 1);
    md_out[md_cap - 1] = 0;
  }
  if (nm_out) {
    for (int i = 0; i < n && i < nm_cap; ++i) {
      nm_out[3 * i + 0] = a.nmPerPosition[i].refPosition;
      nm_out[3 * i + 1] = a.nmPerPosition[i].readPosition;
      nm_out[3 * i + 2] = a.nmPerPosition[i].nm;
    }
  }
  a.clearBuffer();
  a.clearNmPerPosition();
  return threw;
}

// Forward fill only, dumping every direction byte (row-major in corridor layout, exactly the
// reference's directionMatrix) plus best cell. dirs must hold sum(lengths) bytes; cells the
// reference never writes are reported as 0xFF. which = 0: fwdFillMatrixSSESimple (the active
// fill), 1: fwdFillMatrix (scalar alternative, src/ConvexAlignFast.cpp:606-774).
int ref_convex_fill(void* h, char const* ref, char const* qry, int const* offsets,
                    int const* lengths, int height, int which, unsigned char* dirs,
                    float* best_score, int* best_ref, int* best_read) {
  Convex::ConvexAlignFast* aligner = static_cast<Convex::ConvexAlignFast*>(h);
  std::vector<CorridorLine> lines(height > 0 ? height : 1);
  size_t total = 0;
  for (int i = 0; i < height; ++i) {
    lines[i].offset = offsets[i];
    lines[i].length = lengths[i];
    total += lengths[i];
  }
  int const ref_len = (int)strlen(ref), qry_len = (int)strlen(qry);
  if (!aligner->matrix->prepare(ref_len, qry_len, lines.data(), height)) return 1;
  memset(aligner->matrix->directionMatrix, 0xFF, total);
  Convex::ConvexAlignFast::FwdResults fwd;
  memset(&fwd, 0, sizeof(fwd));
  float s = which == 0 ? aligner->fwdFillMatrixSSESimple(ref, qry, fwd, 0)
                       : aligner->fwdFillMatrix(ref, qry, fwd, 0);
  memcpy(dirs, aligner->matrix->directionMatrix, total);
  *best_score = s;
  *best_ref = fwd.best_ref_index;
  *best_read = fwd.best_read_index;
  aligner->matrix->clean();
  return 0;
}

void* ref_ssw_create() { return new StrippedSW(); }
void ref_ssw_destroy(void* h) { delete static_cast<StrippedSW*>(h); }

int ref_ssw_batch_score(void* h, int n, char const* const* refs, char const* const* qrys,
                        float* results) {
  return static_cast<StrippedSW*>(h)->BatchScore(0, n, refs, qrys, results, 0);
}

int ref_ssw_single_score(void* h, char const* ref, char const* qry, float* result) {
  return static_cast<StrippedSW*>(h)->SingleScore(0, 0, ref, qry, *result, 0);
}

}  // extern "C"
