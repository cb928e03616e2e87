/* include/ngmlr_b200_ialignment.h -- C++ declarations that are ABI-compatible with ngmlr's
 * plugin boundary, for building libngmlr_b200.so WITHOUT the ngmlr source tree.
 *
 * When the plugin is compiled inside ngmlr, include ngmlr's own "IAlignment.h" first and define
 * NGMLR_B200_USE_HOST_IALIGNMENT: this header then adds nothing. Otherwise it declares records
 * with the same member order, types and virtual-function order as src/IAlignment.h, so that an
 * object created here can be used through ngmlr's `IAlignment*`:
 *
 *   PositionNM    src/IAlignment.h:16-27      CorridorLine  src/IAlignment.h:29-33
 *   Anchor        src/IAlignment.h:37-52      Interval      src/IAlignment.h:54-108
 *   Align         src/IAlignment.h:112-191    IAlignment    src/IAlignment.h:211-247
 *   pfCreateAlignment / pfDeleteAlignment     src/IAlignment.h:249-250
 *
 * Only layout and vtable order matter here; behaviour lives in ngmlr_b200/csrc/plugin.cpp.
 * tests/test_abi_layout.py checks sizeof/offsetof against the reference header when present.
 */
#ifndef NGMLR_B200_IALIGNMENT_H
#define NGMLR_B200_IALIGNMENT_H

#ifndef NGMLR_B200_USE_HOST_IALIGNMENT

#include <stdlib.h>

typedef long long loc;

struct PositionNM {
  int refPosition = 0;
  int readPosition = 0;
  int nm = 0;
};

struct CorridorLine {
  int offset;
  int length;
  unsigned long offsetInMatrix;
};

struct Anchor {
  int onRead;
  loc onRef;
  float score;
  bool isReverse;
  int type;
  bool isUnique;
};

struct Interval {
  Anchor* anchors = 0;
  int anchorLength = 0;
  int onReadStart = 0;
  int onReadStop = 0;
  loc onRefStart = 0;
  loc onRefStop = 0;
  double m = 0.0;
  double b = 0.0;
  double r = 0.0;
  float score = 0.0f;
  short id = 0;
  bool isReverse = false;
  bool isProcessed = false;
  bool isAssigned = false;

  Interval() {}
  int lengthOnRead() const { return onReadStop - onReadStart; }
  loc lengthOnRef() const { return llabs(onRefStop - onRefStart); }
  virtual ~Interval() {
    if (anchors != 0) {
      delete[] anchors;
      anchors = 0;
      anchorLength = 0;
    }
  }

 private:
  Interval(const Interval&);
};

struct Align {
  Align() {}
  virtual ~Align() {}

  char* pBuffer1 = 0;  /* CIGAR text */
  char* pBuffer2 = 0;  /* MD text */
  PositionNM* nmPerPosition = 0;
  Interval* mappedInterval = 0;
  PositionNM firstPosition;
  PositionNM lastPosition;
  int nmPerPostionLength = 0;
  int alignmentLength = 0;
  int PositionOffset = 0;
  int QStart = 0;
  int QEnd = 0;
  float Score = 0.0f;
  float Identity = 0.0f;
  int NM = 0;
  int MQ = 0;
  int cigarOpCount = 0;
  int maxBufferLength = 20000;
  int maxMdBufferLength = 20000;
  bool skip = false;
  bool primary = false;
  int svType = 0;
};

class IAlignment {
 public:
  virtual int GetScoreBatchSize() const = 0;
  virtual int GetAlignBatchSize() const = 0;
  virtual int BatchScore(int const mode, int const batchSize, char const* const* const refSeqList,
                         char const* const* const qrySeqList, float* const results,
                         void* extData) = 0;
  virtual int SingleAlign(int const mode, int const corridor, char const* const refSeq,
                          char const* const qrySeq, Align& result, void* extData) {
    return 0;
  }
  virtual int SingleAlign(int const mode, CorridorLine* corridor, int const corridorHeight,
                          char const* const refSeq, char const* const qrySeq, Align& result,
                          int const externalQStart, int const externalQEnd, void* extData) {
    return 0;
  }
  virtual int SingleScore(int const mode, int const corridor, char const* const refSeq,
                          char const* const qrySeq, float& result, void* extData) {
    return 0;
  }
  virtual int BatchAlign(int const mode, int const batchSize, char const* const* const refSeqList,
                         char const* const* const qrySeqList, Align* const results,
                         void* extData) = 0;
  virtual ~IAlignment() {}
};

typedef IAlignment* (*pfCreateAlignment)(int const gpu_id);
typedef void (*pfDeleteAlignment)(IAlignment*);

#endif /* NGMLR_B200_USE_HOST_IALIGNMENT */

/* extData of B200Alignment::BatchAlign: the reference leaves BatchAlign unimplemented for the
 * convex aligner (src/ConvexAlignFast.cpp:441-450), so the per-problem corridor arguments of
 * SingleAlign travel here. One entry per problem. */
struct NgmlrB200BatchAlignArgs {
  CorridorLine* corridor;
  int corridorHeight;
  int externalQStart;
  int externalQEnd;
};

extern "C" {
IAlignment* CreateAlignment(int const gpu_id);
void DeleteAlignment(IAlignment* aligner);
/* Scoring used by objects created afterwards (ngmlr passes its Config values, see
 * INTEGRATION.md); defaults are ngmlr's CLI defaults. */
void SetAlignmentScoring(float match, float mismatch, float gapOpen, float gapExtend,
                         float gapExtendMin, float gapDecay);
}

#endif
