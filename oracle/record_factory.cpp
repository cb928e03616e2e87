// oracle/record_factory.cpp -- TEST INFRASTRUCTURE ONLY (see record_aligners.cpp).
// Compiled with -DConvexAlignFast=RefConvexAlignFast -DStrippedSW=RefStrippedSW, like the two reference
// sources it is linked with (src/ConvexAlignFast.cpp, src/StrippedSW.cpp, compiled where they lie): the
// reference's own classes under another name, so that the recording decorators can own the real names.
#include "ConvexAlignFast.h"
#include "StrippedSW.h"

extern "C" IAlignment* rec_make_convex(int stdOutMode, float match, float mismatch, float gapOpen, float gapExtend,
                                       float gapExtendMin, float gapDecay) {
  return new Convex::ConvexAlignFast(stdOutMode, match, mismatch, gapOpen, gapExtend, gapExtendMin, gapDecay);
}
extern "C" IAlignment* rec_make_ssw() { return new StrippedSW(); }
