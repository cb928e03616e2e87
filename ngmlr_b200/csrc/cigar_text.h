// ngmlr_b200/csrc/cigar_text.h -- host-side conversion of a binary CIGAR into the text/statistics
// fields of the reference's `Align` record.
//
// Replaces Convex::ConvexAlignFast::convertCigar (+ addPosition, NumberOfSetBits)
// (src/ConvexAlignFast.cpp:112-333, 76-99, 21-27) and the N-clip probe of SingleAlign (:494-529).
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace nb {

struct AlignText {
  int ret = -1;  // QStart + sum(M) + sum(I) + QEnd, the value SingleAlign returns
  int qstart = 0, qend = 0, nm = 0, alignment_length = 0, cigar_op_count = 0, sv_type = 0;
  float identity = 0.0f;
  int first_ref = 0, first_read = 0, last_ref = 0, last_read = 0;
  std::string cigar, md;
  std::vector<int32_t> nm_positions;  // triples {refPosition, readPosition, nm}
};

// runs: [leading clip][run ...][trailing clip], each len << 4 | op (ops: S 4, I 1, D 2, EQ 7, X 8).
// ref points at refSeq (whole window), ref_position = FwdResults::ref_position.
// Returns false where the reference would `throw 1` (invalid op).
bool binary_cigar_to_text(const int32_t* runs, int n_runs, const char* ref, int ref_len,
                          int ref_position, int ext_qstart, int ext_qend, AlignText& out);

// The peak scan of AlignmentBuffer::detectMisalignment (src/AlignmentBuffer.cpp:1319-1388) over the
// recorded columns: columns with 0 < (32 - nm) / 32 < 0.75 open / extend a low-identity region, the
// 21st column in a row without one closes it. The reference's loop runs to alignment_length although
// only nm_count entries were written; the entries in between are taken as zero (nm = 0: no peak).
// Appends {startInv, stopInv, startInvRead, stopInvRead} of the first `cap` closed regions to
// `regions` (cleared first) and returns how many regions were closed.
int scan_low_identity_regions(const int32_t* nm_positions, int nm_count, int alignment_length,
                              std::vector<int32_t>& regions, int cap);

}  // namespace nb
