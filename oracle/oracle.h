/* oracle/oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, scalar) of the algorithms on ngmlr's alignment hot path. Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * liboracle.so; the product (ngmlr_b200/) never does and fails loudly without its CUDA library.
 *
 * Parity status: PINNED. Every function here is checked against the unmodified reference
 * compiled from /root/reference (oracle/_ref/libngmlr_ref.so, built by oracle/Makefile) in
 * tests/test_oracle.py and tests/test_cs_oracle.py (the whole-reference library oracle/_ref/libngmlr_full.so for the
 * candidate search, window decoding and candidate selection), and against the committed golden vectors that library
 * produced (tests/golden/, generator tests/golden/make_golden.py).
 */
#ifndef NGMLR_B200_ORACLE_H
#define NGMLR_B200_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Direction / CIGAR op codes, reference src/AlignmentMatrixFast.h:15-24 */
enum { OR_I = 1, OR_D = 2, OR_S = 4, OR_EQ = 7, OR_X = 8, OR_STOP = 10 };

/* Convex scoring, reference src/ConvexAlignFast.cpp:29-43 (ctor) */
typedef struct {
  float mat, mis, gap_open_read, gap_open_ref, gap_ext, gap_ext_min, gap_decay;
} or_scoring;

/* Mirrors the fields of the reference's `Align` that SingleAlign fills (src/IAlignment.h:112-191);
 * same layout as RefAlignOut in oracle/ref_shim.cpp. */
typedef struct {
  int ret; /* SingleAlign return: read length covered by the CIGAR incl. clips, or -1 */
  float score;
  int position_offset, qstart, qend, nm, alignment_length, cigar_op_count, sv_type;
  float identity;
  int first_ref, first_read, last_ref, last_read;
  int nm_count;
} or_align_out;

/* Forward fill. rule: 0 = as coded in fwdFillMatrixSSESimple (SSE blocks + scalar left fix-up +
 * scalar tail, src/ConvexAlignFast.cpp:914-1287); 1 = the pure scalar rule of fwdFillMatrix
 * (:606-774). dirs (may be NULL) receives the reference's directionMatrix layout: row y at
 * sum(lengths[0..y)), column x - offsets[y]; never-written cells keep 0xFF. */
int or_convex_fill(const or_scoring* sc, const char* ref, int ref_len, const char* qry, int height,
                   const int* offsets, const int* lengths, int rule, unsigned char* dirs,
                   float* best_score, int* best_ref, int* best_read);

/* Whole SingleAlign (fill + revBacktrack + convertCigar + N-clip probe),
 * src/ConvexAlignFast.cpp:452-559. Returns 0, 1 if the reference would have thrown, 2 if
 * prepare() refuses the matrix (>= max_matrix_mb). nm_out: 3 ints per entry. */
int or_convex_single_align(const or_scoring* sc, const char* ref, const char* qry, const int* offsets,
                           const int* lengths, int height, int ext_qstart, int ext_qend, int rule,
                           or_align_out* out, char* cigar_out, int cigar_cap, char* md_out,
                           int md_cap, int* nm_out, int nm_cap);

/* Number of DP cells SingleAlign evaluates: sum over rows of
 * min(off+len, ref_len) - max(0, off), clamped at 0. */
int64_t or_convex_cells(int ref_len, int height, const int* offsets, const int* lengths);

/* StrippedSW score: src/StrippedSW.cpp:118-202 over ssw.c:366-538 (sw_sse2_word, score only).
 * Returns the float the reference writes (uint16 best score, or -1.0f for over-long input). */
float or_ssw_score(const char* ref, const char* qry);
/* Same, emulating the 8-lane striped evaluation order of sw_sse2_word literally. */
float or_ssw_score_striped(const char* ref, const char* qry);
int or_ssw_batch_score(int n, const char* const* refs, const char* const* qrys, float* results);

/* ---- candidate search (stage 0), reference layout, k-mer index: oracle/cs_oracle.c ---------- */
struct or_cs;
struct or_cs* or_cs_create(int ncontigs, const char* const* contigs, const int64_t* lens);
void or_cs_destroy(struct or_cs* h);
uint64_t or_cs_concat_len(const struct or_cs* h);
int or_cs_ref_count(const struct or_cs* h);
uint64_t or_cs_ref_start(const struct or_cs* h, int i);
const uint8_t* or_cs_encoded(const struct or_cs* h, uint64_t* bytes);
int or_cs_decode(const struct or_cs* h, uint64_t position, uint64_t buffer_len, char* out);
int or_cs_decode_exact(const struct or_cs* h, uint64_t start, uint64_t sequence_len, int corridor, char* out);
int or_cs_build_index(struct or_cs* h, int k, int skip, int bin_shift, int max_freq);
uint32_t or_cs_index_len(const struct or_cs* h);
uint32_t or_cs_npos(const struct or_cs* h);
const uint32_t* or_cs_tab(const struct or_cs* h);
const int8_t* or_cs_rci(const struct or_cs* h);
const uint32_t* or_cs_pos(const struct or_cs* h);
int or_cs_search(const struct or_cs* h, const char* seq, int len, float sensitivity, float min_kmer_hits,
                 float* scores, uint64_t* locs, int* reverse, int cap, float* max_hits);

/* ---- candidate selection after scoring (ScoreBuffer::topNSE / computeMQ): oracle/score_oracle.c - */
int or_score_mq(float best, float second);
int or_score_select(const float* scores, int n, int32_t* order, int* mq);

#ifdef __cplusplus
}
#endif
#endif
