/* oracle/ssw_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Scalar restatement of the sub-read scorer:
 *   StrippedSW::BatchScore / SingleScore   src/StrippedSW.cpp:118-202
 *   nt_table                               src/StrippedSW.cpp:111-116
 *   scoring matrix                         src/StrippedSW.h:20-39  (+1 / -1, anything vs code 4 = 0)
 *   ssw_init -> qP_word                    lib/.../ssw.c:964-989, 342-364
 *   ssw_align(flag=0) -> sw_sse2_word      lib/.../ssw.c:997-1054, 366-538
 *
 * Facts that matter for exactness:
 *   - both lengths passed to SSW are strlen+1: the terminating NUL is part of the sequence and
 *     maps to code 4 (scores 0 against everything);
 *   - gap_open = gap_extension = int32 -1 narrowed to uint8_t => 255;
 *   - score_size = 1 => 16-bit kernel only; H uses signed saturating adds, E/F use unsigned
 *     saturating subtracts; the result is the uint16 maximum of H, returned as float;
 *   - either length >= 100000 => -1.0f.
 */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>

#define SSW_MAX_LEN 100000
#define SSW_GAP 255

static inline int nt_code(char c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
  }
}

static inline int sub_score(int a, int b) {
  if (a == 4 || b == 4) return 0;
  return a == b ? 1 : -1;
}

static inline int sat_add16(int a, int b) { /* _mm_adds_epi16 */
  int s = a + b;
  if (s > 32767) s = 32767;
  if (s < -32768) s = -32768;
  return s;
}

static inline int sat_subu16(int a, int b) { /* _mm_subs_epu16 on values held as int16 bit patterns */
  unsigned ua = (unsigned)a & 0xFFFFu, ub = (unsigned)b & 0xFFFFu;
  unsigned r = ua > ub ? ua - ub : 0u;
  return (int)(int16_t)r;
}

static inline int max16(int a, int b) { return a > b ? a : b; } /* _mm_max_epi16 (signed) */

/* Textbook evaluation order (row-major Gotoh with SSW's conventions). */
float or_ssw_score(const char* ref, const char* qry) {
  int read_len = (int)strlen(qry) + 1, ref_len = (int)strlen(ref) + 1;
  if (read_len >= SSW_MAX_LEN || ref_len >= SSW_MAX_LEN) return -1.0f;
  int* H = (int*)calloc((size_t)read_len, sizeof(int));
  int* E = (int*)calloc((size_t)read_len, sizeof(int));
  int best = 0;
  for (int i = 0; i < ref_len; ++i) {
    int rc = nt_code(ref[i]); /* ref[ref_len-1] is the NUL -> 4 */
    int diag = 0, F = 0;
    for (int j = 0; j < read_len; ++j) {
      int h = sat_add16(diag, sub_score(rc, nt_code(qry[j])));
      h = max16(h, E[j]);
      h = max16(h, F);
      diag = H[j];
      H[j] = h;
      if (h > best) best = h;
      int hg = sat_subu16(h, SSW_GAP);
      E[j] = max16(sat_subu16(E[j], SSW_GAP), hg);
      F = max16(sat_subu16(F, SSW_GAP), hg);
    }
  }
  free(H);
  free(E);
  return (float)(uint16_t)best;
}

/* Literal emulation of the striped evaluation (8 int16 lanes, segLen = (readLen+7)/8), including
 * the lazy-F loop that corrects H but leaves E untouched (ssw.c:454-467). */
float or_ssw_score_striped(const char* ref, const char* qry) {
  int read_len = (int)strlen(qry) + 1, ref_len = (int)strlen(ref) + 1;
  if (read_len >= SSW_MAX_LEN || ref_len >= SSW_MAX_LEN) return -1.0f;
  int seg = (read_len + 7) / 8;
  size_t n = (size_t)seg * 8;
  int* prof = (int*)malloc(sizeof(int) * 5 * n); /* qP_word: prof[nt][i][lane], query index i + lane*seg */
  for (int nt = 0; nt < 5; ++nt)
    for (int i = 0; i < seg; ++i)
      for (int l = 0; l < 8; ++l) {
        int j = i + l * seg;
        prof[((size_t)nt * seg + i) * 8 + l] = j >= read_len ? 0 : sub_score(nt, nt_code(qry[j]));
      }
  int* hstore = (int*)calloc(n, sizeof(int));
  int* hload = (int*)calloc(n, sizeof(int));
  int* ev = (int*)calloc(n, sizeof(int));
  int vmax[8] = {0};
  int best = 0;
  for (int i = 0; i < ref_len; ++i) {
    int vF[8] = {0}, vH[8];
    /* vH = pvHStore[segLen-1] shifted up one lane */
    vH[0] = 0;
    for (int l = 1; l < 8; ++l) vH[l] = hstore[(size_t)(seg - 1) * 8 + l - 1];
    int* tmp = hload; hload = hstore; hstore = tmp;
    const int* vp = prof + (size_t)nt_code(ref[i]) * seg * 8;
    for (int j = 0; j < seg; ++j) {
      for (int l = 0; l < 8; ++l) {
        int h = sat_add16(vH[l], vp[(size_t)j * 8 + l]);
        int e = ev[(size_t)j * 8 + l];
        h = max16(h, e);
        h = max16(h, vF[l]);
        if (h > vmax[l]) vmax[l] = h;
        hstore[(size_t)j * 8 + l] = h;
        h = sat_subu16(h, SSW_GAP);
        e = max16(sat_subu16(e, SSW_GAP), h);
        ev[(size_t)j * 8 + l] = e;
        vF[l] = max16(sat_subu16(vF[l], SSW_GAP), h);
        vH[l] = hload[(size_t)j * 8 + l];
      }
    }
    int done = 0;
    for (int k = 0; k < 8 && !done; ++k) {
      for (int l = 7; l > 0; --l) vF[l] = vF[l - 1];
      vF[0] = 0;
      for (int j = 0; j < seg; ++j) {
        int any = 0;
        for (int l = 0; l < 8; ++l) {
          int h = max16(hstore[(size_t)j * 8 + l], vF[l]);
          hstore[(size_t)j * 8 + l] = h;
          if (h > vmax[l]) vmax[l] = h; /* see note below */
          h = sat_subu16(h, SSW_GAP);
          vF[l] = sat_subu16(vF[l], SSW_GAP);
          if (vF[l] > h) any = 1;
        }
        if (!any) { done = 1; break; }
      }
    }
    /* NOTE: the reference folds only the pre-lazy-F column maximum (vMaxColumn) into vMaxScore;
     * a lazy-F-raised H can never exceed the H it was derived from, so the global maximum is the
     * same whether or not corrected cells are counted. */
  }
  for (int l = 0; l < 8; ++l) if (vmax[l] > best) best = vmax[l];
  free(prof); free(hstore); free(hload); free(ev);
  return (float)(uint16_t)best;
}

int or_ssw_batch_score(int n, const char* const* refs, const char* const* qrys, float* results) {
  for (int i = 0; i < n; ++i) results[i] = or_ssw_score(refs[i], qrys[i]);
  return n;
}
