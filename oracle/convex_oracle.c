/* oracle/convex_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Scalar C restatement of ngmlr's convex-gap banded Smith-Waterman:
 *   fill        Convex::ConvexAlignFast::fwdFillMatrixSSESimple   src/ConvexAlignFast.cpp:914-1287
 *               (and the scalar alternative fwdFillMatrix          src/ConvexAlignFast.cpp:606-774)
 *   storage     Convex::AlignmentMatrixFast                        src/AlignmentMatrixFast.{h,cpp}
 *   traceback   ConvexAlignFast::revBacktrack                      src/ConvexAlignFast.cpp:335-432
 *   text        ConvexAlignFast::convertCigar (+addPosition)       src/ConvexAlignFast.cpp:76-333
 *   driver      ConvexAlignFast::SingleAlign                       src/ConvexAlignFast.cpp:452-559
 *
 * Built with -ffp-contract=off: the reference binary has no FMA (no -march in its flags), every
 * float multiply/add below rounds separately.
 */
#include "oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* One cell of a rolling row: src/AlignmentMatrixFast.h:34-54 */
typedef struct {
  float score;
  short run;
  signed char dir;
} cell_t;

static const cell_t EMPTY_CELL = {0.0f, 0, OR_STOP};

typedef struct {
  const or_scoring* sc;
  const char* ref;
  const char* qry;
  int ref_len, height;
  const int* off;
  const int* len;
  uint64_t* row_start; /* offsetInMatrix, src/AlignmentMatrixFast.cpp:38-41 */
  unsigned char* dirs;
  cell_t *cur, *last;
  int cur_off, cur_len, last_off, last_len; /* currentCorridor / lastCorridor */
  int have_last;
  float best;
  int best_x, best_y;
} fill_t;

static inline float fminf2(float a, float b) { return b < a ? b : a; } /* std::min(a,b) */
static inline float fmaxf2(float a, float b) { return a < b ? b : a; } /* std::max(a,b) */

/* getElementUp, src/AlignmentMatrixFast.h:74-90 */
static inline cell_t up_at(const fill_t* f, int x, int y_up) {
  if (y_up < 0 || x < 0) return EMPTY_CELL;
  if (x < f->last_off || x >= f->last_off + f->last_len) return EMPTY_CELL;
  return f->last[x - f->last_off];
}

/* getElementCurr, src/AlignmentMatrixFast.h:98-112 */
static inline cell_t cur_at(const fill_t* f, int x, int y) {
  if (y < 0 || x < 0) return EMPTY_CELL;
  if (x < f->cur_off || x >= f->cur_off + f->cur_len) return EMPTY_CELL;
  return f->cur[x - f->cur_off];
}

/* Convex extension penalty for a gap that already holds `run` cells,
 * std::min(gap_ext_min, gap_ext + run * gap_decay)  (src/ConvexAlignFast.cpp:672-674) */
static inline float ext_pen(const or_scoring* sc, int run) {
  float r = (float)run;
  float t = r * sc->gap_decay;
  t = sc->gap_ext + t;
  return fminf2(sc->gap_ext_min, t);
}

static inline void put(fill_t* f, int x, int y, cell_t c) {
  f->cur[x - f->cur_off] = c;
  if (f->dirs) f->dirs[f->row_start[y] + (uint64_t)(x - f->off[y])] = (unsigned char)c.dir;
}

static inline void track(fill_t* f, float s, int x, int y) {
  if (s > f->best) { /* strict: first maximum in visiting order (:1165-1170, :1270-1275) */
    f->best = s;
    f->best_x = x;
    f->best_y = y;
  }
}

/* The scalar cell rule: src/ConvexAlignFast.cpp:1179-1277 (tail) == :650-738 (fwdFillMatrix) */
static inline void scalar_cell(fill_t* f, int x, int y) {
  const or_scoring* sc = f->sc;
  float diag_score = up_at(f, x - 1, y - 1).score;
  cell_t up = up_at(f, x, y - 1);
  cell_t left = cur_at(f, x - 1, y);
  int eq = f->qry[y] == f->ref[x];
  float diag_cell = diag_score + (eq ? sc->mat : sc->mis);
  float up_cell, left_cell;
  int ins_run = 0, del_run = 0;
  if (up.dir == OR_I) {
    ins_run = up.run;
    up_cell = (up.score == 0) ? 0.0f : up.score + ext_pen(sc, ins_run);
  } else {
    up_cell = up.score + sc->gap_open_read;
  }
  if (left.dir == OR_D) {
    del_run = left.run;
    left_cell = (left.score == 0) ? 0.0f : left.score + ext_pen(sc, del_run);
  } else {
    left_cell = left.score + sc->gap_open_ref;
  }
  float m = 0.0f;
  m = fmaxf2(left_cell, m);
  m = fmaxf2(diag_cell, m);
  m = fmaxf2(up_cell, m);
  cell_t c;
  if (del_run > 0 && m == left_cell) {
    c.score = m; c.dir = OR_D; c.run = (short)(del_run + 1);
  } else if (ins_run > 0 && m == up_cell) {
    c.score = m; c.dir = OR_I; c.run = (short)(ins_run + 1);
  } else if (m == diag_cell) {
    c.score = m; c.dir = eq ? OR_EQ : OR_X; c.run = 0;
  } else if (m == left_cell) {
    c.score = m; c.dir = OR_D; c.run = 1;
  } else if (m == up_cell) {
    c.score = m; c.dir = OR_I; c.run = 1;
  } else {
    c.score = 0.0f; c.dir = OR_STOP; c.run = 0;
  }
  put(f, x, y, c);
  track(f, m, x, y);
}

/* One 4-wide block of the SSE path, lane by lane: the vector part (:950-1091) has no cross-lane
 * dependency, so each lane is evaluated in turn; the left fix-up (:1103-1174) is sequential in
 * the reference as well. */
static inline void sse_block(fill_t* f, int x0, int y) {
  const or_scoring* sc = f->sc;
  float v_score[4], v_dir[4], v_run[4], v_uprun[4];
  for (int j = 0; j < 4; ++j) {
    int x = x0 + j;
    /* protected and unprotected loads (:966-992) fetch the same cells */
    cell_t up = up_at(f, x, y - 1);
    float diag_score = up_at(f, x - 1, y - 1).score;
    int eq = ((float)f->qry[y] == (float)f->ref[x]); /* _mm_cmpeq_ps on chars cast to float */
    float diag_cell = diag_score + (eq ? sc->mat : sc->mis);
    float up_dir = (float)up.dir, up_run = (float)up.run;
    float up_cell;
    if (up_dir == (float)OR_I) {
      if (up.score == 0.0f) {
        up_cell = 0.0f;
      } else {
        float t = up_run * sc->gap_decay;
        t = sc->gap_ext + t;
        up_cell = up.score + fminf2(sc->gap_ext_min, t); /* _mm_min_ps(a,b) = a<b?a:b; equal for non-NaN */
      }
    } else {
      up_cell = up.score + sc->gap_open_read;
    }
    float m = fmaxf2(fmaxf2(up_cell, diag_cell), 0.0f);
    float dir = (float)OR_STOP, run = 0.0f;
    if (m == up_cell) { dir = (float)OR_I; run = 1.0f; }                      /* cmp_c :1052-1058 */
    if (m == diag_cell) { dir = eq ? (float)OR_EQ : (float)OR_X; run = 0.0f; } /* cmp_b :1071-1079 */
    if (up_run > 0.0f && m == up_cell) { dir = (float)OR_I; run = up_run + 1.0f; } /* cmp_a :1087-1094 */
    v_score[j] = m; v_dir[j] = dir; v_run[j] = run; v_uprun[j] = up_run;
  }
  cell_t left = cur_at(f, x0 - 1, y);
  for (int j = 0; j < 4; ++j) {
    int x = x0 + j;
    float left_cell;
    if (left.dir == OR_D) {
      left_cell = (left.score == 0) ? 0.0f : left.score + ext_pen(sc, left.run);
    } else {
      left_cell = left.score + sc->gap_open_ref;
    }
    cell_t c;
    c.score = v_score[j];
    c.dir = (signed char)v_dir[j];
    c.run = (short)(int)v_run[j];
    if (left_cell >= c.score) {
      if (left.run > 0) {
        c.score = left_cell; c.dir = OR_D; c.run = (short)(left.run + 1);
      } else if (left_cell > c.score || c.dir == OR_STOP || (c.dir == OR_I && v_uprun[j] <= 0)) {
        c.score = left_cell; c.dir = OR_D; c.run = 1;
      }
    }
    put(f, x, y, c);
    track(f, c.score, x, y);
    left = c;
  }
}

/* Derived single-pass form of the SSE path (rule 2), the executable spec of the CUDA kernel's
 * "raw-run" mode: the vector blends + left fix-up of sse_block() collapse to the scalar priority
 * chain with the neighbours' RAW indelRun (not gated by their direction) in the run tests, while
 * the extension formula stays gated by direction. Differs from rule 0 only in that it does not
 * replay the SSE values of the <= 11 cells the tail recomputes into the best-cell tracking. */
static inline void chain_cell(fill_t* f, int x, int y, int raw) {
  const or_scoring* sc = f->sc;
  float diag_score = up_at(f, x - 1, y - 1).score;
  cell_t up = up_at(f, x, y - 1);
  cell_t left = cur_at(f, x - 1, y);
  int eq = f->qry[y] == f->ref[x];
  float diag_cell = diag_score + (eq ? sc->mat : sc->mis);
  float up_cell, left_cell;
  if (up.dir == OR_I) up_cell = (up.score == 0) ? 0.0f : up.score + ext_pen(sc, up.run);
  else up_cell = up.score + sc->gap_open_read;
  if (left.dir == OR_D) left_cell = (left.score == 0) ? 0.0f : left.score + ext_pen(sc, left.run);
  else left_cell = left.score + sc->gap_open_ref;
  int ins_run = (raw || up.dir == OR_I) ? up.run : 0;
  int del_run = (raw || left.dir == OR_D) ? left.run : 0;
  float m = fmaxf2(fmaxf2(fmaxf2(left_cell, 0.0f), diag_cell), up_cell);
  cell_t c;
  if (del_run > 0 && m == left_cell) { c.score = m; c.dir = OR_D; c.run = (short)(del_run + 1); }
  else if (ins_run > 0 && m == up_cell) { c.score = m; c.dir = OR_I; c.run = (short)(ins_run + 1); }
  else if (m == diag_cell) { c.score = m; c.dir = eq ? OR_EQ : OR_X; c.run = 0; }
  else if (m == left_cell) { c.score = m; c.dir = OR_D; c.run = 1; }
  else if (m == up_cell) { c.score = m; c.dir = OR_I; c.run = 1; }
  else { c.score = 0.0f; c.dir = OR_STOP; c.run = 0; }
  put(f, x, y, c);
  track(f, m, x, y);
}

static int fill_run(fill_t* f, int rule) {
  f->best = -1.0f;
  f->best_x = 0;
  f->best_y = 0;
  for (int y = 0; y < f->height; ++y) {
    /* prepareLine: swap rows (src/AlignmentMatrixFast.cpp:197-211); contents are NOT cleared */
    cell_t* tmp = f->last;
    f->last = f->cur;
    f->last_off = f->cur_off;
    f->last_len = f->cur_len;
    f->cur = tmp;
    f->cur_off = f->off[y];
    f->cur_len = f->len[y];
    int x_off = f->off[y];
    int x_max = x_off + f->len[y];
    if (f->ref_len < x_max) x_max = f->ref_len;
    int x_min = x_off > 0 ? x_off : 0;
    if (rule == 0) {
      for (int x = x_min; x < x_max - 4; x += 4) sse_block(f, x, y);
      int t0 = x_max - 4 - 8;
      if (t0 < x_min) t0 = x_min;
      for (int x = t0; x < x_max; ++x) scalar_cell(f, x, y);
    } else if (rule == 2) {
      int t0 = x_max - 4 - 8;
      if (t0 < x_min) t0 = x_min;
      for (int x = x_min; x < x_max; ++x) chain_cell(f, x, y, x < t0);
    } else {
      for (int x = x_min; x < x_max; ++x) scalar_cell(f, x, y);
    }
  }
  return 0;
}

static int fill_init(fill_t* f, const or_scoring* sc, const char* ref, int ref_len, const char* qry,
                     int height, const int* offsets, const int* lengths, unsigned char* dirs,
                     uint64_t* total_out) {
  memset(f, 0, sizeof(*f));
  f->sc = sc; f->ref = ref; f->qry = qry; f->ref_len = ref_len; f->height = height;
  f->off = offsets; f->len = lengths; f->dirs = dirs;
  f->row_start = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(height > 0 ? height : 1));
  uint64_t total = 0;
  int maxlen = 0;
  for (int y = 0; y < height; ++y) {
    f->row_start[y] = total;
    total += (uint64_t)(lengths[y] > 0 ? lengths[y] : 0);
    if (lengths[y] > maxlen) maxlen = lengths[y];
  }
  if (total_out) *total_out = total;
  f->cur = (cell_t*)malloc(sizeof(cell_t) * (size_t)(maxlen + 1));
  f->last = (cell_t*)malloc(sizeof(cell_t) * (size_t)(maxlen + 1));
  for (int i = 0; i <= maxlen; ++i) f->cur[i] = f->last[i] = EMPTY_CELL;
  /* before the first prepareLine both corridors are {0,0,0} (default-constructed CorridorLine) */
  f->cur_off = f->cur_len = f->last_off = f->last_len = 0;
  return 0;
}

static void fill_free(fill_t* f) {
  free(f->row_start);
  free(f->cur);
  free(f->last);
}

int64_t or_convex_cells(int ref_len, int height, const int* offsets, const int* lengths) {
  int64_t n = 0;
  for (int y = 0; y < height; ++y) {
    int lo = offsets[y] > 0 ? offsets[y] : 0;
    int hi = offsets[y] + lengths[y];
    if (hi > ref_len) hi = ref_len;
    if (hi > lo) n += hi - lo;
  }
  return n;
}

int or_convex_fill(const or_scoring* sc, const char* ref, int ref_len, const char* qry, int height,
                   const int* offsets, const int* lengths, int rule, unsigned char* dirs,
                   float* best_score, int* best_ref, int* best_read) {
  fill_t f;
  uint64_t total;
  fill_init(&f, sc, ref, ref_len, qry, height, offsets, lengths, dirs, &total);
  if (dirs) memset(dirs, 0xFF, total);
  fill_run(&f, rule);
  *best_score = f.best;
  *best_ref = f.best_x;
  *best_read = f.best_y;
  fill_free(&f);
  return 0;
}

/* getDirection, src/AlignmentMatrixFast.cpp:185-195 */
static inline int dir_at(const fill_t* f, int x, int y) {
  if (y < 0 || y > f->height - 1 || x < 0) return OR_STOP;
  if (x < f->off[y] || x >= f->off[y] + f->len[y]) return OR_STOP;
  return (signed char)f->dirs[f->row_start[y] + (uint64_t)(x - f->off[y])];
}

/* validPath, src/AlignmentMatrixFast.cpp:213-220 (float math, then truncation to int) */
static inline int valid_path(const fill_t* f, int x, int y) {
  int width = f->len[y];
  int min_c = (int)((float)f->off[y] + 0.1f * (float)width);
  int max_c = (int)((float)(min_c + width) - 0.1f * (float)width);
  return x > min_c && x < max_c;
}

typedef struct {
  int* bc;   /* binaryCigar */
  int cap;   /* maxBinaryCigarLength */
  int offset; /* alignment_offset */
  int ref_position, qstart, qend;
} trace_t;

/* revBacktrack. Returns 1 valid, 0 invalid, -1 "throw 1" (binary CIGAR buffer exhausted). */
static int traceback(const fill_t* f, trace_t* t) {
  if (f->best_y <= 0) return 0;
  int idx = t->cap - 1;
  int op = OR_S;
  int op_len = t->qend;
  int read_len = t->qend;
  int x = f->best_x, y = f->best_y;
  int d;
  while ((d = dir_at(f, x, y)) != OR_STOP) {
    if (!valid_path(f, x, y)) return 0;
    if (d == OR_X || d == OR_EQ) {
      y -= 1; x -= 1; read_len += 1;
    } else if (d == OR_I) {
      y -= 1; read_len += 1;
    } else if (d == OR_D) {
      x -= 1;
    } else {
      return 0;
    }
    if (d == op) {
      op_len += 1;
    } else {
      t->bc[idx--] = (op_len << 4) | op;
      op = d;
      op_len = 1;
    }
    if (idx < 0) return -1;
  }
  t->bc[idx--] = (op_len << 4) | op;
  /* the reference writes the next slot unchecked; idx == -1 here would be a buffer underrun */
  if (idx < 0) return -1;
  t->bc[idx--] = ((y + 1) << 4) | OR_S;
  read_len += y + 1;
  t->ref_position = x + 1;
  t->qstart = y + 1;
  t->offset = idx + 1;
  return f->height == read_len ? 1 : 0;
}

static int popcount32(uint32_t i) { /* NumberOfSetBits :21-27 */
  i = i - ((i >> 1) & 0x55555555u);
  i = (i & 0x33333333u) + ((i >> 2) & 0x33333333u);
  return (int)((((i + (i >> 4)) & 0x0F0F0F0Fu) * 0x01010101u) >> 24);
}

typedef struct {
  char* buf;
  size_t len, cap;
} sbuf;

static void sb_reserve(sbuf* s, size_t extra) {
  if (s->len + extra + 1 > s->cap) {
    while (s->len + extra + 1 > s->cap) s->cap = s->cap * 2 + 64;
    s->buf = (char*)realloc(s->buf, s->cap);
  }
}
static void sb_int(sbuf* s, int v) {
  sb_reserve(s, 16);
  s->len += (size_t)sprintf(s->buf + s->len, "%d", v);
}
static void sb_chr(sbuf* s, char c) {
  sb_reserve(s, 1);
  s->buf[s->len++] = c;
}

/* convertCigar. ref points at refSeq + ref_position. */
static int convert(const trace_t* t, const char* ref, int ext_qstart, int ext_qend, or_align_out* o,
                   sbuf* cigar, sbuf* md, int* nm_out, int nm_cap) {
  uint32_t window = 0;
  int pos_ref = 0, pos_read;
  int ops = 0, nm_idx = 0, exact_len = 0, final_len = 0;
  int j0 = t->offset;
  o->sv_type = 0;
  o->qstart = (t->bc[j0] >> 4) + ext_qstart;
  if (o->qstart > 0) {
    sb_int(cigar, o->qstart); sb_chr(cigar, 'S'); ops++;
    final_len += o->qstart;
  }
  pos_read = t->bc[j0] >> 4;
  o->first_ref = pos_ref;
  o->first_read = pos_read;
  int matches = 0, aln_len = 0, m_len = 0, md_eq = 0, ri = 0, yi = 0;
#define NMPOS(incr_ref, incr_read)                                                   \
  do {                                                                               \
    if (pos_read > 16 && pos_ref > 16) {                                             \
      if (nm_out && nm_idx < nm_cap) {                                               \
        nm_out[3 * nm_idx + 0] = pos_ref - 16;                                       \
        nm_out[3 * nm_idx + 1] = pos_read - 16;                                      \
        nm_out[3 * nm_idx + 2] = yi;                                                 \
      }                                                                              \
      nm_idx++;                                                                      \
    }                                                                                \
    pos_ref += (incr_ref);                                                           \
    pos_read += (incr_read);                                                         \
  } while (0)
  for (int j = j0 + 1; j < t->cap - 1; ++j) {
    int op = t->bc[j] & 15, n = t->bc[j] >> 4;
    aln_len += n;
    switch (op) {
      case OR_X:
        m_len += n;
        for (int k = 0; k < n; ++k) {
          sb_int(md, md_eq); md_eq = 0;
          sb_chr(md, ref[ri++]);
          window = (window << 1) | 1u;
          yi = popcount32(window);
          NMPOS(1, 1);
        }
        exact_len += n;
        break;
      case OR_EQ:
        m_len += n; md_eq += n; matches += n;
        for (int k = 0; k < n; ++k) {
          window <<= 1;
          yi = popcount32(window);
          NMPOS(1, 1);
        }
        ri += n;
        exact_len += n;
        break;
      case OR_D:
        if (m_len > 0) { sb_int(cigar, m_len); sb_chr(cigar, 'M'); ops++; final_len += m_len; m_len = 0; }
        sb_int(cigar, n); sb_chr(cigar, 'D'); ops++;
        sb_int(md, md_eq); md_eq = 0;
        sb_chr(md, '^');
        for (int k = 0; k < n; ++k) {
          sb_chr(md, ref[ri++]);
          window <<= 1;
          if (k < 1) { window |= 1u; yi = yi + 1 > 0 ? yi + 1 : 0; }
          NMPOS(1, 0);
        }
        exact_len += n;
        break;
      case OR_I:
        if (m_len > 0) { sb_int(cigar, m_len); sb_chr(cigar, 'M'); ops++; final_len += m_len; m_len = 0; }
        sb_int(cigar, n); sb_chr(cigar, 'I'); ops++;
        final_len += n;
        for (int k = 0; k < n; ++k) {
          window <<= 1;
          if (k < 1) { window |= 1u; yi = yi + 1 > 0 ? yi + 1 : 0; }
          pos_read += 1;
        }
        exact_len += n;
        break;
      default:
        return -1; /* "Invalid cigar string" -> throw 1 */
    }
  }
#undef NMPOS
  sb_int(md, md_eq);
  if (m_len > 0) { sb_int(cigar, m_len); sb_chr(cigar, 'M'); ops++; final_len += m_len; }
  o->qend = (t->bc[t->cap - 1] >> 4) + ext_qend;
  if (o->qend > 0) { sb_int(cigar, o->qend); sb_chr(cigar, 'S'); ops++; }
  final_len += o->qend;
  o->identity = (float)matches * 1.0f / (float)aln_len;
  o->nm = aln_len - matches;
  o->alignment_length = exact_len;
  o->last_ref = pos_ref;
  o->last_read = pos_read;
  o->cigar_op_count = ops;
  o->nm_count = nm_idx;
  sb_reserve(cigar, 1); cigar->buf[cigar->len] = 0;
  sb_reserve(md, 1); md->buf[md->len] = 0;
  return final_len;
}

int or_convex_single_align(const or_scoring* sc, const char* ref, const char* qry, const int* offsets,
                           const int* lengths, int height, int ext_qstart, int ext_qend, int rule,
                           or_align_out* o, char* cigar_out, int cigar_cap, char* md_out,
                           int md_cap, int* nm_out, int nm_cap) {
  int ref_len = (int)strlen(ref), qry_len = (int)strlen(qry);
  memset(o, 0, sizeof(*o));
  o->sv_type = 0;
  o->score = -1.0f;
  o->ret = -1;
  if (cigar_out && cigar_cap > 0) cigar_out[0] = 0;
  if (md_out && md_cap > 0) md_out[0] = 0;
  fill_t f;
  uint64_t total = 0;
  /* prepare(): matrix height is qryLen, corridor rows = corridorHeight (== qryLen for every caller) */
  fill_init(&f, sc, ref, ref_len, qry, height, offsets, lengths, NULL, &total);
  f.height = qry_len < height ? qry_len : height; /* rows actually filled: matrix->getHeight() */
  if ((unsigned long)((float)total / 1000.0f / 1000.0f) >= 10000ul) { /* maxMatrixSizeMB, IConfig.h:47 */
    fill_free(&f);
    return 2;
  }
  f.dirs = (unsigned char*)malloc(total + 1);
  fill_run(&f, rule);
  trace_t t;
  t.cap = 200000;
  if (t.cap < qry_len) t.cap = qry_len + 1; /* :480-485 */
  t.bc = (int*)malloc(sizeof(int) * (size_t)t.cap);
  t.qend = f.height - f.best_y - 1; /* :1281 */
  int status = 0;
  int valid = traceback(&f, &t);
  if (valid < 0) {
    status = 1;
  } else if (valid) {
    sbuf cigar = {0, 0, 0}, md = {0, 0, 0};
    int r = convert(&t, ref + t.ref_position, ext_qstart, ext_qend, o, &cigar, &md, nm_out, nm_cap);
    if (r < 0) {
      status = 1;
    } else {
      o->ret = r;
      o->position_offset = t.ref_position;
      o->score = f.best;
      /* N-clip probe :498-529; the decoder never emits 'X' so this stays 0 in practice */
      int ncount = 0, probes = 0;
      int lo = t.ref_position - 100 > 0 ? t.ref_position - 100 : 0;
      for (int k = t.ref_position; k > lo; --k) { if (ref[k] == 'X') ncount++; probes++; }
      if ((float)ncount > (float)probes * 0.8f) o->sv_type |= 1;
      ncount = probes = 0;
      int hi = o->last_ref + 100 < ref_len - t.ref_position ? o->last_ref + 100 : ref_len - t.ref_position;
      for (int k = o->last_ref; k < hi; ++k) { if (ref[t.ref_position + k] == 'X') ncount++; probes++; }
      if ((float)ncount > (float)probes * 0.8f) o->sv_type |= 1;
      if (cigar_out && cigar_cap > 0) { strncpy(cigar_out, cigar.buf, (size_t)cigar_cap - 1); cigar_out[cigar_cap - 1] = 0; }
      if (md_out && md_cap > 0) { strncpy(md_out, md.buf, (size_t)md_cap - 1); md_out[md_cap - 1] = 0; }
    }
    free(cigar.buf);
    free(md.buf);
  }
  free(t.bc);
  free(f.dirs);
  fill_free(&f);
  return status;
}
