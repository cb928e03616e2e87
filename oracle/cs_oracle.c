/* oracle/cs_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Scalar C restatement of ngmlr's candidate search (stage 0) and the data it runs on:
 *   reference layout/encoding  _SequenceProvider::Init, enc4/dec4   src/SequenceProvider.cpp:76-105, 292-400
 *   window decode              _SequenceProvider::DecodeRefSequence  src/SequenceProvider.cpp:567-625
 *   k-mer iteration            CS::PrefixIteration                   src/CSstatic.cpp:17-73
 *   index build                CompactPrefixTable::CountKmer / createRefTableIndex / BuildPrefixTable /
 *                              SaveToRefTable / Generate             src/PrefixTable.cpp:233-321, 372-474
 *   reverse complement         revComp                               src/PrefixTable.cpp:70-88
 *   lookup                     CompactPrefixTable::GetRefEntry       src/PrefixTable.cpp:476-532
 *   vote                       CS::PrefixSearch / AddLocationStd     src/CS.cpp:57-149
 *   collect                    CS::CollectResultsStd                 src/CS.cpp:217-269
 * Pinned against oracle/_ref/libngmlr_full.so (the whole unmodified reference) in tests/test_cs_oracle.py.
 *
 * The vote's open-addressing table is an implementation detail of the reference (results do not
 * depend on its size unless it overflows, and then the reference retries with a larger one,
 * src/CS.cpp:324-398); the restatement uses a table sized to the read and keeps exactly what is
 * observable: counts per (bin, strand), the running threshold, and the emission order.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

struct or_cs {
  uint8_t* enc; /* binRef: 2 bases per byte, A0 T1 G2 C3 N4 */
  uint64_t enc_bytes;
  uint64_t concat_len; /* GetConcatRefLen() = 2*bytes - 1 */
  int nref;
  uint64_t* ref_start;
  uint64_t* ref_len;
  int k, skip, bin_shift, max_freq;
  uint32_t index_len; /* 4^k + 1 */
  uint32_t* tab;      /* Index::m_TabIndex */
  int8_t* rci;        /* Index::m_RevCompIndex (used() <=> != 0) */
  uint32_t* pos;      /* Location::m_Location */
  uint32_t npos;
};

static inline int enc4(char c) {
  if (c >= 'a' && c <= 'z') c = (char)(c - 32);
  switch (c) {
    case 'A': return 0;
    case 'T': return 1;
    case 'G': return 2;
    case 'C': return 3;
  }
  return 4;
}
static const char DEC4[5] = {'A', 'T', 'G', 'C', 'N'};

/* k-mer code of CS: A0 C1 T2 G3 = (c >> 1) & 3 (src/CSstatic.cpp:17-20) */
static inline uint64_t kenc(char c) { return (uint64_t)((c >> 1) & 3); }

typedef void (*kmer_fn)(uint64_t prefix, uint64_t pos, void* data);

/* CS::PrefixIteration, the tail recursion unrolled into a loop. */
static void prefix_iteration(const char* seq, uint64_t length, int k, unsigned skip, uint64_t offset,
                             kmer_fn fn, void* data) {
  const uint64_t mask = ((uint64_t)1 << (2 * k)) - 1;
  for (;;) {
    if (length < (uint64_t)k) return;
    if (*seq == 'N') {
      unsigned n_skip = 1;
      while (seq[n_skip] == 'N') ++n_skip;
      seq += n_skip;
      if (n_skip >= (length - (uint64_t)k)) return;
      length -= n_skip;
      offset += n_skip;
    }
    uint64_t prefix = 0;
    uint64_t i;
    int restarted = 0;
    for (i = 0; i < (uint64_t)k - 1; ++i) {
      char c = seq[i];
      if (c == 'N') { restarted = 1; break; }
      prefix = (prefix << 2) | kenc(c);
    }
    if (!restarted) {
      unsigned skipcount = skip;
      for (i = (uint64_t)k - 1; i < length; ++i) {
        char c = seq[i];
        if (c == 'N') { restarted = 1; break; }
        prefix = ((prefix << 2) | kenc(c)) & mask;
        if (skipcount == skip) {
          fn(prefix, offset + i + 1 - (uint64_t)k, data);
          skipcount = 0;
        } else {
          ++skipcount;
        }
      }
    }
    if (!restarted) return;
    /* PrefixIteration(sequence + i + 1, length - i - 1, ..., offset + i + 1) */
    seq += i + 1;
    length -= i + 1;
    offset += i + 1;
  }
}

static inline uint64_t rev_comp(uint64_t prefix, int k) {
  const uint64_t mask = ((uint64_t)1 << (2 * k)) - 1;
  uint64_t c = (prefix ^ 0xAAAAAAAAAAAAAAAAull) & mask;
  uint64_t r = 0;
  for (int i = 0; i < k; ++i) {
    r = (r << 2) | (c & 3);
    c >>= 2;
  }
  return r;
}

/* ---- reference layout -------------------------------------------------------------------- */

struct or_cs* or_cs_create(int ncontigs, const char* const* contigs, const int64_t* lens) {
  struct or_cs* h = (struct or_cs*)calloc(1, sizeof(*h));
  uint64_t bytes = 500;
  int kept = 0;
  for (int i = 0; i < ncontigs; ++i)
    if (lens[i] > 10) { bytes += (uint64_t)(lens[i] + 1) / 2 + 500; ++kept; } /* minRefSeqLen = 10 */
  h->enc = (uint8_t*)malloc(bytes + 16);
  h->ref_start = (uint64_t*)calloc((size_t)kept + 1, sizeof(uint64_t));
  h->ref_len = (uint64_t*)calloc((size_t)kept + 1, sizeof(uint64_t));
  uint64_t bi = 0;
  const uint8_t nn = (uint8_t)((4 << 4) | 4);
  for (int i = 0; i < 500; ++i) h->enc[bi++] = nn; /* padding to avoid negative positions */
  int j = 0;
  for (int i = 0; i < ncontigs; ++i) {
    if (!(lens[i] > 10)) continue;
    h->ref_start[j] = bi * 2;
    h->ref_len[j] = (uint64_t)lens[i];
    ++j;
    const char* s = contigs[i];
    int64_t L = lens[i];
    for (int64_t p = 0; p + 1 < L; p += 2) h->enc[bi++] = (uint8_t)((enc4(s[p]) << 4) | enc4(s[p + 1]));
    if (L & 1) h->enc[bi++] = (uint8_t)((enc4(s[L - 1]) << 4) | 4);
    for (int q = 0; q < 500; ++q) h->enc[bi++] = nn;
  }
  h->nref = kept;
  h->enc_bytes = bi;
  h->concat_len = bi * 2 - 1;
  return h;
}

void or_cs_destroy(struct or_cs* h) {
  if (!h) return;
  free(h->enc); free(h->ref_start); free(h->ref_len); free(h->tab); free(h->rci); free(h->pos);
  free(h);
}

uint64_t or_cs_concat_len(const struct or_cs* h) { return h->concat_len; }
int or_cs_ref_count(const struct or_cs* h) { return h->nref; }
uint64_t or_cs_ref_start(const struct or_cs* h, int i) { return h->ref_start[i]; }
const uint8_t* or_cs_encoded(const struct or_cs* h, uint64_t* bytes) { *bytes = h->enc_bytes; return h->enc; }

static inline char base_at(const struct or_cs* h, uint64_t p) {
  uint8_t b = h->enc[p >> 1];
  return DEC4[(p & 1) ? (b & 0xF) : (b >> 4)];
}

/* DecodeRefSequence(sequence, 0, position, bufferLength). Returns 0 for an invalid position. */
int or_cs_decode(const struct or_cs* h, uint64_t position, uint64_t buffer_len, char* out) {
  uint64_t len = buffer_len - 2;
  if (position >= h->concat_len) return 0;
  uint64_t end = 0;
  if (position + len > h->concat_len) {
    end = position + len - h->concat_len;
    len -= end;
  }
  uint64_t start = (position + 1) / 2;
  uint64_t ci = 0;
  if (position & 1) out[ci++] = DEC4[h->enc[start - 1] & 0xF];
  for (uint64_t i = 0; i < (len + 1) / 2; ++i) {
    out[ci++] = DEC4[h->enc[start + i] >> 4];
    out[ci++] = DEC4[h->enc[start + i] & 0xF];
  }
  if (len & 1) out[ci - 1] = 'x';
  for (uint64_t i = 0; i < end; ++i) out[ci++] = 'x';
  for (uint64_t i = ci; i < buffer_len; ++i) out[i] = 0;
  return 1;
}

/* _SequenceProvider::getChrStart (src/SequenceProvider.cpp:157-178) over refStartPos (:416-424:
 * the start of every contig plus one artificial entry last_start + last_len + 1000). */
static void chr_bounds(const struct or_cs* h, uint64_t position, uint64_t* chr_start, uint64_t* chr_end) {
  /* std::upper_bound: first entry > position */
  int n = h->nref + 1, u = 0;
  while (u < n) {
    uint64_t v = u < h->nref ? h->ref_start[u] : h->ref_start[h->nref - 1] + h->ref_len[h->nref - 1] + 1000;
    if (v > position) break;
    ++u;
  }
  uint64_t up = u < h->nref ? h->ref_start[u] : h->ref_start[h->nref - 1] + h->ref_len[h->nref - 1] + 1000;
  if (up - position < 1000) { /* inside the spacer before the next contig */
    ++u;
    up = u < h->nref ? h->ref_start[u] : h->ref_start[h->nref - 1] + h->ref_len[h->nref - 1] + 1000;
  }
  *chr_start = h->ref_start[u - 1];
  *chr_end = up - 1000;
}

/* _SequenceProvider::decode (src/SequenceProvider.cpp:475-490) */
static uint64_t decode_range(const struct or_cs* h, uint64_t start_pos, uint64_t end_pos, char* out) {
  uint64_t ci = 0, start = (start_pos + 1) / 2, n = (end_pos - start_pos + 1) / 2;
  if (start_pos & 1) out[ci++] = DEC4[h->enc[start - 1] & 0xF];
  for (uint64_t i = 0; i < n; ++i) {
    out[ci++] = DEC4[h->enc[start + i] >> 4];
    out[ci++] = DEC4[h->enc[start + i] & 0xF];
  }
  return ci;
}

/* _SequenceProvider::DecodeRefSequenceExact(sequence, startPosition, sequenceLength, corridor)
 * (src/SequenceProvider.cpp:493-565), the call extractReferenceSequenceForAlignment makes with
 * corridor 0 (src/AlignmentBuffer.cpp:215). `out` needs sequence_len + 2 bytes (decode() writes whole
 * byte pairs and may run up to 2 characters past the requested length, hence the reference's
 * "+ 100"); bytes [0, sequence_len) are the result, NUL-terminated. Returns 0 for an invalid start. */
int or_cs_decode_exact(const struct or_cs* h, uint64_t start, uint64_t sequence_len, int corridor, char* out) {
  if (start >= h->concat_len) return 0;
  memset(out, 'x', sequence_len);
  const uint64_t half = (uint64_t)(corridor / 2);
  uint64_t chr_start, chr_end;
  chr_bounds(h, start, &chr_start, &chr_end);
  uint64_t dstart = start - half;
  const uint64_t end = start + sequence_len - half;
  uint64_t dend = end;
  if (end > chr_end) dend -= end - chr_end;
  if (half > start) {
    dstart = chr_start;
    const uint64_t diff = half - dstart + 1000 - (start - chr_start);
    decode_range(h, dstart, dend, out + diff);
  } else if (dstart < chr_start) {
    if (dend > chr_start) {
      const uint64_t diff = chr_start - dstart;
      dstart += diff;
      decode_range(h, dstart, dend, out + diff);
    }
  } else {
    decode_range(h, dstart, dend, out);
  }
  out[sequence_len - 1] = 0;
  return 1;
}

/* ---- index build --------------------------------------------------------------------------- */

typedef struct {
  struct or_cs* h;
  int32_t* freq;
  uint64_t last_prefix;
  int64_t last_bin;
  uint32_t* fill; /* next free slot per prefix while building */
  int building;
} build_t;

static void build_cb(uint64_t prefix, uint64_t pos, void* data) {
  build_t* b = (build_t*)data;
  int keep;
  if (prefix == b->last_prefix) {
    int64_t bin = (int64_t)(pos >> b->h->bin_shift);
    keep = (bin != b->last_bin || b->last_bin == -1);
    b->last_bin = bin;
  } else {
    b->last_bin = -1;
    keep = 1;
  }
  b->last_prefix = prefix;
  if (!keep) return;
  if (!b->building) {
    b->freq[prefix] += 1;
  } else if (b->h->rci[prefix] != 0) {
    b->h->pos[b->fill[prefix]++] = (uint32_t)pos; /* unit offset is 0 below 4 G positions */
  }
}

static void sweep(build_t* b) {
  struct or_cs* h = b->h;
  for (int r = 0; r < h->nref; ++r) {
    b->last_prefix = 111111;
    b->last_bin = -1;
    uint64_t L = h->ref_len[r];
    char* seq = (char*)malloc(L + 2);
    /* Generate() decodes with DecodeRefSequence(seq, i, offset, len), whose bufferLength argument
     * loses 2 characters (src/SequenceProvider.cpp:569): the last two positions of every contig
     * reach PrefixIteration as 'x' / NUL, which the k-mer code maps to 0 like 'A'. */
    for (uint64_t p = 0; p + 2 < L; ++p) seq[p] = base_at(h, h->ref_start[r] + p);
    if (L >= 2) {
      seq[L - 2] = ((L - 2) & 1) ? 'x' : 0;
      seq[L - 1] = 0;
    }
    seq[L] = 0; seq[L + 1] = 0;
    prefix_iteration(seq, L, h->k, (unsigned)h->skip, h->ref_start[r], build_cb, b);
    free(seq);
  }
}

int or_cs_build_index(struct or_cs* h, int k, int skip, int bin_shift, int max_freq) {
  h->k = k; h->skip = skip; h->bin_shift = bin_shift; h->max_freq = max_freq;
  uint32_t length = ((uint32_t)1 << (2 * k)) + 1;
  h->index_len = length;
  build_t b;
  memset(&b, 0, sizeof(b));
  b.h = h;
  b.freq = (int32_t*)calloc(length, sizeof(int32_t));
  b.building = 0;
  sweep(&b);
  h->tab = (uint32_t*)calloc((size_t)length + 1, sizeof(uint32_t));
  h->rci = (int8_t*)calloc((size_t)length + 1, 1);
  uint32_t next = 0, i;
  for (i = 0; i < length - 1; ++i) {
    int f = b.freq[i];
    int total = f + b.freq[rev_comp(i, k)];
    h->tab[i] = next + 1;
    if (f > 0 && total < max_freq) {
      h->rci[i] = (int8_t)(char)((float)(max_freq - total) * 100.0f / (float)max_freq);
      next += (uint32_t)f;
    }
  }
  h->tab[i] = next + 1;
  h->npos = next;
  h->pos = (uint32_t*)calloc((size_t)next + 2, sizeof(uint32_t));
  b.fill = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)length);
  for (i = 0; i < length - 1; ++i) b.fill[i] = h->tab[i] - 1;
  b.building = 1;
  sweep(&b);
  free(b.fill);
  free(b.freq);
  return 0;
}

uint32_t or_cs_index_len(const struct or_cs* h) { return h->index_len; }
uint32_t or_cs_npos(const struct or_cs* h) { return h->npos; }
const uint32_t* or_cs_tab(const struct or_cs* h) { return h->tab; }
const int8_t* or_cs_rci(const struct or_cs* h) { return h->rci; }
const uint32_t* or_cs_pos(const struct or_cs* h) { return h->pos; }

/* ---- vote ---------------------------------------------------------------------------------- */

typedef struct {
  uint64_t key;
  float f, r;
  int used, listed;
} vote_t;

typedef struct {
  const struct or_cs* h;
  int read_len;
  vote_t* tab;
  uint32_t cap; /* power of two */
  uint32_t* order;
  int n_order;
  float max_hits, thresh, sens;
} search_t;

static void vote(search_t* s, uint64_t bin, int reverse) {
  uint32_t i = (uint32_t)((bin * 11400714819323198485ull) >> 32) & (s->cap - 1);
  while (s->tab[i].used && s->tab[i].key != bin) i = (i + 1) & (s->cap - 1);
  vote_t* e = &s->tab[i];
  float score;
  if (!e->used) {
    e->used = 1; e->key = bin; e->listed = 0;
    e->f = reverse ? 0.0f : 1.0f;
    e->r = reverse ? 1.0f : 0.0f;
    score = 1.0f;
  } else {
    score = reverse ? (e->r += 1.0f) : (e->f += 1.0f);
  }
  if (score > s->max_hits) {
    s->max_hits = score;
    s->thresh = s->max_hits * s->sens;
  }
  if (!e->listed && score >= s->thresh) {
    e->listed = 1;
    s->order[s->n_order++] = i;
  }
}

static void search_cb(uint64_t prefix, uint64_t pos, void* data) {
  search_t* s = (search_t*)data;
  const struct or_cs* h = s->h;
  /* GetRefEntry: forward list, then the reverse-complement k-mer's list */
  if (h->rci[prefix] != 0) {
    uint32_t start = h->tab[prefix] - 1, n = h->tab[prefix + 1] - 1 - start;
    for (uint32_t i = 0; i < n; ++i) vote(s, ((uint64_t)h->pos[start + i] - pos) >> h->bin_shift, 0);
  }
  uint64_t rc = rev_comp(prefix, h->k);
  if (h->rci[rc] != 0) {
    uint32_t start = h->tab[rc] - 1, n = h->tab[rc + 1] - 1 - start;
    uint64_t corr = (uint64_t)s->read_len - (pos + (uint64_t)h->k);
    for (uint32_t i = 0; i < n; ++i) vote(s, ((uint64_t)h->pos[start + i] - corr) >> h->bin_shift, 1);
  }
}

typedef struct { uint64_t hits; const struct or_cs* h; } count_t;
static void count_cb(uint64_t prefix, uint64_t pos, void* data) {
  count_t* c = (count_t*)data;
  (void)pos;
  if (c->h->rci[prefix] != 0) c->hits += c->h->tab[prefix + 1] - c->h->tab[prefix];
  uint64_t rc = rev_comp(prefix, c->h->k);
  if (c->h->rci[rc] != 0) c->hits += c->h->tab[rc + 1] - c->h->tab[rc];
}

/* One (sub-)read: RunRead without the hand-over to ScoreBuffer. Returns the candidate count;
 * outputs in the reference's emission order; *max_hits = maxHitNumber. */
int or_cs_search(const struct or_cs* h, const char* seq, int len, float sensitivity, float min_kmer_hits,
                 float* scores, uint64_t* locs, int* reverse, int cap, float* max_hits) {
  count_t c = {0, h};
  prefix_iteration(seq, (uint64_t)len, h->k, 0, 0, count_cb, &c);
  search_t s;
  memset(&s, 0, sizeof(s));
  s.h = h; s.read_len = len; s.sens = sensitivity;
  s.cap = 64;
  while (s.cap < 2 * c.hits + 2) s.cap <<= 1;
  s.tab = (vote_t*)calloc(s.cap, sizeof(vote_t));
  s.order = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(c.hits + 1));
  prefix_iteration(seq, (uint64_t)len, h->k, 0, 0, search_cb, &s);
  float thr = min_kmer_hits > s.thresh ? min_kmer_hits : s.thresh;
  int n = 0;
  const uint64_t half = h->bin_shift > 0 ? ((uint64_t)1 << (h->bin_shift - 1)) : 0;
  for (int i = 0; i < s.n_order; ++i) {
    const vote_t* e = &s.tab[s.order[i]];
    uint64_t loc = ((uint64_t)(uint32_t)e->key << h->bin_shift) + half; /* CSTableEntry::m_Location is a uint */
    if (e->f >= thr) { if (n < cap) { scores[n] = e->f; locs[n] = loc; reverse[n] = 0; } ++n; }
    if (e->r >= thr) { if (n < cap) { scores[n] = e->r; locs[n] = loc; reverse[n] = 1; } ++n; }
  }
  if (max_hits) *max_hits = s.max_hits;
  free(s.tab);
  free(s.order);
  return n;
}
