/* oracle/score_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * CPU restatement of the candidate selection ngmlr runs once every sub-read's candidates have their
 * Smith-Waterman scores: ScoreBuffer::topNSE (src/ScoreBuffer.cpp:170-192) and ScoreBuffer::computeMQ
 * (src/ScoreBuffer.cpp:33-45).
 *
 * topNSE orders the candidates with std::sort and the comparator a.Score.f > b.Score.f
 * (src/ScoreBuffer.cpp:25-27). std::sort is not stable and ties are the common case (scores are small
 * integers), so the resulting order is part of the behaviour. std::sort is a dependency that is not in
 * /root/reference: GNU libstdc++ (bits/stl_algo.h; the reference is built with the system g++, 13.3.0
 * here). Its published algorithm is restated below: introsort (median-of-three quicksort down to
 * partitions of <= 16 elements, heapsort once the depth budget 2*floor(log2 n) is spent) followed by
 * one final insertion sort. Pinned against the compiled reference in tests/test_oracle.py
 * (ScoreBuffer::topNSE called through oracle/ref_cs_shim.cpp). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

typedef struct {
  float score;
  int32_t idx;
} item_t;

static int before(const item_t* a, const item_t* b) { return a->score > b->score; }

static void swap_items(item_t* a, item_t* b) {
  item_t t = *a;
  *a = *b;
  *b = t;
}

/* std::__unguarded_linear_insert */
static void linear_insert(item_t* last) {
  item_t val = *last;
  item_t* next = last - 1;
  while (before(&val, next)) {
    *last = *next;
    last = next;
    --next;
  }
  *last = val;
}

/* std::__insertion_sort */
static void insertion_sort(item_t* first, item_t* last) {
  if (first == last) return;
  for (item_t* i = first + 1; i != last; ++i) {
    if (before(i, first)) {
      item_t val = *i;
      memmove(first + 1, first, (size_t)(i - first) * sizeof(item_t));
      *first = val;
    } else {
      linear_insert(i);
    }
  }
}

/* std::__adjust_heap followed by std::__push_heap */
static void adjust_heap(item_t* first, long hole, long len, item_t value) {
  const long top = hole;
  long child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (before(first + child, first + (child - 1))) --child;
    first[hole] = first[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    first[hole] = first[child - 1];
    hole = child - 1;
  }
  long parent = (hole - 1) / 2;
  while (hole > top && before(first + parent, &value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}

/* std::__partial_sort(first, last, last): make_heap + sort_heap */
static void heap_sort(item_t* first, item_t* last) {
  const long len = last - first;
  if (len >= 2) {
    long parent = (len - 2) / 2;
    for (;;) {
      item_t v = first[parent];
      adjust_heap(first, parent, len, v);
      if (parent == 0) break;
      --parent;
    }
  }
  while (last - first > 1) {
    --last;
    item_t v = *last;
    *last = *first;
    adjust_heap(first, 0, last - first, v);
  }
}

/* std::__move_median_to_first */
static void median_to_first(item_t* result, item_t* a, item_t* b, item_t* c) {
  if (before(a, b)) {
    if (before(b, c))
      swap_items(result, b);
    else if (before(a, c))
      swap_items(result, c);
    else
      swap_items(result, a);
  } else if (before(a, c)) {
    swap_items(result, a);
  } else if (before(b, c)) {
    swap_items(result, c);
  } else {
    swap_items(result, b);
  }
}

/* std::__unguarded_partition */
static item_t* partition(item_t* first, item_t* last, const item_t* pivot) {
  for (;;) {
    while (before(first, pivot)) ++first;
    --last;
    while (before(pivot, last)) --last;
    if (!(first < last)) return first;
    swap_items(first, last);
    ++first;
  }
}

/* std::__introsort_loop */
static void introsort_loop(item_t* first, item_t* last, long depth) {
  while (last - first > 16) {
    if (depth == 0) {
      heap_sort(first, last);
      return;
    }
    --depth;
    item_t* mid = first + (last - first) / 2;
    median_to_first(first, first + 1, mid, last - 1);
    item_t* cut = partition(first + 1, last, first);
    introsort_loop(cut, last, depth);
    last = cut;
  }
}

static void std_sort(item_t* first, item_t* last) {
  if (first == last) return;
  long n = last - first, lg = 0;
  while ((n >> (lg + 1)) != 0) ++lg;
  introsort_loop(first, last, 2 * lg);
  if (last - first > 16) {
    insertion_sort(first, first + 16);
    for (item_t* i = first + 16; i != last; ++i) linear_insert(i);
  } else {
    insertion_sort(first, last);
  }
}

/* ScoreBuffer::computeMQ(float, float) (src/ScoreBuffer.cpp:33-36); MAX_MQ = 60.0f (:16). */
int or_score_mq(float best, float second) {
  const float q = 60.0f * (best - second) / best;
  return (int)ceil((double)q);
}

/* ScoreBuffer::topNSE for one sub-read with n scored candidates. order[] receives, for every sorted
 * position, the index the candidate had on input. Returns the number of candidates kept for
 * alignment (MappedRead::Calculated after the call); *mq = MappedRead::mappingQlty. */
int or_score_select(const float* scores, int n, int32_t* order, int* mq) {
  item_t stack_items[64];
  item_t* it = stack_items;
  item_t* heap_items = 0;
  if (n > 64) {
    heap_items = (item_t*)malloc((size_t)n * sizeof(item_t));
    it = heap_items;
  }
  for (int i = 0; i < n; ++i) {
    it[i].score = scores[i];
    it[i].idx = i;
  }
  std_sort(it, it + n);
  int kept = n;
  if (n > 1) {
    const float min_score = it[0].score * 0.75f;
    int i = 1;
    while (i < n && it[i].score > min_score) ++i;
    kept = i;
  }
  *mq = 60;
  if (n > 1) *mq = or_score_mq(it[0].score, it[1].score);
  for (int i = 0; i < n; ++i) order[i] = it[i].idx;
  free(heap_items);
  return kept;
}
