// ngmlr_b200/csrc/cigar_text.cpp -- see cigar_text.h.
#include "cigar_text.h"

#include "device_types.h"

namespace nb {

namespace {

// Fast decimal append (the reference uses sprintf("%d"); the digits are the same).
inline void put_int(std::string& s, int v) {
  char buf[12];
  int n = 0;
  unsigned u = v < 0 ? 0u - (unsigned)v : (unsigned)v;
  do {
    buf[n++] = (char)('0' + u % 10);
    u /= 10;
  } while (u);
  if (v < 0) buf[n++] = '-';
  while (n) s.push_back(buf[--n]);
}

inline void put_op(std::string& s, int len, char op, int& count) {
  put_int(s, len);
  s.push_back(op);
  ++count;
}

// 32-event sliding window of "was this alignment column an error" used for the per-position
// mismatch density (inversion detection input, src/ConvexAlignFast.cpp:117-119, 193-269).
struct ErrorWindow {
  uint32_t bits = 0;
  int ones = 0;   // popcount(bits), kept incrementally (one bit leaves, one enters per event)
  int level = 0;  // what the reference records: NumberOfSetBits after a match/mismatch, but only
                  // "previous + 1" after the first base of an indel (it does not recount there)
  void shift(uint32_t in) {
    ones += (int)in - (int)(bits >> 31);
    bits = (bits << 1) | in;
  }
  void match() {
    shift(0u);
    level = ones;
  }
  void mismatch() {
    shift(1u);
    level = ones;
  }
  void gap_base(bool first) {  // only the first base of an indel run counts (maxIndelLength = 1)
    shift(first ? 1u : 0u);
    if (first) level = level + 1 > 0 ? level + 1 : 0;
  }
};

}  // namespace

bool binary_cigar_to_text(const int32_t* runs, int n_runs, const char* ref, int ref_len,
                          int ref_position, int ext_qstart, int ext_qend, AlignText& out) {
  // reuse the caller's buffers (capacity survives across batches)
  out.cigar.clear();
  out.md.clear();
  out.ret = -1;
  out.sv_type = 0;
  if (n_runs < 2) {
    out.nm_positions.clear();
    return false;
  }
  const char* aref = ref + ref_position;  // convertCigar receives refSeq + ref_position (:489)
  std::string& cg = out.cigar;
  std::string& md = out.md;
  cg.reserve((size_t)n_runs * 4 + 16);
  md.reserve((size_t)n_runs * 3 + 16);

  const int lead = runs[0] >> 4;
  const int trail = runs[n_runs - 1] >> 4;
  int ops = 0, covered = 0;
  out.qstart = lead + ext_qstart;
  if (out.qstart > 0) {
    put_op(cg, out.qstart, 'S', ops);
    covered += out.qstart;
  }
  int pos_ref = 0, pos_read = lead;
  out.first_ref = pos_ref;
  out.first_read = pos_read;

  int matches = 0, columns = 0, exact = 0;
  int pending_m = 0, md_run = 0, ri = 0;
  ErrorWindow win;
  // nmPerPosition records are written through a raw pointer into a buffer sized for every
  // alignment column (the hot loop of this stage: 12 bytes per column)
  {
    size_t cols = 0;
    for (int j = 1; j < n_runs - 1; ++j)
      if ((runs[j] & 15) != OP_I) cols += (size_t)(runs[j] >> 4);
    out.nm_positions.resize(cols * 3);  // not cleared first: only growth is zero-filled, every kept record is rewritten
  }
  int32_t* np = out.nm_positions.data();
  auto note = [&](int pr, int pq) {  // addPosition (:76-99)
    if (pq > 16 && pr > 16) {
      np[0] = pr - 16;
      np[1] = pq - 16;
      np[2] = win.level;
      np += 3;
    }
  };
  auto flush_m = [&]() {
    if (pending_m > 0) {
      put_op(cg, pending_m, 'M', ops);
      covered += pending_m;
      pending_m = 0;
    }
  };

  for (int j = 1; j < n_runs - 1; ++j) {
    const int op = runs[j] & 15, n = runs[j] >> 4;
    columns += n;
    exact += n;
    switch (op) {
      case OP_X:
        pending_m += n;
        for (int k = 0; k < n; ++k) {
          put_int(md, md_run);
          md_run = 0;
          md.push_back(aref[ri++]);
          win.mismatch();
          note(pos_ref, pos_read);
          ++pos_ref;
          ++pos_read;
        }
        break;
      case OP_EQ: {
        pending_m += n;
        md_run += n;
        matches += n;
        int k = 0;
        // while error bits are still sliding out of the 32-event window, or the positions have not
        // passed the 16-base margin, go column by column
        for (; k < n && (win.bits != 0u || pos_ref <= 16 || pos_read <= 16); ++k) {
          win.match();
          note(pos_ref, pos_read);
          ++pos_ref;
          ++pos_read;
        }
        // then the window is empty and stays empty: level 0 for the rest of the run
        win.level = k < n ? 0 : win.level;
        for (; k < n; ++k) {
          np[0] = pos_ref - 16;
          np[1] = pos_read - 16;
          np[2] = 0;
          np += 3;
          ++pos_ref;
          ++pos_read;
        }
        ri += n;
        break;
      }
      case OP_D:
        flush_m();
        put_op(cg, n, 'D', ops);
        put_int(md, md_run);
        md_run = 0;
        md.push_back('^');
        for (int k = 0; k < n; ++k) {
          md.push_back(aref[ri++]);
          win.gap_base(k == 0);
          note(pos_ref, pos_read);
          ++pos_ref;
        }
        break;
      case OP_I:
        flush_m();
        put_op(cg, n, 'I', ops);
        covered += n;
        for (int k = 0; k < n; ++k) win.gap_base(k == 0);
        pos_read += n;
        break;
      default:
        out.nm_positions.clear();
        return false;  // "Invalid cigar string" -> throw 1 (:272-274)
    }
  }
  out.nm_positions.resize((size_t)(np - out.nm_positions.data()));
  put_int(md, md_run);
  flush_m();
  out.qend = trail + ext_qend;
  if (out.qend > 0) put_op(cg, out.qend, 'S', ops);
  covered += out.qend;

  out.identity = (float)matches * 1.0f / (float)columns;
  out.nm = columns - matches;
  out.alignment_length = exact;
  out.last_ref = pos_ref;
  out.last_read = pos_read;
  out.cigar_op_count = ops;
  out.ret = covered;

  // Was the clipping caused by N in the reference? (:494-529). The reference probes for 'X',
  // which its own decoder never emits, so sv_type stays 0 on real data; kept for equality.
  out.sv_type = 0;
  {
    int n_count = 0, probes = 0;
    const int lo = ref_position - 100 > 0 ? ref_position - 100 : 0;
    for (int k = ref_position; k > lo; --k) {
      if (ref[k] == 'X') ++n_count;
      ++probes;
    }
    if ((float)n_count > (float)probes * 0.8f) out.sv_type |= 1;
    n_count = probes = 0;
    const int rest = ref_len - ref_position;
    const int hi = out.last_ref + 100 < rest ? out.last_ref + 100 : rest;
    for (int k = out.last_ref; k < hi; ++k) {
      if (ref[ref_position + k] == 'X') ++n_count;
      ++probes;
    }
    if ((float)n_count > (float)probes * 0.8f) out.sv_type |= 1;
  }
  return true;
}

int scan_low_identity_regions(const int32_t* nm_positions, int nm_count, int alignment_length,
                              std::vector<int32_t>& regions, int cap) {
  regions.clear();
  int closed = 0;
  int start_ref = -1, stop_ref = -1, start_read = -1, stop_read = -1;
  int distance = 20;  // maxDistance
  for (int i = 0; i < alignment_length; ++i) {
    const int nm = i < nm_count ? nm_positions[3 * i + 2] : 0;
    const int pr = i < nm_count ? nm_positions[3 * i + 0] : 0;
    const int pq = i < nm_count ? nm_positions[3 * i + 1] : 0;
    const float id = (float)(32 - nm) / 32.0f;
    const bool peak = id > 0.0f && id < 0.75f;  // isInversion (:1143-1148)
    if (start_ref == -1) {
      if (peak) {
        start_ref = stop_ref = pr;
        start_read = stop_read = pq;
      }
    } else if (peak) {
      stop_ref = pr;
      stop_read = pq;
      distance = 20;
    } else if (distance == 0) {
      if (closed < cap) {
        regions.push_back(start_ref);
        regions.push_back(stop_ref);
        regions.push_back(start_read);
        regions.push_back(stop_read);
      }
      ++closed;
      start_ref = stop_ref = start_read = stop_read = -1;
      distance = 20;
    } else {
      --distance;
    }
  }
  return closed;
}

}  // namespace nb
