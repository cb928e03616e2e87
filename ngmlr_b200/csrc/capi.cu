// ngmlr_b200/csrc/capi.cu -- host runtime + plain-C ABI (include/ngmlr_b200.h).
//
// Owns the device arenas (sequences, corridor rows, descriptors, direction arena, traceback
// strips), the pinned staging buffers and the stream; packs a batch of SingleAlign problems,
// launches fill -> traceback (which also compacts the binary CIGARs), and turns the binary CIGARs into the reference's `Align`
// fields. There is no CPU compute path: every entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <numeric>
#include <stdlib.h>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../include/ngmlr_b200.h"
#include "cigar_text.h"
#include "device_types.h"
#include "kernels.h"

namespace {

using namespace nb;

std::string g_create_error;

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  cudaError_t reserve(size_t n, bool keep = false, cudaStream_t st = 0) {
    if (n <= cap) return cudaSuccess;
    size_t want = std::max(n, cap + cap / 2);
    T* q = nullptr;
    cudaError_t e = cudaMalloc(&q, want * sizeof(T));
    if (e != cudaSuccess) return e;
    if (keep && p && cap) cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, st);
    if (p) {
      cudaStreamSynchronize(st);
      cudaFree(p);
    }
    p = q;
    cap = want;
    return cudaSuccess;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

template <typename T>
struct PinBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    size_t want = std::max(n, cap + cap / 2);
    T* q = nullptr;
    cudaError_t e = cudaMallocHost(&q, want * sizeof(T));
    if (e != cudaSuccess) return e;
    if (p) cudaFreeHost(p);
    p = q;
    cap = want;
    return cudaSuccess;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Host worker threads for packing and CIGAR/MD text (NGMLR_B200_HOST_THREADS overrides).
int host_threads() {
  static int n = [] {
    const char* e = getenv("NGMLR_B200_HOST_THREADS");
    int v = e ? atoi(e) : 0;
    if (v <= 0) v = (int)std::min(32u, std::max(1u, std::thread::hardware_concurrency()));
    return v;
  }();
  return n;
}

// fn(i) for i in [0, n), dynamically scheduled in chunks over host_threads() threads.
template <typename F>
void parallel_for(int n, int chunk, F fn) {
  const int threads = std::min(host_threads(), (n + chunk - 1) / chunk);
  if (threads <= 1) {
    for (int i = 0; i < n; ++i) fn(i);
    return;
  }
  std::atomic<int> next(0);
  auto work = [&]() {
    for (;;) {
      const int b = next.fetch_add(chunk);
      if (b >= n) break;
      const int e = std::min(n, b + chunk);
      for (int i = b; i < e; ++i) fn(i);
    }
  };
  std::vector<std::thread> pool;
  pool.reserve(threads - 1);
  for (int t = 1; t < threads; ++t) pool.emplace_back(work);
  work();
  for (auto& t : pool) t.join();
}

// bytes readable past the end of every staged sequence (the fill kernel stages whole 64-column
// chunks, see convex_fill.cu) and slack of the per-warp boundary strip
constexpr size_t SEQ_PAD = 192;
constexpr size_t STRIP_SLACK = 192;

// Scorings for which the as-coded SSE fill (raw indelRun in the run tests) can differ from the
// scalar rule: a gap-open out of a cell that was itself reached through the *other* gap type
// would have to tie with or beat the diagonal. Sufficient condition for equivalence, with a
// margin far above float rounding at alignment-score magnitudes:
//   open < 0 and max(open, ext_min) + open <= min(match, mismatch) - 0.25   (both gap kinds)
// Default scoring (2,-5,-5,-5,-1,0.15): -6 <= -5.25 -> scalar kernel. Otherwise the RAW kernel.
bool scoring_needs_raw(const Scoring& s) {
  const float sub_min = std::min(s.mat, s.mis);
  const bool ok = s.open_read < 0.0f && s.open_ref < 0.0f &&
                  std::max(s.open_ref, s.ext_min) + s.open_read <= sub_min - 0.25f &&
                  std::max(s.open_read, s.ext_min) + s.open_ref <= sub_min - 0.25f;
  return !ok;
}

}  // namespace

void nb_cs_release(ngmlr_b200_ctx* ctx);  // candidate-search state lives in a side table (below)

struct ngmlr_b200_ctx {
  int device = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  cudaEvent_t ev[6] = {};
  Scoring sc{};
  bool raw = false;
  int force_raw = -1;
  std::string error;

  // ---- convex batch state ----
  int n = 0;
  size_t seq_bytes = 0, rows = 0, nblocks = 0, tb_ints = 0;
  int max_len = 0;
  int max_ref_len = 0;
  int wide_problems = 0;  // problems whose corridor is >= 352 columns wide
  int force_team = -1;
  int fill_ctas_cap = 0;  // 0 = full occupancy
  PinBuf<unsigned long long> h_win;   // windows mode: start | arena offset | (sequenceLength, span) pairs
  DevBuf<unsigned long long> d_win;
  int64_t upload_d2h_bytes = 0;
  int ctas_per_sm[4] = {0, 0, 0, 0};  // occupancy of the four fill-kernel variants
  long long debug_arena_words = -1;    // test hook: initial direction-arena size
  PinBuf<uint8_t> h_seq;
  PinBuf<int32_t> h_coff, h_clen, h_order, h_blkbase;
  PinBuf<int8_t> h_delta;
  std::vector<uint8_t> is_packed;
  int no_corridor_packing = 0;  // NGMLR_B200_NO_CORRIDOR_PACKING=1: always ship raw CorridorLines
  PinBuf<AlnDesc> h_desc;
  PinBuf<FillOut> h_fill;
  PinBuf<TraceOut> h_trace;
  PinBuf<int32_t> h_runs;
  PinBuf<unsigned long long> h_counters;
  std::vector<int32_t> ext_qs, ext_qe;
  DevBuf<uint8_t> d_seq;
  DevBuf<int32_t> d_coff, d_clen, d_order, d_blkbase;
  DevBuf<int8_t> d_delta;
  DevBuf<AlnDesc> d_desc;
  DevBuf<BlockRec> d_blocks;
  DevBuf<uint32_t> d_dir;
  DevBuf<BndEntry> d_bnd;
  DevBuf<FillOut> d_fill;
  DevBuf<int32_t> d_scratch;
  DevBuf<TraceOut> d_trace;
  DevBuf<int32_t> d_runs;
  DevBuf<unsigned long long> d_counters;  // [0] dir_alloc, [1] runs_alloc, [2] work counter (as int)
  size_t dir_words_needed = 0;
  bool ran = false;
  unsigned long long runs_used = 0, dir_used = 0;
  int fill_grid = 0;
  ngmlr_b200_batch_stats stats{};
  std::vector<AlignText> texts;

  // ---- sw state ----
  PinBuf<uint8_t> h_sw_seq;
  PinBuf<uint64_t> h_sw_off;
  PinBuf<int32_t> h_sw_len;
  PinBuf<float> h_sw_out;
  DevBuf<uint8_t> d_sw_seq;
  DevBuf<uint64_t> d_sw_off;
  DevBuf<int32_t> d_sw_len;
  DevBuf<float> d_sw_out;
  DevBuf<int32_t> d_sw_scratch;

  int fail(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    error = buf;
    return -1;
  }
};

#define CU(call)                                                                          \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess)                                                               \
      return ctx->fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

extern "C" {

int ngmlr_b200_abi_version(void) { return NGMLR_B200_ABI_VERSION; }

int ngmlr_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

const char* ngmlr_b200_last_error(const ngmlr_b200_ctx* ctx) {
  return ctx ? ctx->error.c_str() : g_create_error.c_str();
}

int ngmlr_b200_create(int gpu_id, const ngmlr_b200_scoring* s, ngmlr_b200_ctx** out) {
  if (!out) return -1;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_create_error = std::string("ngmlr_b200: no CUDA device available (") + cudaGetErrorString(e) +
                     "); this library has no CPU fallback";
    return -1;
  }
  if (gpu_id < 0 || gpu_id >= count) {
    g_create_error = "ngmlr_b200: gpu_id out of range";
    return -1;
  }
  ngmlr_b200_ctx* ctx = new ngmlr_b200_ctx();
  ctx->device = gpu_id;
  if ((e = cudaSetDevice(gpu_id)) != cudaSuccess) {
    g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
    delete ctx;
    return -1;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, gpu_id);
  ctx->num_sms = prop.multiProcessorCount;
  if (prop.major < 10) {
    g_create_error = "ngmlr_b200: kernels are built for sm_100a only; device is sm_" +
                     std::to_string(prop.major) + std::to_string(prop.minor);
    delete ctx;
    return -1;
  }
  cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  for (auto& ev : ctx->ev) cudaEventCreate(&ev);
  ngmlr_b200_scoring d = {2.0f, -5.0f, -5.0f, -5.0f, -1.0f, 0.15f};
  if (s) d = *s;
  ctx->sc.mat = d.match;
  ctx->sc.mis = d.mismatch;
  ctx->sc.open_read = d.gap_open;  // gap_open_read = gap_open_ref = gapOpen (:39-40)
  ctx->sc.open_ref = d.gap_open;
  ctx->sc.gap_ext = d.gap_extend;
  ctx->sc.ext_min = d.gap_extend_min;
  ctx->sc.decay = d.gap_decay;
  ctx->raw = scoring_needs_raw(ctx->sc);
  if (const char* e = getenv("NGMLR_B200_FILL_TEAM")) ctx->force_team = atoi(e);
  if (const char* e = getenv("NGMLR_B200_FILL_CTAS_PER_SM")) ctx->fill_ctas_cap = std::max(0, atoi(e));
  if (const char* e = getenv("NGMLR_B200_NO_CORRIDOR_PACKING")) ctx->no_corridor_packing = atoi(e);
  *out = ctx;
  return 0;
}

void ngmlr_b200_destroy(ngmlr_b200_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  nb_cs_release(ctx);
  ctx->h_seq.release(); ctx->h_coff.release(); ctx->h_clen.release(); ctx->h_order.release(); ctx->h_blkbase.release(); ctx->h_delta.release();
  ctx->h_desc.release(); ctx->h_fill.release(); ctx->h_trace.release(); ctx->h_runs.release();
  ctx->h_counters.release();
  ctx->d_seq.release(); ctx->d_coff.release(); ctx->d_clen.release(); ctx->d_order.release(); ctx->d_blkbase.release(); ctx->d_delta.release(); ctx->d_win.release(); ctx->h_win.release();
  ctx->d_desc.release(); ctx->d_blocks.release(); ctx->d_dir.release(); ctx->d_bnd.release();
  ctx->d_fill.release(); ctx->d_scratch.release(); ctx->d_trace.release(); ctx->d_runs.release();
  ctx->d_counters.release();
  ctx->h_sw_seq.release(); ctx->h_sw_off.release(); ctx->h_sw_len.release(); ctx->h_sw_out.release();
  ctx->d_sw_seq.release(); ctx->d_sw_off.release(); ctx->d_sw_len.release(); ctx->d_sw_out.release();
  ctx->d_sw_scratch.release();
  for (auto& ev : ctx->ev) cudaEventDestroy(ev);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int ngmlr_b200_set_stream(ngmlr_b200_ctx* ctx, void* s) {
  if (!ctx) return -1;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  if (s) {
    ctx->stream = (cudaStream_t)s;
    ctx->own_stream = false;
  } else {
    cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    ctx->own_stream = true;
  }
  return 0;
}

void* ngmlr_b200_get_stream(ngmlr_b200_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

// Test hook: start convex_run with a direction arena of `words` 32-bit words (-1 = estimate).
int ngmlr_b200_debug_set_arena_words(ngmlr_b200_ctx* ctx, long long words) {
  if (!ctx) return -1;
  ctx->debug_arena_words = words;
  return 0;
}

// Cap the persistent fill grid at v CTAs per SM (0 = full occupancy, the default). With several
// contexts sharing a GPU a smaller grid per launch lets the launches of different contexts -- and
// their memory-bound candidate-search / traceback kernels -- reside on the SMs together.
// NGMLR_B200_FILL_CTAS_PER_SM sets the initial value. Tuning hook.
int ngmlr_b200_set_fill_ctas_per_sm(ngmlr_b200_ctx* ctx, int v) {
  if (!ctx) return -1;
  ctx->fill_ctas_cap = v > 0 ? v : 0;
  return 0;
}

// Test hook (host only, no context): the CIGAR/MD/nmPerPosition text stage on a given binary CIGAR.
// Outputs: ints[0..11] = ret, qstart, qend, nm, alignment_length, cigar_op_count, sv_type, first_ref,
// first_read, last_ref, last_read, nm_count (triples); *identity; cigar/md NUL-terminated (truncated
// to the caps); nm_out receives min(3 * nm_count, nm_cap) ints. Returns 1, or 0 where the reference throws.
int ngmlr_b200_debug_cigar_text(const int32_t* runs, int n_runs, const char* ref, int ref_len, int ref_position,
                                int ext_qstart, int ext_qend, int32_t* ints, float* identity, char* cigar,
                                int cigar_cap, char* md, int md_cap, int32_t* nm_out, int nm_cap) {
  AlignText t;
  t.nm_positions.assign(7, -1);  // stale content of a reused buffer must not leak into the result
  const bool ok = binary_cigar_to_text(runs, n_runs, ref, ref_len, ref_position, ext_qstart, ext_qend, t);
  const int v[12] = {t.ret, t.qstart, t.qend, t.nm, t.alignment_length, t.cigar_op_count, t.sv_type, t.first_ref,
                     t.first_read, t.last_ref, t.last_read, (int)(t.nm_positions.size() / 3)};
  memcpy(ints, v, sizeof(v));
  *identity = t.identity;
  snprintf(cigar, (size_t)cigar_cap, "%s", t.cigar.c_str());
  snprintf(md, (size_t)md_cap, "%s", t.md.c_str());
  const size_t n = std::min(t.nm_positions.size(), (size_t)std::max(nm_cap, 0));
  if (n) memcpy(nm_out, t.nm_positions.data(), n * sizeof(int32_t));
  return ok ? 1 : 0;
}

// force_team: -1 auto, 0 one warp per problem, 1 four-warp teams. Test / tuning hook.
int ngmlr_b200_set_force_team(ngmlr_b200_ctx* ctx, int v) {
  if (!ctx) return -1;
  ctx->force_team = v;
  return 0;
}

// force_raw: -1 auto (by scoring), 0 scalar-rule kernel, 1 as-coded (RAW) kernel. Test hook.
int ngmlr_b200_set_force_raw(ngmlr_b200_ctx* ctx, int v) {
  if (!ctx) return -1;
  ctx->force_raw = v;
  return 0;
}

}  // extern "C"

namespace {

// Reference windows decoded on the device instead of shipped as text (convex_upload_windows).
struct RefWindows {
  const uint8_t* d_enc;
  const unsigned long long* d_ref_starts;
  int n_starts;
  const uint64_t* win_start;  // host, n entries
};

// refs == nullptr <=> windows mode: the sequence arena then holds all reference windows first (filled
// by decode_windows_kernel and copied back for the host CIGAR/MD stage), then all reads.
int convex_upload_impl(ngmlr_b200_ctx* ctx, int n, const char* const* refs, const RefWindows* win,
                       const int32_t* ref_lens, const char* const* qrys, const int32_t* qry_lens,
                       const int32_t* corridor_offsets, const int32_t* corridor_lengths,
                       const int64_t* row_start, const int32_t* ext_qstart, const int32_t* ext_qend) {
  if (!ctx) return -1;
  if (n < 0) return ctx->fail("convex_upload: n < 0");
  CU(cudaSetDevice(ctx->device));
  ctx->ran = false;
  ctx->n = n;
  if (n == 0) return 0;
  // ---- sizes ----
  size_t seq_bytes = 0, rows = 0, nblocks = 0, tb_ints = 0;
  for (int i = 0; i < n; ++i) {
    if (ref_lens[i] < 0 || qry_lens[i] < 0) return ctx->fail("convex_upload: negative length at %d", i);
    if (row_start[i + 1] - row_start[i] != (int64_t)qry_lens[i])
      return ctx->fail("convex_upload: problem %d has %lld corridor rows for a %d-base read "
                       "(corridorHeight must equal qryLen)", i,
                       (long long)(row_start[i + 1] - row_start[i]), qry_lens[i]);
    seq_bytes += align_up((size_t)ref_lens[i] + SEQ_PAD, 16) + align_up((size_t)qry_lens[i] + SEQ_PAD, 16);
    rows += (size_t)qry_lens[i];
    nblocks += ((size_t)qry_lens[i] + 31) / 32;
  }
  CU(ctx->h_seq.reserve(seq_bytes + 64));
  CU(ctx->h_coff.reserve(rows + 32));
  CU(ctx->h_clen.reserve(rows + 32));
  CU(ctx->h_delta.reserve(rows + 32));
  CU(ctx->h_blkbase.reserve(nblocks + 1));
  ctx->is_packed.assign(n, 0);
  CU(ctx->h_desc.reserve(n));
  CU(ctx->h_order.reserve(n));
  ctx->ext_qs.assign(n, 0);
  ctx->ext_qe.assign(n, 0);
  // ---- pack (parallel over problems) ----
  const double t_pack0 = now_ms();
  const int64_t r0 = row_start[0];
  std::vector<size_t> ref_at(n), qry_at(n), blk_at(n), tb_at(n);
  size_t ref_region = 0;  // windows mode: bytes of the leading reference region
  if (!refs)
    for (int i = 0; i < n; ++i) ref_region += align_up((size_t)ref_lens[i] + SEQ_PAD, 16);
  {
    size_t so_ = 0, ro_ = 0, qo_ = ref_region, bo_ = 0, tb_ = 0;
    for (int i = 0; i < n; ++i) {
      const int rl = ref_lens[i], ql = qry_lens[i];
      const size_t rspan = align_up((size_t)rl + SEQ_PAD, 16), qspan = align_up((size_t)ql + SEQ_PAD, 16);
      if (refs) {
        ref_at[i] = so_;
        qry_at[i] = so_ + rspan;
        so_ += rspan + qspan;
      } else {
        ref_at[i] = ro_;
        qry_at[i] = qo_;
        ro_ += rspan;
        qo_ += qspan;
      }
      blk_at[i] = bo_;
      bo_ += ((size_t)ql + 31) / 32;
      tb_at[i] = tb_;
      const long long ref_cap = ql > 200000 ? (long long)ql + 1 : 200000;  // maxBinaryCigarLength (:480-485)
      tb_ += (size_t)std::min<long long>((long long)ql + rl + 4, ref_cap);
    }
    tb_ints = tb_;
  }
  std::vector<unsigned long long> est(n);
  std::vector<size_t> dirw(n);
  std::vector<int> maxlen(n);
  parallel_for(n, 16, [&](int i) {
    AlnDesc& d = ctx->h_desc.p[i];
    memset(&d, 0, sizeof(d));
    const int rl = ref_lens[i], ql = qry_lens[i];
    size_t so = ref_at[i];
    d.ref_off = so;
    if (refs) {
      memcpy(ctx->h_seq.p + so, refs[i], rl);
      memset(ctx->h_seq.p + so + rl, 0, align_up((size_t)rl + SEQ_PAD, 16) - rl);
    }
    so = qry_at[i];
    d.qry_off = so;
    memcpy(ctx->h_seq.p + so, qrys[i], ql);
    memset(ctx->h_seq.p + so + ql, 0, align_up((size_t)ql + SEQ_PAD, 16) - ql);
    d.row_off = (uint64_t)(row_start[i] - r0);
    d.blk_off = blk_at[i];
    d.ref_len = rl;
    d.height = ql;
    d.ref_cap = ql > 200000 ? ql + 1 : 200000;
    d.tb_cap = (int)std::min<long long>((long long)ql + rl + 4, d.ref_cap);
    d.tb_off = tb_at[i];
    // Corridor rows: one pass that writes the packed form (int8 offset deltas + one base per 32-row
    // block) and finds out whether it is exact for this problem (constant length, |delta| < 128);
    // only problems that fail ship their raw CorridorLines.
    const int32_t* src_off = corridor_offsets + row_start[i];
    const int32_t* src_len = corridor_lengths + row_start[i];
    int8_t* delta = ctx->h_delta.p + d.row_off;
    int32_t* blkbase = ctx->h_blkbase.p + d.blk_off;
    int ml = 0;
    unsigned long long sum = 0;
    bool packable = !ctx->no_corridor_packing && ql > 0;
    const int len0 = ql ? src_len[0] : 0;
    for (int y = 0; y < ql; ++y) {
      const int ln = src_len[y];
      ml = std::max(ml, ln);
      sum += (unsigned long long)std::max(ln, 0);
      const long long dl = y ? (long long)src_off[y] - (long long)src_off[y - 1] : 0;
      packable = packable && ln == len0 && dl >= -128 && dl <= 127;
      delta[y] = (int8_t)dl;
      if ((y & 31) == 0) blkbase[y >> 5] = src_off[y];
    }
    int32_t* lens = ctx->h_clen.p + d.row_off;
    int32_t* offs = ctx->h_coff.p + d.row_off;
    if (!packable && ql) {
      memcpy(offs, src_off, (size_t)ql * sizeof(int32_t));
      memcpy(lens, src_len, (size_t)ql * sizeof(int32_t));
    }
    d.packed = packable ? 1 : 0;
    d.const_len = len0;
    ctx->is_packed[i] = packable ? 1 : 0;
    offs = const_cast<int32_t*>(src_off);  // the arena estimate below only needs the end points
    d.max_len = ml;
    maxlen[i] = std::min(ml, rl);
    est[i] = sum;
    // direction arena estimate: per 32-row block, steps = row width + 31 (stagger) + corridor
    // advance over the block; the exact figure is computed by the kernel (bump allocation) and an
    // overflow triggers a re-run with a larger arena.
    dirw[i] = 0;
    if (ql > 0) {
      const long long adv_total = std::max<long long>(0, (long long)offs[ql - 1] - offs[0]);
      const long long adv = (adv_total * 32 + std::max(ql - 1, 1) - 1) / std::max(ql - 1, 1) + 2;
      const long long w = std::min<long long>(ml, (long long)rl);
      const long long steps = w + 31 + adv;
      dirw[i] = (size_t)(((size_t)ql + 31) / 32) * (size_t)((steps + 15) / 16 + 1) * 32;
    }
  });
  size_t so = 0, bo = 0;
  int max_len_all = 0;
  size_t dir_words = 0;
  ctx->max_ref_len = 0;
  ctx->wide_problems = 0;
  for (int i = 0; i < n; ++i) {
    if (maxlen[i] >= 352) ctx->wide_problems++;
    max_len_all = std::max(max_len_all, maxlen[i]);
    ctx->max_ref_len = std::max(ctx->max_ref_len, ref_lens[i]);
    dir_words += dirw[i];
    if (ext_qstart) ctx->ext_qs[i] = ext_qstart[i];
    if (ext_qend) ctx->ext_qe[i] = ext_qend[i];
  }
  so = seq_bytes;
  bo = nblocks;
  std::iota(ctx->h_order.p, ctx->h_order.p + n, 0);
  std::stable_sort(ctx->h_order.p, ctx->h_order.p + n,
                   [&](int a, int b) { return est[a] > est[b]; });
  ctx->seq_bytes = so;
  ctx->rows = rows;
  ctx->nblocks = bo;
  ctx->tb_ints = tb_ints;
  ctx->max_len = max_len_all;
  ctx->dir_words_needed = dir_words + dir_words / 16 + 1024;
  const double t_pack1 = now_ms();
  // ---- device arenas + H2D ----
  cudaStream_t st = ctx->stream;
  CU(ctx->d_seq.reserve(so + 64));
  CU(ctx->d_coff.reserve(rows + 32));
  CU(ctx->d_clen.reserve(rows + 32));
  CU(ctx->d_desc.reserve(n));
  CU(ctx->d_order.reserve(n));
  CU(ctx->d_blocks.reserve(bo + 1));
  CU(ctx->d_fill.reserve(n));
  CU(ctx->d_trace.reserve(n));
  CU(ctx->d_scratch.reserve(tb_ints + 32));
  CU(ctx->d_runs.reserve(tb_ints / 4 + 4096));
  CU(ctx->d_counters.reserve(4));
  CU(ctx->h_counters.reserve(4));
  CU(ctx->h_fill.reserve(n));
  CU(ctx->h_trace.reserve(n));
  CU(cudaMemcpyAsync(ctx->d_seq.p + ref_region, ctx->h_seq.p + ref_region, so - ref_region,
                     cudaMemcpyHostToDevice, st));
  if (!refs) {
    // window descriptors -> device, decode into the reference region, bring the text back for fetch()
    CU(ctx->h_win.reserve((size_t)n * 3 + 2));
    CU(ctx->d_win.reserve((size_t)n * 3 + 2));
    unsigned long long* hw = ctx->h_win.p;
    int32_t* hl = reinterpret_cast<int32_t*>(hw + 2 * (size_t)n);
    for (int i = 0; i < n; ++i) {
      hw[i] = win->win_start[i];
      hw[n + i] = ctx->h_desc.p[i].ref_off;
      hl[i] = ref_lens[i] + 1;                                               // sequenceLength incl. NUL
      hl[n + i] = (int32_t)align_up((size_t)ref_lens[i] + SEQ_PAD, 16);     // text + zero padding
    }
    CU(cudaMemcpyAsync(ctx->d_win.p, hw, (size_t)n * 24, cudaMemcpyHostToDevice, st));
    RefDecodeParams rp;
    rp.enc = win->d_enc;
    rp.ref_starts = win->d_ref_starts;
    rp.n_starts = win->n_starts;
    rp.n = n;
    rp.win_start = ctx->d_win.p;
    rp.out_off = reinterpret_cast<const uint64_t*>(ctx->d_win.p + n);
    rp.win_len = reinterpret_cast<const int32_t*>(ctx->d_win.p + 2 * (size_t)n);
    rp.out_span = rp.win_len + n;
    rp.out = ctx->d_seq.p;
    CU(launch_decode_windows(rp, st));
    CU(cudaMemcpyAsync(ctx->h_seq.p, ctx->d_seq.p, ref_region, cudaMemcpyDeviceToHost, st));
  }
  size_t raw_rows = 0;
  int raw_problems = 0;
  for (int i = 0; i < n; ++i)
    if (!ctx->is_packed[i]) {
      raw_rows += (size_t)qry_lens[i];
      ++raw_problems;
    }
  CU(ctx->d_delta.reserve(rows + 32));
  CU(ctx->d_blkbase.reserve(bo + 1));
  if (rows) {
    CU(cudaMemcpyAsync(ctx->d_delta.p, ctx->h_delta.p, rows, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->d_blkbase.p, ctx->h_blkbase.p, bo * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    if (raw_problems > 64 || raw_rows * 2 > rows) {  // many raw problems: ship the arrays whole
      CU(cudaMemcpyAsync(ctx->d_coff.p, ctx->h_coff.p, rows * sizeof(int32_t), cudaMemcpyHostToDevice, st));
      CU(cudaMemcpyAsync(ctx->d_clen.p, ctx->h_clen.p, rows * sizeof(int32_t), cudaMemcpyHostToDevice, st));
      raw_rows = rows;
    } else {
      for (int i = 0; i < n; ++i) {
        if (ctx->is_packed[i] || !qry_lens[i]) continue;
        const size_t ro = (size_t)ctx->h_desc.p[i].row_off, nb = (size_t)qry_lens[i] * sizeof(int32_t);
        CU(cudaMemcpyAsync(ctx->d_coff.p + ro, ctx->h_coff.p + ro, nb, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ctx->d_clen.p + ro, ctx->h_clen.p + ro, nb, cudaMemcpyHostToDevice, st));
      }
    }
  }
  CU(cudaMemcpyAsync(ctx->d_desc.p, ctx->h_desc.p, (size_t)n * sizeof(AlnDesc), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(ctx->d_order.p, ctx->h_order.p, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  CU(cudaStreamSynchronize(st));
  ctx->stats = ngmlr_b200_batch_stats();
  ctx->stats.host_pack_ms = (float)(t_pack1 - t_pack0);
  ctx->stats.host_h2d_ms = (float)(now_ms() - t_pack1);
  ctx->stats.host_threads = host_threads();
  ctx->stats.h2d_bytes = (int64_t)(so - ref_region + (refs ? 0 : (size_t)n * 24) + rows + bo * 4 + raw_rows * 8 +
                                   (size_t)n * (sizeof(AlnDesc) + 4));
  ctx->upload_d2h_bytes = (int64_t)ref_region;
  ctx->stats.seq_bytes = (int64_t)so;
  return 0;
}

}  // namespace

extern "C" {

int ngmlr_b200_convex_upload(ngmlr_b200_ctx* ctx, int n, const char* const* refs,
                             const int32_t* ref_lens, const char* const* qrys,
                             const int32_t* qry_lens, const int32_t* corridor_offsets,
                             const int32_t* corridor_lengths, const int64_t* row_start,
                             const int32_t* ext_qstart, const int32_t* ext_qend) {
  if (ctx && n > 0 && !refs) return ctx->fail("convex_upload: refs is NULL");
  return convex_upload_impl(ctx, n, refs, nullptr, ref_lens, qrys, qry_lens, corridor_offsets, corridor_lengths,
                            row_start, ext_qstart, ext_qend);
}

int ngmlr_b200_convex_run(ngmlr_b200_ctx* ctx) {
  if (!ctx) return -1;
  if (ctx->n == 0) {
    ctx->ran = true;
    return 0;
  }
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int n = ctx->n;
  const double t_run0 = now_ms();
  const bool raw = ctx->force_raw < 0 ? (ctx->raw || ctx->max_len > 32767) : (ctx->force_raw != 0);
  // Team mode (4 warps pipeline one problem) when the corridors are wide enough for the pipeline to
  // stay full (a warp must still be busy with its block when the fourth warp behind it has produced
  // the first chunk of the next one: ~4 x 100 columns); NGMLR_B200_FILL_TEAM=0/1 overrides.
  bool team = ctx->wide_problems * 2 > n;
  if (ctx->force_team >= 0) team = ctx->force_team != 0;
  const int variant = (raw ? 1 : 0) | (team ? 2 : 0);
  if (!ctx->ctas_per_sm[variant]) ctx->ctas_per_sm[variant] = std::max(1, fill_max_ctas_per_sm(raw, team));
  int per_sm = ctx->ctas_per_sm[variant];
  if (ctx->fill_ctas_cap > 0) per_sm = std::min(per_sm, ctx->fill_ctas_cap);
  const int max_grid = ctx->num_sms * per_sm;
  const int want_grid = team ? n : (n + FILL_WARPS_PER_CTA - 1) / FILL_WARPS_PER_CTA;
  const int grid = std::max(1, std::min(max_grid, want_grid));
  ctx->fill_grid = grid;
  const size_t warps = (size_t)grid * FILL_WARPS_PER_CTA;
  const size_t bnd_stride = align_up((size_t)ctx->max_ref_len + STRIP_SLACK, 8);
  CU(ctx->d_bnd.reserve(warps * bnd_stride));
  size_t dir_words = std::max(ctx->dir_words_needed, (size_t)4096);
  if (ctx->debug_arena_words >= 0) {  // force the overflow -> grow -> re-run path (tests)
    dir_words = (size_t)ctx->debug_arena_words;
    ctx->d_dir.release();
  }
  size_t runs_cap = ctx->d_runs.cap;

  for (int attempt = 0; attempt < 16; ++attempt) {
    CU(ctx->d_dir.reserve(dir_words));
    CU(ctx->d_runs.reserve(runs_cap));
    CU(cudaMemsetAsync(ctx->d_counters.p, 0, 4 * sizeof(unsigned long long), st));
    FillParams fp;
    fp.seq = ctx->d_seq.p;
    fp.c_off = ctx->d_coff.p;
    fp.c_len = ctx->d_clen.p;
    fp.c_blkbase = ctx->d_blkbase.p;
    fp.c_delta = ctx->d_delta.p;
    fp.desc = ctx->d_desc.p;
    fp.order = ctx->d_order.p;
    fp.n = n;
    fp.blocks = ctx->d_blocks.p;
    fp.dir = ctx->d_dir.p;
    fp.dir_capacity = ctx->d_dir.cap;
    fp.dir_alloc = ctx->d_counters.p + 0;
    fp.work_counter = reinterpret_cast<int*>(ctx->d_counters.p + 2);
    fp.bnd = ctx->d_bnd.p;
    fp.bnd_stride = bnd_stride;
    fp.out = ctx->d_fill.p;
    fp.sc = ctx->sc;
    TraceParams tp;
    tp.seq = ctx->d_seq.p;
    tp.c_off = ctx->d_coff.p;
    tp.c_len = ctx->d_clen.p;
    tp.c_blkbase = ctx->d_blkbase.p;
    tp.c_delta = ctx->d_delta.p;
    tp.desc = ctx->d_desc.p;
    tp.order = ctx->d_order.p;
    tp.n = n;
    tp.blocks = ctx->d_blocks.p;
    tp.dir = ctx->d_dir.p;
    tp.fill = ctx->d_fill.p;
    tp.scratch = ctx->d_scratch.p;
    tp.out = ctx->d_trace.p;
    tp.runs = ctx->d_runs.p;
    tp.runs_capacity = ctx->d_runs.cap;
    tp.runs_alloc = ctx->d_counters.p + 1;

    CU(cudaEventRecord(ctx->ev[0], st));
    CU(launch_convex_fill(fp, raw, team, grid, st));
    CU(cudaEventRecord(ctx->ev[1], st));
    CU(launch_convex_traceback(tp, st));
    CU(cudaEventRecord(ctx->ev[2], st));
    CU(cudaMemcpyAsync(ctx->h_counters.p, ctx->d_counters.p, 4 * sizeof(unsigned long long),
                       cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    ctx->stats.fill_launches++;
    ctx->stats.traceback_launches++;
    const unsigned long long dir_used = ctx->h_counters.p[0], runs_used = ctx->h_counters.p[1];
    bool again = false;
    if (dir_used > ctx->d_dir.cap) {
      // the counter under-reports after an overflow (warps stop allocating), so also double and
      // fall back to the host's estimate
      dir_words = std::max({(size_t)dir_used + (size_t)dir_used / 8 + 4096, (size_t)ctx->d_dir.cap * 2,
                            ctx->dir_words_needed});
      again = true;
    }
    if (runs_used > ctx->d_runs.cap) {
      runs_cap = (size_t)runs_used + 4096;
      again = true;
    }
    ctx->dir_used = dir_used;
    ctx->runs_used = runs_used;
    if (!again) break;
    if (attempt == 15) return ctx->fail("convex_run: arena sizing did not converge");
  }
  ctx->dir_words_needed = std::max(ctx->dir_words_needed, (size_t)ctx->dir_used);
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
  ctx->stats.fill_ms = ms;
  cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]);
  ctx->stats.traceback_ms = ms;
  ctx->stats.compact_ms = 0.0f;  // compaction is fused into the traceback kernel (fields kept for ABI stability)
  ctx->stats.dir_bytes = (int64_t)ctx->dir_used * 4;
  ctx->stats.cigar_runs = (int64_t)ctx->runs_used;
  ctx->stats.host_run_ms = (float)(now_ms() - t_run0);
  ctx->ran = true;
  return 0;
}

int ngmlr_b200_convex_fetch(ngmlr_b200_ctx* ctx, ngmlr_b200_align_result* results) {
  if (!ctx) return -1;
  if (!ctx->ran) return ctx->fail("convex_fetch: call convex_run first");
  const int n = ctx->n;
  if (n == 0) return 0;
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const double t_f0 = now_ms();
  CU(ctx->h_runs.reserve((size_t)ctx->runs_used + 16));
  CU(cudaMemcpyAsync(ctx->h_fill.p, ctx->d_fill.p, (size_t)n * sizeof(FillOut), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(ctx->h_trace.p, ctx->d_trace.p, (size_t)n * sizeof(TraceOut), cudaMemcpyDeviceToHost, st));
  if (ctx->runs_used)
    CU(cudaMemcpyAsync(ctx->h_runs.p, ctx->d_runs.p, (size_t)ctx->runs_used * sizeof(int32_t),
                       cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  const double t_f1 = now_ms();
  ctx->stats.d2h_bytes = (int64_t)((size_t)n * (sizeof(FillOut) + sizeof(TraceOut)) + ctx->runs_used * 4) +
                         ctx->upload_d2h_bytes;
  if ((int)ctx->texts.size() < n) ctx->texts.resize(n);
  for (int i = 0; i < n; ++i)
    if (ctx->h_trace.p[i].status == ST_DIR_OVERFLOW)
      return ctx->fail("convex_fetch: internal arena overflow survived run()");
  parallel_for(n, 8, [&](int i) {
    const AlnDesc& d = ctx->h_desc.p[i];
    const FillOut& f = ctx->h_fill.p[i];
    const TraceOut& t = ctx->h_trace.p[i];
    ngmlr_b200_align_result& r = results[i];
    memset(&r, 0, sizeof(r));
    r.ret = -1;
    r.score = -1.0f;  // align.Score = -1.0f on entry and on failure (:457, :537)
    r.cigar = "";
    r.md = "";
    r.cells = (int64_t)f.cells;
    if (t.status == ST_THROW) {
      r.threw = 1;
      return;
    }
    if (t.status != ST_OK) return;
    AlignText& tx = ctx->texts[i];
    const char* ref = reinterpret_cast<const char*>(ctx->h_seq.p + d.ref_off);
    if (!binary_cigar_to_text(ctx->h_runs.p + t.run_off, t.n_runs, ref, d.ref_len, t.ref_position,
                              ctx->ext_qs[i], ctx->ext_qe[i], tx)) {
      r.threw = 1;
      return;
    }
    r.ret = tx.ret;
    r.score = f.best_score;
    r.identity = tx.identity;
    r.position_offset = t.ref_position;
    r.qstart = tx.qstart;
    r.qend = tx.qend;
    r.nm = tx.nm;
    r.alignment_length = tx.alignment_length;
    r.cigar_op_count = tx.cigar_op_count;
    r.sv_type = tx.sv_type;
    r.first_ref = tx.first_ref;
    r.first_read = tx.first_read;
    r.last_ref = tx.last_ref;
    r.last_read = tx.last_read;
    r.nm_count = (int32_t)(tx.nm_positions.size() / 3);
    r.cigar_len = (int32_t)tx.cigar.size();
    r.md_len = (int32_t)tx.md.size();
    r.cigar = tx.cigar.c_str();
    r.md = tx.md.c_str();
    r.nm_positions = tx.nm_positions.data();
  });
  int64_t cells = 0, steps = 0;
  for (int i = 0; i < n; ++i) {
    cells += (int64_t)ctx->h_fill.p[i].cells;
    steps += ctx->h_trace.p[i].steps;
  }
  ctx->stats.cells = cells;
  ctx->stats.path_steps = steps;
  ctx->stats.host_d2h_ms = (float)(t_f1 - t_f0);
  ctx->stats.host_text_ms = (float)(now_ms() - t_f1);
  return 0;
}

int ngmlr_b200_convex_align_batch(ngmlr_b200_ctx* ctx, int n, const char* const* refs,
                                  const int32_t* ref_lens, const char* const* qrys,
                                  const int32_t* qry_lens, const int32_t* corridor_offsets,
                                  const int32_t* corridor_lengths, const int64_t* row_start,
                                  const int32_t* ext_qstart, const int32_t* ext_qend,
                                  ngmlr_b200_align_result* results) {
  int rc = ngmlr_b200_convex_upload(ctx, n, refs, ref_lens, qrys, qry_lens, corridor_offsets,
                                    corridor_lengths, row_start, ext_qstart, ext_qend);
  if (rc) return rc;
  rc = ngmlr_b200_convex_run(ctx);
  if (rc) return rc;
  return ngmlr_b200_convex_fetch(ctx, results);
}

int ngmlr_b200_convex_stats(ngmlr_b200_ctx* ctx, ngmlr_b200_batch_stats* out) {
  if (!ctx || !out) return -1;
  *out = ctx->stats;
  return 0;
}

int ngmlr_b200_convex_debug_directions(ngmlr_b200_ctx* ctx, int i, uint8_t* dirs, size_t dirs_cap,
                                       float* best_score, int32_t* best_ref, int32_t* best_read) {
  if (!ctx) return -1;
  if (!ctx->ran || i < 0 || i >= ctx->n) return ctx->fail("debug_directions: bad state/index");
  CU(cudaSetDevice(ctx->device));
  const AlnDesc& d = ctx->h_desc.p[i];
  const int H = d.height;
  const size_t nblk = ((size_t)H + 31) / 32;
  std::vector<BlockRec> blocks(nblk);
  FillOut f;
  CU(cudaMemcpy(&f, ctx->d_fill.p + i, sizeof(f), cudaMemcpyDeviceToHost));
  if (nblk) CU(cudaMemcpy(blocks.data(), ctx->d_blocks.p + d.blk_off, nblk * sizeof(BlockRec), cudaMemcpyDeviceToHost));
  std::vector<int32_t> offs_v(H), lens_v(H);
  for (int y = 0; y < H; ++y) {
    if (d.packed) {
      offs_v[y] = (y & 31) ? offs_v[y - 1] + ctx->h_delta.p[d.row_off + y] : ctx->h_blkbase.p[d.blk_off + (y >> 5)];
      lens_v[y] = d.const_len;
    } else {
      offs_v[y] = ctx->h_coff.p[d.row_off + y];
      lens_v[y] = ctx->h_clen.p[d.row_off + y];
    }
  }
  const int32_t* offs = offs_v.data();
  const int32_t* lens = lens_v.data();
  const char* ref = reinterpret_cast<const char*>(ctx->h_seq.p + d.ref_off);
  const char* qry = reinterpret_cast<const char*>(ctx->h_seq.p + d.qry_off);
  size_t total = 0;
  for (int y = 0; y < H; ++y) total += (size_t)std::max(lens[y], 0);
  if (total > dirs_cap) return ctx->fail("debug_directions: buffer too small (%zu > %zu)", total, dirs_cap);
  memset(dirs, 0xFF, total);
  std::vector<uint32_t> words;
  size_t row_base = 0;
  for (size_t b = 0; b < nblk; ++b) {
    const BlockRec& br = blocks[b];
    const size_t nw = (size_t)((br.nsteps + 15) / 16) * 32;
    words.resize(nw);
    if (nw) CU(cudaMemcpy(words.data(), ctx->d_dir.p + br.word_off, nw * 4, cudaMemcpyDeviceToHost));
    for (int t = 0; t < 32; ++t) {
      const int y = (int)b * 32 + t;
      if (y >= H) break;
      const int lo = std::max(offs[y], 0);
      const int hi = (int)std::min<long long>((long long)offs[y] + lens[y], d.ref_len);
      for (int x = lo; x < hi; ++x) {
        const int s = x - br.base + t;
        const uint32_t code = (words[(size_t)(s >> 4) * 32 + t] >> ((s & 15) * 2)) & 3u;
        uint8_t v = OP_STOP;
        if (code == DIR_DIAG) v = qry[y] == ref[x] ? OP_EQ : OP_X;
        else if (code == DIR_I) v = OP_I;
        else if (code == DIR_D) v = OP_D;
        dirs[row_base + (size_t)(x - offs[y])] = v;
      }
      row_base += (size_t)std::max(lens[y], 0);
    }
  }
  if (best_score) *best_score = f.best_score;
  if (best_ref) *best_ref = f.best_x;
  if (best_read) *best_read = f.best_y;
  return 0;
}

int ngmlr_b200_sw_score_batch(ngmlr_b200_ctx* ctx, int n, const char* const* refs,
                              const char* const* qrys, float* results) {
  if (!ctx) return -1;
  if (n <= 0) return 0;
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  CU(ctx->h_sw_off.reserve((size_t)2 * n));
  CU(ctx->h_sw_len.reserve((size_t)2 * n));
  CU(ctx->h_sw_out.reserve(n));
  size_t bytes = 0;
  int max_ref = 0, max_qry = 0;
  for (int i = 0; i < n; ++i) {
    const size_t rl = strlen(refs[i]) + 1, ql = strlen(qrys[i]) + 1;  // NUL included (:130-131)
    ctx->h_sw_off.p[i] = bytes;
    ctx->h_sw_len.p[i] = (int32_t)std::min<size_t>(rl, 1u << 30);
    bytes += align_up(rl, 16);
    ctx->h_sw_off.p[n + i] = bytes;
    ctx->h_sw_len.p[n + i] = (int32_t)std::min<size_t>(ql, 1u << 30);
    bytes += align_up(ql, 16);
    if (rl < 100000 && ql < 100000) {
      max_ref = std::max(max_ref, (int)rl);
      max_qry = std::max(max_qry, (int)ql);
    }
  }
  CU(ctx->h_sw_seq.reserve(bytes + 16));
  for (int i = 0; i < n; ++i) {
    memcpy(ctx->h_sw_seq.p + ctx->h_sw_off.p[i], refs[i], (size_t)ctx->h_sw_len.p[i]);
    memcpy(ctx->h_sw_seq.p + ctx->h_sw_off.p[n + i], qrys[i], (size_t)ctx->h_sw_len.p[n + i]);
  }
  CU(ctx->d_sw_seq.reserve(bytes + 16));
  CU(ctx->d_sw_off.reserve((size_t)2 * n));
  CU(ctx->d_sw_len.reserve((size_t)2 * n));
  CU(ctx->d_sw_out.reserve(n));
  const int warps_per_cta = 4;
  const int grid = std::max(1, std::min(ctx->num_sms * 8, (n + warps_per_cta - 1) / warps_per_cta));
  const size_t stride = max_qry > 288 ? align_up((size_t)max_ref + 2, 4) : 4;
  CU(ctx->d_sw_scratch.reserve((size_t)grid * warps_per_cta * stride * 2));
  CU(cudaMemcpyAsync(ctx->d_sw_seq.p, ctx->h_sw_seq.p, bytes, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(ctx->d_sw_off.p, ctx->h_sw_off.p, (size_t)2 * n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(ctx->d_sw_len.p, ctx->h_sw_len.p, (size_t)2 * n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  SwParams sp;
  sp.seq = ctx->d_sw_seq.p;
  sp.ref_off = ctx->d_sw_off.p;
  sp.qry_off = ctx->d_sw_off.p + n;
  sp.ref_len = ctx->d_sw_len.p;
  sp.qry_len = ctx->d_sw_len.p + n;
  sp.out = ctx->d_sw_out.p;
  sp.n = n;
  sp.scratch = ctx->d_sw_scratch.p;
  sp.scratch_stride = stride;
  CU(cudaEventRecord(ctx->ev[4], st));
  CU(launch_sw_score(sp, grid, st));
  CU(cudaEventRecord(ctx->ev[5], st));
  CU(cudaMemcpyAsync(ctx->h_sw_out.p, ctx->d_sw_out.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  memcpy(results, ctx->h_sw_out.p, (size_t)n * sizeof(float));
  return n;
}

float ngmlr_b200_sw_last_kernel_ms(ngmlr_b200_ctx* ctx) {
  float ms = 0;
  if (ctx) cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]);
  return ms;
}

}  // extern "C"

// ============================================================================================
// candidate search
// ============================================================================================
namespace {

struct CsState {
  DevBuf<uint8_t> d_packed;
  DevBuf<uint32_t> d_tab, d_pos, d_order;
  DevBuf<uint32_t> d_used;  // bitmap
  DevBuf<uint8_t> d_seq, d_tables;
  DevBuf<uint64_t> d_off;     // seq_off | table_off | order_off | out_off
  DevBuf<int32_t> d_len, d_count;
  DevBuf<uint32_t> d_cap;
  DevBuf<unsigned long long> d_hits;
  DevBuf<float> d_max;
  DevBuf<CsCandidate> d_out;
  uint32_t index_len = 0, n_pos = 0;
  uint64_t unit_offset = 0;
  int k = 0, bin_shift = 0;
  std::vector<float> scores;
  std::vector<uint64_t> locs;
  std::vector<uint8_t> reverse;
  std::vector<float> sw_scores;
  float last_ms = 0;
  // candidate scoring
  DevBuf<uint8_t> d_enc, d_rev;
  uint64_t enc_bytes = 0, concat_len = 0;
  DevBuf<unsigned long long> d_ref_starts;   // refStartPos (set_ref_starts)
  std::vector<unsigned long long> ref_starts;
  DevBuf<unsigned long long> d_winpos;
  DevBuf<uint64_t> d_qoff;
  DevBuf<int32_t> d_qlen;
  DevBuf<float> d_sw;
  DevBuf<int32_t> d_swscratch;
  std::vector<uint64_t> last_seq_off;  // arena offsets of the reads of the last search
  // resident pipeline
  int rn = 0;                       // reads uploaded
  size_t rbytes = 0;
  DevBuf<unsigned long long> d_a, d_b, d_c, d_sa, d_sb, d_sc, d_cnt64, d_cstart, d_cloc;
  DevBuf<uint8_t> d_scan_tmp;
  DevBuf<float> d_cscore;
  long long n_cand = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<int64_t> h_cstart;
  // pinned staging for the resident pipeline's upload / fetch
  PinBuf<uint8_t> p_seq, p_rev;
  PinBuf<float> p_score, p_sw;
  PinBuf<uint64_t> p_loc;
};

std::vector<std::pair<ngmlr_b200_ctx*, CsState*>> g_cs_states;
std::mutex* g_cs_mutex = new std::mutex();

CsState* cs_state(ngmlr_b200_ctx* ctx, bool create) {
  std::lock_guard<std::mutex> lock(*g_cs_mutex);
  for (auto& kv : g_cs_states)
    if (kv.first == ctx) return kv.second;
  if (!create) return nullptr;
  g_cs_states.emplace_back(ctx, new CsState());
  return g_cs_states.back().second;
}

}  // namespace

void nb_cs_release(ngmlr_b200_ctx* ctx) {
  std::lock_guard<std::mutex> lock(*g_cs_mutex);
  for (size_t i = 0; i < g_cs_states.size(); ++i) {
    if (g_cs_states[i].first != ctx) continue;
    CsState* cs = g_cs_states[i].second;
    cs->d_packed.release(); cs->d_tab.release(); cs->d_pos.release(); cs->d_order.release(); cs->d_ref_starts.release();
    cs->d_used.release(); cs->d_seq.release(); cs->d_tables.release(); cs->d_off.release();
    cs->d_len.release(); cs->d_count.release(); cs->d_cap.release(); cs->d_hits.release();
    cs->d_max.release(); cs->d_out.release(); cs->d_enc.release(); cs->d_rev.release();
    cs->d_winpos.release(); cs->d_qoff.release(); cs->d_qlen.release(); cs->d_sw.release();
    cs->d_swscratch.release(); cs->d_a.release(); cs->d_b.release(); cs->d_c.release();
    cs->d_sa.release(); cs->d_sb.release(); cs->d_sc.release(); cs->d_cnt64.release();
    cs->d_cstart.release(); cs->d_cloc.release(); cs->d_scan_tmp.release(); cs->d_cscore.release();
    if (cs->ev0) cudaEventDestroy(cs->ev0);
    if (cs->ev1) cudaEventDestroy(cs->ev1);
    cs->p_seq.release(); cs->p_rev.release(); cs->p_score.release(); cs->p_sw.release(); cs->p_loc.release();
    delete cs;
    g_cs_states.erase(g_cs_states.begin() + i);
    return;
  }
}

extern "C" {

int ngmlr_b200_cs_set_index(ngmlr_b200_ctx* ctx, const void* packed_index, uint32_t index_len,
                            const uint32_t* positions, uint32_t n_positions, uint64_t unit_offset,
                            int k, int bin_shift) {
  if (!ctx) return -1;
  if (k < 1 || k > 16) return ctx->fail("cs_set_index: k must be in 1..16");
  if (index_len != (1u << (2 * k)) + 1u) return ctx->fail("cs_set_index: index_len must be 4^k + 1");
  CU(cudaSetDevice(ctx->device));
  CsState* cs = cs_state(ctx, true);
  cudaStream_t st = ctx->stream;
  CU(cs->d_packed.reserve((size_t)index_len * 5));
  CU(cs->d_tab.reserve((size_t)index_len + 1));
  CU(cs->d_used.reserve(((size_t)index_len + 255) / 256 * 8 + 8));  // one word per 32 prefixes, whole blocks
  CU(cs->d_pos.reserve((size_t)n_positions + 1));
  CU(cudaMemcpyAsync(cs->d_packed.p, packed_index, (size_t)index_len * 5, cudaMemcpyHostToDevice, st));
  if (n_positions)
    CU(cudaMemcpyAsync(cs->d_pos.p, positions, (size_t)n_positions * 4, cudaMemcpyHostToDevice, st));
  CU(launch_unpack_index(cs->d_packed.p, index_len, cs->d_tab.p, cs->d_used.p, st));
  CU(cudaStreamSynchronize(st));
  cs->d_packed.release();
  cs->index_len = index_len;
  cs->n_pos = n_positions;
  cs->unit_offset = unit_offset;
  cs->k = k;
  cs->bin_shift = bin_shift;
  return 0;
}

int ngmlr_b200_cs_search_batch(ngmlr_b200_ctx* ctx, int n, const char* const* seqs,
                               const int32_t* lens, float sensitivity, float min_kmer_hits,
                               int64_t* cand_start, const float** scores, const uint64_t** locs,
                               const uint8_t** reverse, float* max_hits) {
  if (!ctx) return -1;
  CsState* cs = cs_state(ctx, false);
  if (!cs || !cs->index_len) return ctx->fail("cs_search_batch: call cs_set_index first");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  cs->scores.clear();
  cs->locs.clear();
  cs->reverse.clear();
  cand_start[0] = 0;
  if (n <= 0) return 0;
  // ---- reads -> device ----
  std::vector<uint64_t> seq_off(n);
  size_t bytes = 0;
  for (int i = 0; i < n; ++i) {
    seq_off[i] = bytes;
    bytes += align_up((size_t)std::max(lens[i], 0) + 1, 16);
  }
  cs->last_seq_off = seq_off;
  std::vector<uint8_t> hseq(bytes + 16, 0);
  parallel_for(n, 256, [&](int i) { memcpy(hseq.data() + seq_off[i], seqs[i], (size_t)std::max(lens[i], 0)); });
  CU(cs->d_seq.reserve(bytes + 16));
  CU(cs->d_off.reserve((size_t)4 * n));
  CU(cs->d_len.reserve(n));
  CU(cs->d_hits.reserve(n));
  CU(cs->d_cap.reserve(n));
  CU(cs->d_count.reserve(n));
  CU(cs->d_max.reserve(n));
  CU(cudaMemcpyAsync(cs->d_seq.p, hseq.data(), bytes, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(cs->d_off.p, seq_off.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(cs->d_len.p, lens, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  CsParams p;
  memset(&p, 0, sizeof(p));
  p.tab = cs->d_tab.p;
  p.used_bits = cs->d_used.p;
  p.pos = cs->d_pos.p;
  p.unit_offset = cs->unit_offset;
  p.k = cs->k;
  p.bin_shift = cs->bin_shift;
  p.sensitivity = sensitivity;
  p.min_kmer_hits = min_kmer_hits;
  p.seq = cs->d_seq.p;
  p.seq_off = cs->d_off.p;
  p.seq_len = cs->d_len.p;
  p.n = n;
  p.hits = cs->d_hits.p;
  // ---- pass 1: hits per read ----
  CU(cudaEventRecord(ctx->ev[4], st));
  CU(launch_cs_search(p, true, st));
  std::vector<unsigned long long> hits(n);
  CU(cudaMemcpyAsync(hits.data(), cs->d_hits.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  // ---- pass 2 in chunks bounded by a device-memory budget ----
  const size_t budget = (size_t)4 << 30;
  std::vector<uint64_t> toff(n), ooff(n), roff(n);
  std::vector<uint32_t> caps(n);
  std::vector<int32_t> counts(n);
  std::vector<CsCandidate> hout;
  int first = 0;
  while (first < n) {
    size_t tent = 0, oent = 0, rent = 0;
    int last = first;
    while (last < n) {
      uint32_t cap = 16;
      while ((unsigned long long)cap < 2 * hits[last] + 2) cap <<= 1;
      const size_t need = (size_t)cap * 16 + (size_t)hits[last] * 4 + (size_t)hits[last] * 2 * 16;
      if (last > first && (tent * 16 + oent * 4 + rent * 16 + need) > budget) break;
      caps[last] = cap;
      toff[last] = tent;
      ooff[last] = oent;
      roff[last] = rent;
      tent += cap;
      oent += (size_t)hits[last];
      rent += (size_t)hits[last] * 2;
      ++last;
    }
    const int m = last - first;
    CU(cs->d_tables.reserve(tent * 16 + 16));
    CU(cs->d_order.reserve(oent + 4));
    CU(cs->d_out.reserve(rent + 4));
    CU(cudaMemsetAsync(cs->d_tables.p, 0, tent * 16, st));
    CU(cudaMemcpyAsync(cs->d_off.p + (size_t)n, toff.data() + first, (size_t)m * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(cs->d_off.p + (size_t)2 * n, ooff.data() + first, (size_t)m * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(cs->d_off.p + (size_t)3 * n, roff.data() + first, (size_t)m * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(cs->d_cap.p, caps.data() + first, (size_t)m * 4, cudaMemcpyHostToDevice, st));
    CsParams q = p;
    q.seq_off = cs->d_off.p + first;
    q.seq_len = cs->d_len.p + first;
    q.n = m;
    q.tables = cs->d_tables.p;
    q.table_off = cs->d_off.p + (size_t)n;
    q.table_cap = cs->d_cap.p;
    q.order = cs->d_order.p;
    q.order_off = cs->d_off.p + (size_t)2 * n;
    q.out = cs->d_out.p;
    q.out_off = cs->d_off.p + (size_t)3 * n;
    q.out_count = cs->d_count.p;
    q.max_hits = cs->d_max.p;
    CU(launch_cs_search(q, false, st));
    hout.resize(rent + 1);
    CU(cudaMemcpyAsync(counts.data() + first, cs->d_count.p, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(max_hits + first, cs->d_max.p, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
    if (rent) CU(cudaMemcpyAsync(hout.data(), cs->d_out.p, rent * sizeof(CsCandidate), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int i = first; i < last; ++i) {
      const CsCandidate* c = hout.data() + roff[i];
      for (int j = 0; j < counts[i]; ++j) {
        cs->scores.push_back(c[j].score);
        cs->locs.push_back(c[j].loc);
        cs->reverse.push_back((uint8_t)c[j].reverse);
      }
      cand_start[i + 1] = (int64_t)cs->scores.size();
    }
    first = last;
  }
  CU(cudaEventRecord(ctx->ev[5], st));
  CU(cudaStreamSynchronize(st));
  *scores = cs->scores.data();
  *locs = cs->locs.data();
  *reverse = cs->reverse.data();
  return n;
}

}  // extern "C"

extern "C" {

int ngmlr_b200_cs_set_reference(ngmlr_b200_ctx* ctx, const uint8_t* bin_ref, uint64_t n_bytes,
                                uint64_t concat_len) {
  if (!ctx) return -1;
  if (concat_len > 2 * n_bytes) return ctx->fail("cs_set_reference: concat_len exceeds 2 * n_bytes");
  CU(cudaSetDevice(ctx->device));
  CsState* cs = cs_state(ctx, true);
  CU(cs->d_enc.reserve(n_bytes + 64));
  CU(cudaMemsetAsync(cs->d_enc.p, 0x44, n_bytes + 64, ctx->stream));  // 'N','N' past the end
  CU(cudaMemcpyAsync(cs->d_enc.p, bin_ref, n_bytes, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  cs->enc_bytes = n_bytes;
  cs->concat_len = concat_len;
  return 0;
}

int ngmlr_b200_set_ref_starts(ngmlr_b200_ctx* ctx, const uint64_t* ref_start_pos, int n_entries) {
  if (!ctx) return -1;
  if (n_entries < 2 || !ref_start_pos) return ctx->fail("set_ref_starts: need the contig starts plus the end entry");
  for (int i = 1; i < n_entries; ++i)
    if (ref_start_pos[i] <= ref_start_pos[i - 1]) return ctx->fail("set_ref_starts: entries must increase");
  CU(cudaSetDevice(ctx->device));
  CsState* cs = cs_state(ctx, true);
  cs->ref_starts.assign(ref_start_pos, ref_start_pos + n_entries);
  CU(cs->d_ref_starts.reserve((size_t)n_entries));
  CU(cudaMemcpyAsync(cs->d_ref_starts.p, cs->ref_starts.data(), (size_t)n_entries * 8, cudaMemcpyHostToDevice,
                     ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

namespace {

// The contract of the window calls: the window starts inside a contig or inside the 1000-N spacer in
// front of one (where the reference itself is well defined).
int check_windows(ngmlr_b200_ctx* ctx, const CsState* cs, int n, const uint64_t* start, const char* who) {
  if (!cs || !cs->enc_bytes || cs->ref_starts.empty())
    return ctx->fail("%s: call cs_set_reference and set_ref_starts first", who);
  const auto& rs = cs->ref_starts;
  for (int i = 0; i < n; ++i) {
    const uint64_t p = start[i];
    if (p >= cs->concat_len || p >= rs.back() || p == 0)
      return ctx->fail("%s: window %d starts outside the reference", who, i);
    size_t u = std::upper_bound(rs.begin(), rs.end(), (unsigned long long)p) - rs.begin();
    if (rs[u] - p < 1000ull) ++u;
    if (u >= rs.size() || p > rs[u] - 1000ull)
      return ctx->fail("%s: window %d starts behind the end of a contig", who, i);
  }
  return 0;
}

}  // namespace

int ngmlr_b200_decode_windows(ngmlr_b200_ctx* ctx, int n, const uint64_t* start, const int32_t* seq_len,
                              char* out, const int64_t* out_off) {
  if (!ctx) return -1;
  if (n <= 0) return 0;
  CsState* cs = cs_state(ctx, false);
  if (check_windows(ctx, cs, n, start, "decode_windows")) return -1;
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  size_t total = 0;
  for (int i = 0; i < n; ++i) {
    if (seq_len[i] < 1) return ctx->fail("decode_windows: sequence length %d < 1", seq_len[i]);
    total += align_up((size_t)seq_len[i], 16);
  }
  CU(ctx->h_win.reserve((size_t)n * 3 + 2));
  CU(ctx->d_win.reserve((size_t)n * 3 + 2));
  CU(ctx->d_sw_seq.reserve(total + 64));
  CU(ctx->h_sw_seq.reserve(total + 64));
  unsigned long long* hw = ctx->h_win.p;
  int32_t* hl = reinterpret_cast<int32_t*>(hw + 2 * (size_t)n);
  size_t at = 0;
  for (int i = 0; i < n; ++i) {
    hw[i] = start[i];
    hw[n + i] = at;
    hl[i] = seq_len[i];
    hl[n + i] = seq_len[i];
    at += align_up((size_t)seq_len[i], 16);
  }
  CU(cudaMemcpyAsync(ctx->d_win.p, hw, (size_t)n * 24, cudaMemcpyHostToDevice, st));
  RefDecodeParams rp;
  rp.enc = cs->d_enc.p;
  rp.ref_starts = cs->d_ref_starts.p;
  rp.n_starts = (int)cs->ref_starts.size();
  rp.n = n;
  rp.win_start = ctx->d_win.p;
  rp.out_off = reinterpret_cast<const uint64_t*>(ctx->d_win.p + n);
  rp.win_len = reinterpret_cast<const int32_t*>(ctx->d_win.p + 2 * (size_t)n);
  rp.out_span = rp.win_len + n;
  rp.out = ctx->d_sw_seq.p;
  CU(launch_decode_windows(rp, st));
  CU(cudaMemcpyAsync(ctx->h_sw_seq.p, ctx->d_sw_seq.p, total, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  parallel_for(n, 64, [&](int i) { memcpy(out + out_off[i], ctx->h_sw_seq.p + hw[n + i], (size_t)seq_len[i]); });
  return n;
}

int ngmlr_b200_convex_upload_windows(ngmlr_b200_ctx* ctx, int n, const uint64_t* on_ref_start,
                                     const uint64_t* on_ref_stop, const char* const* qrys,
                                     const int32_t* qry_lens, const int32_t* corridor_offsets,
                                     const int32_t* corridor_lengths, const int64_t* row_start,
                                     const int32_t* ext_qstart, const int32_t* ext_qend) {
  if (!ctx) return -1;
  if (n < 0) return ctx->fail("convex_upload_windows: n < 0");
  CsState* cs = cs_state(ctx, false);
  if (n > 0 && check_windows(ctx, cs, n, on_ref_start, "convex_upload_windows")) return -1;
  std::vector<int32_t> ref_lens((size_t)std::max(n, 1));
  for (int i = 0; i < n; ++i) {
    // extractReferenceSequenceForAlignment: onRefStart >= onRefStop -> no sequence (src/AlignmentBuffer.cpp:204-207)
    if (on_ref_start[i] >= on_ref_stop[i] || on_ref_stop[i] - on_ref_start[i] > 0x7ffffff0ull)
      return ctx->fail("convex_upload_windows: window %d is empty or too long", i);
    ref_lens[i] = (int32_t)(on_ref_stop[i] - on_ref_start[i]);  // strlen of the decoded refSeqLength = stop-start+1 buffer
  }
  RefWindows w;
  w.d_enc = cs ? cs->d_enc.p : nullptr;
  w.d_ref_starts = cs ? cs->d_ref_starts.p : nullptr;
  w.n_starts = cs ? (int)cs->ref_starts.size() : 0;
  w.win_start = on_ref_start;
  return convex_upload_impl(ctx, n, nullptr, &w, ref_lens.data(), qrys, qry_lens, corridor_offsets,
                            corridor_lengths, row_start, ext_qstart, ext_qend);
}

int ngmlr_b200_cs_score_batch(ngmlr_b200_ctx* ctx, int n, const char* const* seqs,
                              const int32_t* lens, float sensitivity, float min_kmer_hits,
                              int corridor, int read_part_length, int64_t* cand_start,
                              const float** cs_scores, const uint64_t** locs, const uint8_t** reverse,
                              const float** sw_scores, float* max_hits) {
  if (!ctx) return -1;
  CsState* cs = cs_state(ctx, false);
  if (!cs || !cs->enc_bytes) return ctx->fail("cs_score_batch: call cs_set_reference first");
  int rc = ngmlr_b200_cs_search_batch(ctx, n, seqs, lens, sensitivity, min_kmer_hits, cand_start,
                                      cs_scores, locs, reverse, max_hits);
  if (rc < 0) return rc;
  cs->sw_scores.clear();
  *sw_scores = cs->sw_scores.data();
  if (n <= 0) return rc;
  const size_t m = (size_t)cand_start[n];
  cs->sw_scores.assign(m, -1.0f);
  *sw_scores = cs->sw_scores.data();
  if (!m) return rc;
  cudaStream_t st = ctx->stream;
  // per candidate: window position, strand, and the arena location of its read (already on device)
  std::vector<unsigned long long> win(m);
  std::vector<uint64_t> qoff(m);
  std::vector<int32_t> qlen(m);
  int max_q = 0;
  for (int i = 0; i < n; ++i) {
    for (int64_t j = cand_start[i]; j < cand_start[i + 1]; ++j) {
      win[j] = (unsigned long long)cs->locs[j] - (unsigned long long)(corridor >> 1);  // uloc arithmetic (:110)
      qoff[j] = cs->last_seq_off[i];
      qlen[j] = lens[i] + 1;  // strlen + 1
    }
    max_q = std::max(max_q, lens[i] + 1);
  }
  const int win_len = ((read_part_length + 10 + corridor) | 1) + 1;  // refMaxLen, src/ScoreBuffer.h:71-72
  CU(cs->d_winpos.reserve(m));
  CU(cs->d_qoff.reserve(m));
  CU(cs->d_qlen.reserve(m));
  CU(cs->d_rev.reserve(m));
  CU(cs->d_sw.reserve(m));
  CU(cudaMemcpyAsync(cs->d_winpos.p, win.data(), m * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(cs->d_qoff.p, qoff.data(), m * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(cs->d_qlen.p, qlen.data(), m * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(cs->d_rev.p, cs->reverse.data(), m, cudaMemcpyHostToDevice, st));
  const int warps_per_cta = 4;
  const int grid = (int)std::max<size_t>(1, std::min<size_t>((size_t)ctx->num_sms * 8, (m + warps_per_cta - 1) / warps_per_cta));
  const size_t stride = max_q > 288 ? align_up((size_t)win_len + 4, 4) : 4;
  CU(cs->d_swscratch.reserve((size_t)grid * warps_per_cta * stride * 2));
  SwParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.seq = cs->d_seq.p;
  sp.ref_off = cs->d_qoff.p;  // unused in gather mode
  sp.qry_off = cs->d_qoff.p;
  sp.ref_len = cs->d_qlen.p;  // unused in gather mode
  sp.qry_len = cs->d_qlen.p;
  sp.out = cs->d_sw.p;
  sp.n = (int)m;
  sp.scratch = cs->d_swscratch.p;
  sp.scratch_stride = stride;
  sp.enc = cs->d_enc.p;
  sp.concat_len = cs->concat_len;
  sp.win_pos = cs->d_winpos.p;
  sp.rev = cs->d_rev.p;
  sp.win_len = win_len;
  CU(launch_sw_score_gather(sp, grid, st));
  CU(cudaMemcpyAsync(cs->sw_scores.data(), cs->d_sw.p, m * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return rc;
}

}  // extern "C"

extern "C" {

int ngmlr_b200_cs_upload(ngmlr_b200_ctx* ctx, int n, const char* const* seqs, const int32_t* lens) {
  if (!ctx) return -1;
  if (n < 0) return ctx->fail("cs_upload: n < 0");
  CU(cudaSetDevice(ctx->device));
  CsState* cs = cs_state(ctx, true);
  cudaStream_t st = ctx->stream;
  std::vector<uint64_t> seq_off((size_t)n + 1);
  size_t bytes = 0;
  for (int i = 0; i < n; ++i) {
    seq_off[i] = bytes;
    bytes += align_up((size_t)std::max(lens[i], 0) + 1, 16);
  }
  CU(cs->p_seq.reserve(bytes + 16));
  uint8_t* hseq = cs->p_seq.p;
  parallel_for(n, 256, [&](int i) {
    const size_t L = (size_t)std::max(lens[i], 0);
    memcpy(hseq + seq_off[i], seqs[i], L);
    memset(hseq + seq_off[i] + L, 0, align_up(L + 1, 16) - L);
  });
  CU(cs->d_seq.reserve(bytes + 16));
  CU(cs->d_off.reserve((size_t)n + 1));
  CU(cs->d_len.reserve((size_t)n + 1));
  CU(cudaMemcpyAsync(cs->d_seq.p, hseq, bytes, cudaMemcpyHostToDevice, st));
  if (n) {
    CU(cudaMemcpyAsync(cs->d_off.p, seq_off.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(cs->d_len.p, lens, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  }
  CU(cudaStreamSynchronize(st));
  cs->rn = n;
  cs->rbytes = bytes;
  cs->n_cand = 0;
  return 0;
}

int ngmlr_b200_cs_run(ngmlr_b200_ctx* ctx, float sensitivity, float min_kmer_hits, int corridor,
                      int read_part_length, int64_t* n_candidates, float* kernel_ms) {
  if (!ctx) return -1;
  CsState* cs = cs_state(ctx, false);
  if (!cs || !cs->index_len) return ctx->fail("cs_run: call cs_set_index first");
  if (!cs->enc_bytes) return ctx->fail("cs_run: call cs_set_reference first");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int n = cs->rn;
  if (n_candidates) *n_candidates = 0;
  if (kernel_ms) *kernel_ms = 0.0f;
  if (n <= 0) return 0;
  if (!cs->ev0) {
    CU(cudaEventCreate(&cs->ev0));
    CU(cudaEventCreate(&cs->ev1));
  }
  const size_t n1 = (size_t)n + 1;
  CU(cs->d_hits.reserve(n1));
  CU(cs->d_cap.reserve(n1));
  CU(cs->d_count.reserve(n1));
  CU(cs->d_max.reserve(n1));
  CU(cs->d_a.reserve(n1)); CU(cs->d_b.reserve(n1)); CU(cs->d_c.reserve(n1));
  CU(cs->d_sa.reserve(n1)); CU(cs->d_sb.reserve(n1)); CU(cs->d_sc.reserve(n1));
  CU(cs->d_cnt64.reserve(n1)); CU(cs->d_cstart.reserve(n1));
  size_t tmp_bytes = 0;
  CU(cs_exclusive_scan(nullptr, tmp_bytes, cs->d_a.p, cs->d_sa.p, (int)n1, st));
  CU(cs->d_scan_tmp.reserve(tmp_bytes + 256));
  CsParams p;
  memset(&p, 0, sizeof(p));
  p.tab = cs->d_tab.p;
  p.used_bits = cs->d_used.p;
  p.pos = cs->d_pos.p;
  p.unit_offset = cs->unit_offset;
  p.k = cs->k;
  p.bin_shift = cs->bin_shift;
  p.sensitivity = sensitivity;
  p.min_kmer_hits = min_kmer_hits;
  p.seq = cs->d_seq.p;
  p.seq_off = cs->d_off.p;
  p.seq_len = cs->d_len.p;
  p.n = n;
  p.hits = cs->d_hits.p;
  CU(cudaEventRecord(cs->ev0, st));
  CU(launch_cs_search(p, true, st));
  CU(launch_cs_sizes(cs->d_hits.p, n, cs->d_cap.p, cs->d_a.p, cs->d_b.p, cs->d_c.p, st));
  size_t tb = cs->d_scan_tmp.cap;
  CU(cs_exclusive_scan(cs->d_scan_tmp.p, tb, cs->d_a.p, cs->d_sa.p, (int)n1, st));
  tb = cs->d_scan_tmp.cap;
  CU(cs_exclusive_scan(cs->d_scan_tmp.p, tb, cs->d_b.p, cs->d_sb.p, (int)n1, st));
  tb = cs->d_scan_tmp.cap;
  CU(cs_exclusive_scan(cs->d_scan_tmp.p, tb, cs->d_c.p, cs->d_sc.p, (int)n1, st));
  unsigned long long totals[3] = {0, 0, 0};
  CU(cudaMemcpyAsync(&totals[0], cs->d_sa.p + n, 8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(&totals[1], cs->d_sb.p + n, 8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(&totals[2], cs->d_sc.p + n, 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  const size_t need = (size_t)totals[0] * 16 + (size_t)totals[1] * 4 + (size_t)totals[2] * 16;
  if (need > ((size_t)48 << 30))
    return ctx->fail("cs_run: %zu bytes of vote tables needed; use cs_score_batch (chunked) for this batch", need);
  CU(cs->d_tables.reserve((size_t)totals[0] * 16 + 16));
  CU(cs->d_order.reserve((size_t)totals[1] + 4));
  CU(cs->d_out.reserve((size_t)totals[2] + 4));
  CU(cudaMemsetAsync(cs->d_tables.p, 0, (size_t)totals[0] * 16, st));
  p.tables = cs->d_tables.p;
  p.table_off = reinterpret_cast<const uint64_t*>(cs->d_sa.p);
  p.table_cap = cs->d_cap.p;
  p.order = cs->d_order.p;
  p.order_off = reinterpret_cast<const uint64_t*>(cs->d_sb.p);
  p.out = cs->d_out.p;
  p.out_off = reinterpret_cast<const uint64_t*>(cs->d_sc.p);
  p.out_count = cs->d_count.p;
  p.max_hits = cs->d_max.p;
  CU(launch_cs_search(p, false, st));
  CU(launch_cs_count_to_u64(cs->d_count.p, n, cs->d_cnt64.p, st));
  tb = cs->d_scan_tmp.cap;
  CU(cs_exclusive_scan(cs->d_scan_tmp.p, tb, cs->d_cnt64.p, cs->d_cstart.p, (int)n1, st));
  unsigned long long m64 = 0;
  CU(cudaMemcpyAsync(&m64, cs->d_cstart.p + n, 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  const size_t m = (size_t)m64;
  cs->n_cand = (long long)m;
  CU(cs->d_cloc.reserve(m + 1));
  CU(cs->d_cscore.reserve(m + 1));
  CU(cs->d_rev.reserve(m + 1));
  CU(cs->d_winpos.reserve(m + 1));
  CU(cs->d_qoff.reserve(m + 1));
  CU(cs->d_qlen.reserve(m + 1));
  CU(cs->d_sw.reserve(m + 1));
  CU(launch_cs_compact(cs->d_out.p, reinterpret_cast<const uint64_t*>(cs->d_sc.p), cs->d_count.p, cs->d_cstart.p,
                       cs->d_off.p, cs->d_len.p, n, corridor >> 1, cs->d_cloc.p, cs->d_cscore.p, cs->d_rev.p,
                       cs->d_winpos.p, cs->d_qoff.p, cs->d_qlen.p, st));
  if (m) {
    const int win_len = ((read_part_length + 10 + corridor) | 1) + 1;  // refMaxLen, src/ScoreBuffer.h:71-72
    const int warps_per_cta = 4;
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((size_t)ctx->num_sms * 8, (m + warps_per_cta - 1) / warps_per_cta));
    const size_t stride = align_up((size_t)win_len + 4, 4);
    CU(cs->d_swscratch.reserve((size_t)grid * warps_per_cta * stride * 2));
    SwParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.seq = cs->d_seq.p;
    sp.ref_off = cs->d_qoff.p;
    sp.qry_off = cs->d_qoff.p;
    sp.ref_len = cs->d_qlen.p;
    sp.qry_len = cs->d_qlen.p;
    sp.out = cs->d_sw.p;
    sp.n = (int)m;
    sp.scratch = cs->d_swscratch.p;
    sp.scratch_stride = stride;
    sp.enc = cs->d_enc.p;
    sp.concat_len = cs->concat_len;
    sp.win_pos = cs->d_winpos.p;
    sp.rev = cs->d_rev.p;
    sp.win_len = win_len;
    CU(launch_sw_score_gather(sp, grid, st));
  }
  CU(cudaEventRecord(cs->ev1, st));
  CU(cudaStreamSynchronize(st));
  if (n_candidates) *n_candidates = (int64_t)m;
  if (kernel_ms) cudaEventElapsedTime(kernel_ms, cs->ev0, cs->ev1);
  return 0;
}

int ngmlr_b200_cs_fetch(ngmlr_b200_ctx* ctx, int64_t* cand_start, const float** cs_scores,
                        const uint64_t** locs, const uint8_t** reverse, const float** sw_scores,
                        float* max_hits) {
  if (!ctx) return -1;
  CsState* cs = cs_state(ctx, false);
  if (!cs) return ctx->fail("cs_fetch: nothing to fetch");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int n = cs->rn;
  const size_t m = (size_t)cs->n_cand;
  CU(cs->p_score.reserve(m + 1));
  CU(cs->p_loc.reserve(m + 1));
  CU(cs->p_rev.reserve(m + 1));
  CU(cs->p_sw.reserve(m + 1));
  static_assert(sizeof(unsigned long long) == sizeof(int64_t), "");
  if (n) CU(cudaMemcpyAsync(cand_start, cs->d_cstart.p, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, st));
  else cand_start[0] = 0;
  if (n && max_hits) CU(cudaMemcpyAsync(max_hits, cs->d_max.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  if (m) {
    CU(cudaMemcpyAsync(cs->p_score.p, cs->d_cscore.p, m * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(cs->p_loc.p, cs->d_cloc.p, m * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(cs->p_rev.p, cs->d_rev.p, m, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(cs->p_sw.p, cs->d_sw.p, m * 4, cudaMemcpyDeviceToHost, st));
  }
  CU(cudaStreamSynchronize(st));
  *cs_scores = cs->p_score.p;
  *locs = cs->p_loc.p;
  *reverse = cs->p_rev.p;
  *sw_scores = cs->p_sw.p;
  return n;
}

// Host glue after scoring: ScoreBuffer::topNSE + computeMQ (src/ScoreBuffer.cpp:170-192, 33-45).
// Candidates of one (sub-)read are ordered with std::sort and the reference's comparator
// (a.Score.f > b.Score.f, :25-27) -- the same libstdc++ routine, because the order of tied scores is
// part of the behaviour --, then those scoring above 0.75 x best are kept.
int ngmlr_b200_select_candidates(int n, const int64_t* cand_start, const float* sw_scores, int32_t* order,
                                 int32_t* kept, int32_t* mq) {
  if (n < 0 || (n > 0 && (!cand_start || !order || !kept || !mq))) return -1;
  struct Item {
    float score;
    int32_t idx;
  };
  parallel_for(n, 512, [&](int i) {
    const int64_t b = cand_start[i];
    const int m = (int)(cand_start[i + 1] - b);
    Item small[32];
    std::vector<Item> big;
    Item* it = small;
    if (m > 32) {
      big.resize((size_t)m);
      it = big.data();
    }
    for (int j = 0; j < m; ++j) {
      it[j].score = sw_scores[b + j];
      it[j].idx = (int32_t)(b + j);
    }
    std::sort(it, it + m, [](Item x, Item y) { return x.score > y.score; });
    int keep = m;
    if (m > 1) {
      const float min_score = it[0].score * 0.75f;
      int j = 1;
      while (j < m && it[j].score > min_score) ++j;
      keep = j;
    }
    int q = 60;  // MAX_MQ (src/ScoreBuffer.cpp:16)
    if (m > 1) q = (int)ceil(60.0f * (it[0].score - it[1].score) / it[0].score);
    for (int j = 0; j < m; ++j) order[b + j] = it[j].idx;
    kept[i] = keep;
    mq[i] = q;
  });
  return n;
}

}  // extern "C"
