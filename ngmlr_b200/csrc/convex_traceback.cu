// ngmlr_b200/csrc/convex_traceback.cu -- traceback + binary-CIGAR compaction for sm_100a.
//
// Replaces Convex::ConvexAlignFast::revBacktrack (src/ConvexAlignFast.cpp:335-432) with
// AlignmentMatrixFast::getDirection / validPath (src/AlignmentMatrixFast.cpp:185-195, 213-220).
//
// The walk is a pointer chase (each step depends on the previous direction), at most H + W steps
// against H x W cells of fill, so it is latency- not bandwidth-bound: one thread per problem, many
// problems in flight. Runs are emitted back-to-front into a per-problem scratch strip exactly like
// the reference's binaryCigar (element = len << 4 | op, EQ and X separate ops), then a second
// kernel copies each strip, warp-coalesced, into a compact arena so the host needs one D2H copy.
#include <cuda_runtime.h>

#include "device_types.h"
#include "kernels.h"

namespace nb {

namespace {

__global__ void __launch_bounds__(128) convex_traceback_kernel(const TraceParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const AlnDesc d = p.desc[i];
  const FillOut f = p.fill[i];
  TraceOut o;
  o.status = ST_INVALID;
  o.n_runs = 0;
  o.ref_position = 0;
  o.qstart = 0;
  o.qend = 0;
  o.steps = 0;
  o.run_off = 0;
  if (f.status != ST_OK) {
    o.status = f.status;
    p.out[i] = o;
    return;
  }
  const int H = d.height;
  const uint8_t* __restrict__ ref = p.seq + d.ref_off;
  const uint8_t* __restrict__ qry = p.seq + d.qry_off;
  const int32_t* __restrict__ coff = p.c_off + d.row_off;
  const int32_t* __restrict__ clen = p.c_len + d.row_off;
  const BlockRec* __restrict__ blocks = p.blocks + d.blk_off;
  int32_t* __restrict__ bc = p.scratch + d.tb_off;
  const int cap = d.tb_cap;        // our strip
  const int ref_cap = d.ref_cap;   // the reference's binaryCigar length (for its overflow throw)

  int x = f.best_x, y = f.best_y;
  const int qend = H - y - 1;  // (:1281)
  o.qend = qend;
  if (y <= 0) {  // (:338)
    p.out[i] = o;
    return;
  }
  int idx = cap - 1;       // next free slot, filled downwards
  int used = 0;            // slots the reference would have consumed
  int op = OP_S;
  int op_len = qend;
  int read_len = qend;
  int steps = 0;
  int cached_blk = -1;
  BlockRec br;
  br.word_off = 0; br.base = 0; br.nsteps = 0;
  bool ok = true, threw = false;

  for (;;) {
    // getDirection(x, y)
    int dir = OP_STOP;
    int off = 0, len = 0;
    if (y >= 0 && y <= H - 1 && x >= 0) {
      off = coff[y];
      len = clen[y];
      if (x >= off && (long long)x < (long long)off + (long long)len) {
        const int blk = y >> 5, t = y & 31;
        if (blk != cached_blk) {
          br = blocks[blk];
          cached_blk = blk;
        }
        const int s = x - br.base + t;
        const uint32_t wd = p.dir[br.word_off + (unsigned long long)(s >> 4) * 32ull + (unsigned)t];
        const uint32_t code = (wd >> ((s & 15) * 2)) & 3u;
        if (code == DIR_DIAG) dir = (qry[y] == ref[x]) ? OP_EQ : OP_X;
        else if (code == DIR_I) dir = OP_I;
        else if (code == DIR_D) dir = OP_D;
      }
    }
    if (dir == OP_STOP) break;
    // validPath(x, y): float math then truncation, as in the reference
    {
      const float w = (float)len;
      const int min_c = (int)__fadd_rn((float)off, __fmul_rn(0.1f, w));
      const int max_c = (int)__fsub_rn((float)(min_c + len), __fmul_rn(0.1f, w));
      if (!(x > min_c && x < max_c)) {
        ok = false;
        break;
      }
    }
    ++steps;
    if (dir == OP_EQ || dir == OP_X) {
      --y; --x; ++read_len;
    } else if (dir == OP_I) {
      --y; ++read_len;
    } else {
      --x;
    }
    if (dir == op) {
      ++op_len;
    } else {
      if (idx >= 0) bc[idx] = (op_len << 4) | op;
      --idx;
      ++used;
      op = dir;
      op_len = 1;
    }
    if (used >= ref_cap || idx < 0) {  // binaryCigarIndex < 0 -> throw 1 (:404-407)
      threw = true;
      break;
    }
  }
  o.steps = steps;
  if (threw) {
    o.status = ST_THROW;
    p.out[i] = o;
    return;
  }
  if (!ok) {
    p.out[i] = o;
    return;
  }
  // last run + leading clip (:411-416); the reference writes both slots unchecked
  if (idx < 1 || used + 2 > ref_cap) {
    o.status = ST_THROW;
    p.out[i] = o;
    return;
  }
  bc[idx--] = (op_len << 4) | op;
  bc[idx--] = ((y + 1) << 4) | OP_S;
  read_len += y + 1;
  o.ref_position = x + 1;
  o.qstart = y + 1;
  const int n = cap - 1 - idx;
  o.n_runs = n;
  if (H != read_len) {  // (:424-428)
    p.out[i] = o;
    return;
  }
  const unsigned long long at = atomicAdd(p.runs_alloc, (unsigned long long)n);
  o.run_off = at;
  o.status = (at + (unsigned long long)n <= p.runs_capacity) ? ST_OK : ST_DIR_OVERFLOW;
  p.out[i] = o;
}

// One warp per problem: strip [cap-n, cap) -> compact arena [run_off, run_off+n), same order
// (leading clip, runs..., trailing clip).
__global__ void __launch_bounds__(256) convex_compact_kernel(const TraceParams p) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= p.n) return;
  const TraceOut o = p.out[warp];
  if (o.status != ST_OK) return;
  const AlnDesc d = p.desc[warp];
  const int32_t* __restrict__ src = p.scratch + d.tb_off + (d.tb_cap - o.n_runs);
  int32_t* __restrict__ dst = p.runs + o.run_off;
  for (int k = lane; k < o.n_runs; k += 32) dst[k] = src[k];
}

}  // namespace

cudaError_t launch_convex_traceback(const TraceParams& p, cudaStream_t stream) {
  if (p.n <= 0) return cudaSuccess;
  convex_traceback_kernel<<<(p.n + 127) / 128, 128, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_convex_compact(const TraceParams& p, cudaStream_t stream) {
  if (p.n <= 0) return cudaSuccess;
  const int warps_per_cta = 8;
  convex_compact_kernel<<<(p.n + warps_per_cta - 1) / warps_per_cta, warps_per_cta * 32, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace nb
