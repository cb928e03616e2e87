// ngmlr_b200/csrc/convex_fill.cu -- convex-gap banded Smith-Waterman forward fill for sm_100a.
//
// Replaces Convex::ConvexAlignFast::fwdFillMatrixSSESimple (src/ConvexAlignFast.cpp:914-1287)
// over Convex::AlignmentMatrixFast (src/AlignmentMatrixFast.{h,cpp}).
//
// Mapping. One warp owns one alignment problem (a persistent grid pulls problems, largest
// first, from an atomic counter). The warp walks the DP matrix in blocks of 32 rows: lane t owns
// row y = 32*b + t and sweeps it left to right, staggered one step behind lane t-1, so that at
// step s lane t evaluates column x = base + s - t. That is an anti-diagonal wavefront:
//   up   (x,   y-1)  = what lane t-1 produced one step ago   -> __shfl_up
//   diag (x-1, y-1)  = what lane t-1 produced two steps ago  -> last step's "up", kept in a register
//   left (x-1, y)    = this lane's previous cell             -> registers
// The rolling rows of the reference (2 x W x 8 B, AlignmentMatrixFast.h:34-54) therefore never
// leave the register file. Row 32*b+31 is handed to lane 0 of the next block through a per-warp
// strip in global memory indexed by absolute column (it stays in L2; ~16 B/column/block), staged
// through shared memory in 64-step chunks: all 32 lanes copy a chunk of the strip (+ the reference
// bytes for those columns, merged into the same 16-byte records) into shared memory one chunk
// ahead, lane 0 then needs a single LDS.128 per step and lane 31 a single STS.128; finished
// chunks are flushed back coalesced. Because lane 31 always trails lane 0 by 31 columns the strip
// is updated in place for any corridor shape.
//
// HBM traffic. The only per-cell output is the traceback direction: 2 bits per cell (EQ/X are
// re-derived by the traceback), 16 steps per 32-bit word, written as one fully coalesced 128-byte
// warp store every 16 steps (word index = group*32 + lane). The reference spends 1 byte per cell
// (AlignmentMatrixFast.h:261).
//
// Arithmetic is float32 exactly as the reference: every add/multiply is a separately rounded
// __fadd_rn/__fmul_rn (the reference binary has no FMA), comparisons are exact, and the
// direction priority is the reference's. `RAW` selects the as-coded SSE semantics in which the
// run tests use the neighbours' raw indelRun for all but the last <=12 cells of a row (executable
// spec: oracle/convex_oracle.c chain_cell, rule 2); RAW=false is the scalar rule, which is
// identical for every scoring in the "default class" (see capi.cu scoring_needs_raw()).
#include <cuda_runtime.h>
#include <limits.h>

#include "device_types.h"
#include "kernels.h"

namespace nb {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int CHUNK = 64;        // steps staged through shared memory at a time
constexpr int STRIP_PAD = 32;    // strip index = column + STRIP_PAD (lane 31 trails lane 0 by 31)

__device__ __forceinline__ uint4 ld_strip(const uint4* p) { return __ldcg(p); }
__device__ __forceinline__ void st_strip(uint4* p, uint4 v) { __stcg(p, v); }

template <bool RAW>
__global__ void __launch_bounds__(FILL_WARPS_PER_CTA * 32, FILL_CTAS_PER_SM)
convex_fill_kernel(const FillParams p) {
  __shared__ uint4 s_in[FILL_WARPS_PER_CTA][CHUNK + 1];  // +1: lane 31 reads one record ahead
  __shared__ uint4 s_out[FILL_WARPS_PER_CTA][CHUNK];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int warp_global = blockIdx.x * FILL_WARPS_PER_CTA + wib;
  uint4* const in_s = s_in[wib];
  uint4* const out_s = s_out[wib];
  // strip[x + STRIP_PAD] = {S, U, pack, -} of column x of the last finished block's bottom row
  uint4* const strip = reinterpret_cast<uint4*>(p.bnd) + (size_t)warp_global * p.bnd_stride + STRIP_PAD;
  const Scoring sc = p.sc;
  const uint32_t empty_pack = RAW ? (DIR_STOP << 16) : 0u;  // scalar kernel: run 0.0f
  const uint4 EMPTY = make_uint4(0u, __float_as_uint(sc.open_read), empty_pack, 0u);
  const bool is0 = lane == 0, is31 = lane == 31;
  const int src_lane = (lane + 31) & 31;  // rotate: lane 0 receives what lane 31 staged for it

  for (;;) {
    int w = 0;
    if (is0) w = atomicAdd(p.work_counter, 1);
    w = __shfl_sync(FULL, w, 0);
    if (w >= p.n) break;
    const int ai = p.order[w];
    const AlnDesc d = p.desc[ai];
    const uint8_t* __restrict__ ref = p.seq + d.ref_off;
    const uint8_t* __restrict__ qry = p.seq + d.qry_off;
    const int32_t* __restrict__ coff = p.c_off + d.row_off;
    const int32_t* __restrict__ clen = p.c_len + d.row_off;
    const int H = d.height, ref_len = d.ref_len;
    const int nblk = (H + 31) >> 5;

    // curr_max starts at -1 (:921), so the first visited cell is always recorded and only strictly
    // larger scores replace it: track improvements over 0 here and remember the first visited cell.
    float bestS = 0.0f;
    int bestX = 0, bestY = 0;
    int firstX = 0, firstY = -1;
    unsigned long long cells = 0;
    int status = ST_OK;

    // columns of the strip that hold valid records of the row above the current block
    int wlo = 0, whi = 0;  // nothing written yet: the row above block 0 does not exist

    // rows of the next block are fetched one block ahead
    int n_off = 0, n_len = 0;
    uint32_t n_q = 0x100u;  // never equals a byte
    if (lane < H) {
      n_off = coff[lane];
      n_len = clen[lane];
      n_q = qry[lane];
    }

    for (int b = 0; b < nblk; ++b) {
      const int y = (b << 5) + lane;
      const int off = n_off, len = n_len;
      const uint32_t q = n_q;
      {
        const int yn = y + 32;
        n_off = 0; n_len = 0; n_q = 0x100u;
        if (yn < H) {
          n_off = coff[yn];
          n_len = clen[yn];
          n_q = qry[yn];
        }
      }
      // columns of this row: [max(0,off), min(off+len, refLen))   (:943-950)
      const int xlo = off > 0 ? off : 0;
      const long long hi64 = (long long)off + (long long)len;
      const int xhi = hi64 < (long long)ref_len ? (int)hi64 : ref_len;
      const unsigned rlen = xhi > xlo ? (unsigned)(xhi - xlo) : 0u;
      // base = leftmost column of the block: lane t first becomes active at step >= t, i.e. after the
      // reference byte for its column has travelled down the shuffle chain from lane 0
      const int lo_key = rlen ? xlo : INT_MAX;
      const int hi_key = rlen ? xhi + lane : INT_MIN;
      int base = __reduce_min_sync(FULL, lo_key);
      const int send = __reduce_max_sync(FULL, hi_key);
      int nsteps = 0;
      if (base != INT_MAX) nsteps = send - base; else base = 0;
      const int ngroups = (nsteps + 15) >> 4;
      const int nchunks = (ngroups + 3) >> 2;
      cells += rlen;
      if (firstY < 0) {
        const unsigned any = __ballot_sync(FULL, rlen != 0);
        if (any) {
          const int l0 = __ffs(any) - 1;
          firstY = (b << 5) + l0;
          firstX = __shfl_sync(FULL, xlo, l0);
        }
      }

      unsigned long long word_off = 0;
      if (is0) {
        word_off = atomicAdd(p.dir_alloc, (unsigned long long)ngroups * 32ull);
        BlockRec br;
        br.word_off = word_off;
        br.base = base;
        br.nsteps = nsteps;
        p.blocks[d.blk_off + b] = br;
      }
      word_off = __shfl_sync(FULL, word_off, 0);
      if (word_off + (unsigned long long)ngroups * 32ull > p.dir_capacity) {
        status = ST_DIR_OVERFLOW;
        break;
      }
      uint32_t* __restrict__ dwp = p.dir + word_off + lane;

      // The strip must hold a record of the row above for every column lane 0 will visit,
      // [base-1, base + nchunks*64): what the previous block did not write is EMPTY.
      {
        const int need_lo = base - 1, need_hi = base + nchunks * CHUNK;
        const int l_hi = wlo < need_hi ? wlo : need_hi;
        for (int x = need_lo + lane; x < l_hi; x += 32) st_strip(strip + x, EMPTY);
        const int r_lo = whi > need_lo ? whi : need_lo;
        for (int x = r_lo + lane; x < need_hi; x += 32) st_strip(strip + x, EMPTY);
      }
      __syncwarp();

      int rel = base - lane - xlo;  // x - xlo at step 0; inside the corridor iff (unsigned)rel < rlen
      const int t0rel = (int)rlen > 12 ? (int)rlen - 12 : 0;  // tail start max(x0, xMax-12) - xlo (:1179)

      // What this lane hands down / keeps for its right neighbour. EMPTY = {0, 0, STOP}.
      // oP: scalar kernel = the I-run length as a float (0 unless the cell is an insertion);
      //     RAW kernel    = indelRun (low 16 bits) | direction << 16.
      float oS = 0.0f, oU = sc.open_read;
      uint32_t oP = empty_pack, oC = 0u;
      float dS = 0.0f;         // score of (x-1, y-1)
      float lL = sc.open_ref;  // left_cell contribution of (x-1, y)
      float lRunF = 0.0f;      // scalar kernel: D-run length of (x-1, y), 0 unless it is a deletion
      int lRun = 0;            // RAW kernel: raw indelRun / direction of (x-1, y)
      uint32_t lDir = DIR_STOP;
      if (is0) dS = __uint_as_float(ld_strip(strip + base - 1).x);
      float kS = bestS;
      int kStep = -1;

      // stage chunk 0: strip records + reference bytes for columns [base, base+64)
      uint4 pa, pb;
      uint32_t ra, rb;
      {
        const int x0 = base + lane;
        pa = ld_strip(strip + x0);
        pb = ld_strip(strip + x0 + 32);
        ra = __ldg(ref + x0);
        rb = __ldg(ref + x0 + 32);
      }

      for (int c = 0; c < nchunks; ++c) {
        pa.w = ra;  // the reference byte of the column rides in the record
        pb.w = rb;
        in_s[lane] = pa;
        in_s[lane + 32] = pb;
        __syncwarp();
        if (is31) {  // lane 31's shuffle sources carry the strip record lane 0 needs next
          const uint4 t = in_s[0];
          oS = __uint_as_float(t.x);
          oU = __uint_as_float(t.y);
          oP = t.z;
          oC = t.w;
        }
        if (c + 1 < nchunks) {  // fetch the next chunk while this one is computed
          const int x0 = base + (c + 1) * CHUNK + lane;
          pa = ld_strip(strip + x0);
          pb = ld_strip(strip + x0 + 32);
          ra = __ldg(ref + x0);
          rb = __ldg(ref + x0 + 32);
        }
        const int g_end = min(ngroups, (c + 1) << 2);
        for (int g = c << 2; g < g_end; ++g) {
          const uint4* in_g = in_s + ((g & 3) << 4);
          uint4* out_g = out_s + ((g & 3) << 4);
          uint32_t dw = 0;
#pragma unroll 1
          for (int k4 = 0; k4 < 16; k4 += 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int s = (g << 4) + k4 + k;
              uint4 v;
              v.x = __float_as_uint(__shfl_sync(FULL, oS, src_lane));
              v.y = __float_as_uint(__shfl_sync(FULL, oU, src_lane));
              v.z = __shfl_sync(FULL, oP, src_lane);
              v.w = __shfl_sync(FULL, oC, src_lane);
              const float nS = __uint_as_float(v.x), nU = __uint_as_float(v.y);
              const uint32_t r = v.w;
              const bool act = (unsigned)rel < rlen;
              const float sub = (r == q) ? sc.mat : sc.mis;
              const float dg = __fadd_rn(dS, sub);
              dS = nS;
              // Outside the corridor the cell must degenerate to {0, 0, STOP}: a NaN maximum makes every
              // equality below false, and fmaxf(NaN, 0) = 0 gives the score.
              const float m = act ? fmaxf(fmaxf(fmaxf(lL, 0.0f), dg), nU) : __int_as_float(0x7fffffff);
              const bool eL = (m == lL), eU = (m == nU), eG = (m == dg);
              const float S = fmaxf(m, 0.0f);  // STOP implies m == 0
              uint32_t code;
              float U, L;
              if (RAW) {
                const uint32_t nP = v.z;
                const int upRaw = (int)(short)(nP & 0xffffu);
                const uint32_t upDir = (nP >> 16) & 3u;
                const bool rawHere = rel < t0rel;
                const int upRun = (rawHere || upDir == DIR_I) ? upRaw : 0;
                const int leftRun = (rawHere || lDir == DIR_D) ? lRun : 0;
                // priority (:1232-1267): continue D, continue I, diagonal, open D, open I, STOP
                const bool dc = eL && (leftRun > 0);
                const bool ic = !dc && eU && (upRun > 0);
                const bool gg = !dc && !ic && eG;
                const bool resolved = dc || ic || gg;
                const bool isD = dc || (!resolved && eL);
                const bool isI = ic || (!resolved && !eL && eU);
                int run = dc ? leftRun : (ic ? upRun : 0);
                run = (isD || isI) ? run + 1 : 0;
                run = (int)(short)run;  // MatrixElement::indelRun is a short
                code = gg ? DIR_DIAG : (isI ? DIR_I : (isD ? DIR_D : DIR_STOP));
                // what the neighbours will see: S + min(ext_min, gap_ext + run*decay), 0 if S == 0 (:666-676)
                const float pen = fminf(sc.ext_min, __fadd_rn(sc.gap_ext, __fmul_rn((float)run, sc.decay)));
                float e = __fadd_rn(S, pen);
                if (S == 0.0f) e = 0.0f;
                U = isI ? e : __fadd_rn(S, sc.open_read);
                L = isD ? e : __fadd_rn(S, sc.open_ref);
                oP = (code << 16) | ((uint32_t)run & 0xffffu);
                lRun = run;
                lDir = code;
              } else {
                // Same priority chain as a 3-input predicate network (verified exhaustively):
                //   D  <=>  eL && (lr || !((eU && ur) || eG))
                //   I  <=>  !D && eU && (ur || !eG)
                // with run lengths kept as floats (exact below 2^24; rows are < 32768 wide here).
                const float upRunF = __uint_as_float(v.z);
                const bool lr = lRunF > 0.0f, ur = upRunF > 0.0f;
                const bool X = (eU && ur) || eG;
                const bool pD = eL && (lr || !X);
                const bool pI = !pD && eU && (ur || !eG);
                code = pD ? DIR_D : (pI ? DIR_I : (eG ? DIR_DIAG : DIR_STOP));
                const float runF = pD ? __fadd_rn(lRunF, 1.0f) : (pI ? __fadd_rn(upRunF, 1.0f) : 0.0f);
                const float pen = fminf(sc.ext_min, __fadd_rn(sc.gap_ext, __fmul_rn(runF, sc.decay)));
                float e = __fadd_rn(S, pen);
                if (S == 0.0f) e = 0.0f;
                U = pI ? e : __fadd_rn(S, sc.open_read);
                L = pD ? e : __fadd_rn(S, sc.open_ref);
                oP = __float_as_uint(pI ? runF : 0.0f);
                lRunF = pD ? runF : 0.0f;
              }
              oS = S;
              oU = U;
              oC = r;
              lL = L;
              if (S > kS) {  // strict: first maximum in row-major order (:1165-1170)
                kS = S;
                kStep = s;
              }
              dw = __funnelshift_r(dw, code, 2);
              if (is31) {
                out_g[k4 + k] = make_uint4(__float_as_uint(S), __float_as_uint(U), oP, 0u);
                const uint4 t = in_g[k4 + k + 1];
                oS = __uint_as_float(t.x);
                oU = __uint_as_float(t.y);
                oP = t.z;
                oC = t.w;
              }
              ++rel;
            }
          }
          dwp[(size_t)g * 32] = dw;
        }
        __syncwarp();
        // flush lane 31's records of this chunk: columns [base - 31 + 64c, ...)
        {
          const int done = (g_end - (c << 2)) << 4;  // steps executed in this chunk
          const int xo = base - 31 + c * CHUNK;
          if (lane < done) st_strip(strip + xo + lane, out_s[lane]);
          if (lane + 32 < done) st_strip(strip + xo + lane + 32, out_s[lane + 32]);
        }
        __syncwarp();
      }

      if (kStep >= 0) {
        bestS = kS;
        bestY = y;
        bestX = base + kStep - lane;
      }
      wlo = base - 31;
      whi = base - 31 + (ngroups << 4);
      __syncwarp();
    }

    // first maximum in row-major order across lanes: larger score, then smaller y, then smaller x
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float s2 = __shfl_xor_sync(FULL, bestS, o);
      const int y2 = __shfl_xor_sync(FULL, bestY, o);
      const int x2 = __shfl_xor_sync(FULL, bestX, o);
      const unsigned long long c2 = __shfl_xor_sync(FULL, cells, o);
      const bool take = (s2 > bestS) || (s2 == bestS && (y2 < bestY || (y2 == bestY && x2 < bestX)));
      if (take) {
        bestS = s2;
        bestY = y2;
        bestX = x2;
      }
      cells += c2;
    }
    if (bestS == 0.0f) {  // no positive score anywhere: the first visited cell stands (or nothing was visited)
      bestS = firstY >= 0 ? 0.0f : -1.0f;
      bestX = firstY >= 0 ? firstX : 0;
      bestY = firstY >= 0 ? firstY : 0;
    }
    if (is0) {
      FillOut o;
      o.best_score = bestS;
      o.best_x = bestX;
      o.best_y = bestY;
      o.status = status;
      o.cells = cells;
      p.out[ai] = o;
    }
  }
}

}  // namespace

cudaError_t launch_convex_fill(const FillParams& p, bool raw, int grid, cudaStream_t stream) {
  if (raw)
    convex_fill_kernel<true><<<grid, FILL_WARPS_PER_CTA * 32, 0, stream>>>(p);
  else
    convex_fill_kernel<false><<<grid, FILL_WARPS_PER_CTA * 32, 0, stream>>>(p);
  return cudaGetLastError();
}

int fill_max_ctas_per_sm(bool raw) {
  int n = 0;
  if (raw)
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, convex_fill_kernel<true>, FILL_WARPS_PER_CTA * 32, 0);
  else
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, convex_fill_kernel<false>, FILL_WARPS_PER_CTA * 32, 0);
  return n;
}

}  // namespace nb
