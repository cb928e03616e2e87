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
// double-buffered strip in global memory that stays in L2 (one 16-byte store by lane 31 and one
// 16-byte load by lane 0 per step, the load prefetched one step ahead).
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

__device__ __forceinline__ BndEntry ld_bnd(const BndEntry* p) {
  uint4 v = __ldcg(reinterpret_cast<const uint4*>(p));
  BndEntry e;
  e.S = __uint_as_float(v.x);
  e.U = __uint_as_float(v.y);
  e.pack = v.z;
  e.pad = 0;
  return e;
}

__device__ __forceinline__ void st_bnd(BndEntry* p, float S, float U, uint32_t pack) {
  __stcg(reinterpret_cast<uint4*>(p), make_uint4(__float_as_uint(S), __float_as_uint(U), pack, 0u));
}

template <bool RAW>
__global__ void __launch_bounds__(FILL_WARPS_PER_CTA * 32, FILL_CTAS_PER_SM)
convex_fill_kernel(const FillParams p) {
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * FILL_WARPS_PER_CTA + (threadIdx.x >> 5);
  BndEntry* const bnd0 = p.bnd + (size_t)warp_global * 2 * p.bnd_stride;
  const Scoring sc = p.sc;
  const uint32_t empty_pack = RAW ? (DIR_STOP << 16) : 0u;

  for (;;) {
    int w = 0;
    if (lane == 0) w = atomicAdd(p.work_counter, 1);
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

    float bestS = -1.0f;  // curr_max starts at -1 (:921)
    int bestX = 0, bestY = 0;
    unsigned long long cells = 0;
    int status = ST_OK;

    // Row above the first block does not exist: an empty boundary (sentinel at index 0).
    int pxlo = 0;
    unsigned plen = 0;
    int cur = 0;
    if (lane == 0) st_bnd(bnd0, 0.0f, sc.open_read, empty_pack);
    __syncwarp();

    for (int b = 0; b < nblk; ++b) {
      const int y = (b << 5) + lane;
      int off = 0, len = 0;
      uint32_t q = 0x100u;  // never equals a byte
      if (y < H) {
        off = coff[y];
        len = clen[y];
        q = qry[y];
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
      cells += rlen;

      unsigned long long word_off = 0;
      if (lane == 0) {
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
      const BndEntry* bin = bnd0 + (size_t)cur * p.bnd_stride;
      BndEntry* bout = bnd0 + (size_t)(cur ^ 1) * p.bnd_stride;

      int rel = base - lane - xlo;  // x - xlo at step 0; the cell is inside the corridor iff (unsigned)rel < rlen
      const int t0rel = (int)rlen > 12 ? (int)rlen - 12 : 0;  // tail start max(x0, xMax-12) relative to xlo (:1179)
      int brel = base - pxlo;       // lane 0: index of column `base` in the boundary strip

      float oS = 0.0f, oU = sc.open_read;  // what this lane hands down: EMPTY = {0,0,STOP}
      uint32_t oPack = empty_pack;
      float dS = 0.0f;                     // score of (x-1, y-1)
      float lL = sc.open_ref;              // left_cell contribution of (x-1, y)
      int lRun = 0;
      uint32_t lDir = DIR_STOP;
      BndEntry pf;
      pf.S = 0.0f; pf.U = 0.0f; pf.pack = 0u; pf.pad = 0u;
      uint32_t pfr = 0;
      if (lane == 0) {
        const unsigned im1 = min((unsigned)(brel - 1), plen);
        dS = ld_bnd(bin + im1).S;
        pf = ld_bnd(bin + min((unsigned)brel, plen));
        pfr = __ldg(ref + base);
      }
      float kS = bestS;
      int kStep = -1;

      for (int g = 0; g < ngroups; ++g) {
        uint32_t dw = 0;
#pragma unroll 1
        for (int k4 = 0; k4 < 16; k4 += 4) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int s = (g << 4) + k4 + k;
            float nS = __shfl_up_sync(FULL, oS, 1);
            float nU = __shfl_up_sync(FULL, oU, 1);
            uint32_t nP = __shfl_up_sync(FULL, oPack, 1);
            if (lane == 0) {
              nS = pf.S;
              nU = pf.U;
              nP = (pf.pack & 0x00ffffffu) | (pfr << 24);
              ++brel;
              pf = ld_bnd(bin + min((unsigned)brel, plen));
              pfr = __ldg(ref + base + s + 1);
            }
            const uint32_t r = nP >> 24;
            const bool act = (unsigned)rel < rlen;
            int upRun, leftRun;
            if (RAW) {
              const int upRaw = (int)(short)(nP & 0xffffu);
              const uint32_t upDir = (nP >> 16) & 3u;
              const bool rawHere = rel < t0rel;
              upRun = (rawHere || upDir == DIR_I) ? upRaw : 0;
              leftRun = (rawHere || lDir == DIR_D) ? lRun : 0;
            } else {
              upRun = (int)(nP & 0xffffu);
              leftRun = lRun;
            }
            const float sub = (r == q) ? sc.mat : sc.mis;
            const float dg = __fadd_rn(dS, sub);
            dS = nS;
            const float m = fmaxf(fmaxf(fmaxf(lL, 0.0f), dg), nU);
            const bool eL = (m == lL), eU = (m == nU), eG = (m == dg);
            // priority (:1232-1267): continue D, continue I, diagonal, open D, open I, STOP
            const bool dc = eL && (leftRun > 0);
            const bool ic = !dc && eU && (upRun > 0);
            const bool gg = !dc && !ic && eG;
            const bool resolved = dc || ic || gg;
            const bool dn = !resolved && eL;
            const bool in = !resolved && !eL && eU;
            const bool isD = dc || dn, isI = ic || in;
            int run = dc ? leftRun + 1 : (ic ? upRun + 1 : ((dn || in) ? 1 : 0));
            if (RAW) run = (int)(short)run;  // MatrixElement::indelRun is a short
            uint32_t code = gg ? DIR_DIAG : (isI ? DIR_I : (isD ? DIR_D : DIR_STOP));
            float S = m;  // STOP implies m == 0
            // what the neighbours will see: S + min(ext_min, gap_ext + run*decay), 0 if S == 0 (:666-676)
            const float pen = fminf(sc.ext_min, __fadd_rn(sc.gap_ext, __fmul_rn((float)run, sc.decay)));
            float e = __fadd_rn(S, pen);
            if (S == 0.0f) e = 0.0f;
            float U = isI ? e : __fadd_rn(S, sc.open_read);
            float L = isD ? e : __fadd_rn(S, sc.open_ref);
            if (!act) {  // outside the corridor / reference: reads as {0, 0, STOP}
              S = 0.0f;
              U = sc.open_read;
              L = sc.open_ref;
              run = 0;
              code = DIR_STOP;
            }
            oS = S;
            oU = U;
            lL = L;
            if (RAW) {
              oPack = (r << 24) | (code << 16) | ((uint32_t)run & 0xffffu);
              lRun = run;
              lDir = code;
            } else {
              oPack = (r << 24) | (uint32_t)((act && isI) ? run : 0);
              lRun = (act && isD) ? run : 0;
            }
            if (act && S > kS) {  // strict: first maximum in row-major order (:1165-1170)
              kS = S;
              kStep = s;
            }
            dw = __funnelshift_r(dw, code, 2);
            if (lane == 31 && act) st_bnd(bout + rel, S, U, oPack & 0x00ffffffu);
            ++rel;
          }
        }
        dwp[(size_t)g * 32] = dw;
      }

      if (kStep >= 0) {
        bestS = kS;
        bestY = y;
        bestX = base + kStep - lane;
      }
      if (lane == 31) st_bnd(bout + rlen, 0.0f, sc.open_read, empty_pack);  // sentinel: EMPTY
      pxlo = __shfl_sync(FULL, xlo, 31);
      plen = __shfl_sync(FULL, rlen, 31);
      cur ^= 1;
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
    if (lane == 0) {
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
