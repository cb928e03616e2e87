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
// Team mode (NW = 4). One warp per problem leaves a long tail behind the largest matrices (the
// reference sees 93 M-cell matrices). In team mode the 4 warps of a CTA share one problem: warp w
// takes the 32-row blocks b = w, w+4, w+8, ... and block b+1 consumes the strip records of block b
// as they are produced (16-step chunks, progress published through shared memory), so the four
// warps run as a software pipeline about 100 columns apart and a problem finishes ~4x sooner at
// the same total work.
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

#include <type_traits>

#include "device_types.h"
#include "kernels.h"

namespace nb {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int STRIP_PAD = 32;    // strip index = column + STRIP_PAD (lane 31 trails lane 0 by 31)
constexpr unsigned X_BIAS = 1u << 20;
#ifndef FILL_TEAM_CHUNK
#define FILL_TEAM_CHUNK 16
#endif
#ifndef FILL_LANE0_LDS
#define FILL_LANE0_LDS 0  // 1: lane 0 loads its strip record itself instead of receiving it through lane 31's shuffle sources
#endif

__device__ __forceinline__ uint4 ld_strip(const uint4* p) { return __ldcg(p); }
__device__ __forceinline__ void st_strip(uint4* p, uint4 v) { __stcg(p, v); }

// CTA-scope fence ordering the strip records (st.cg / ld.cg) against the progress word in shared memory
// (fence.acq_rel.cta instead of the sequentially consistent fence measured no difference)
__device__ __forceinline__ void team_fence() { __threadfence_block(); }

// progress words live in shared memory; plain 32-bit shared addresses (computed once per kernel) instead of the
// generic-address sequence the compiler emits for a volatile shared array indexed at run time
__device__ __forceinline__ unsigned long long ld_prog(uint32_t saddr) {
  unsigned long long v;
  asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ void st_prog(uint32_t saddr, unsigned long long v) {
  asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(saddr), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long prog_key(int blk, int x) {
  return ((unsigned long long)(unsigned)(blk + 1) << 32) | (unsigned long long)((unsigned)x + X_BIAS);
}

// Geometry of one 32-row block, identical for the warp that fills it and the warp that consumes it.
struct BlockGeom {
  int base, ngroups;
};

__device__ __forceinline__ void row_span(int off, int len, int ref_len, int& xlo, int& xhi, unsigned& rlen) {
  // columns of a row: [max(0,off), min(off+len, refLen))   (:943-950)
  xlo = off > 0 ? off : 0;
  const long long hi64 = (long long)off + (long long)len;
  xhi = hi64 < (long long)ref_len ? (int)hi64 : ref_len;
  rlen = xhi > xlo ? (unsigned)(xhi - xlo) : 0u;
}

__device__ __forceinline__ BlockGeom block_geom(int xlo, int xhi, unsigned rlen, int lane, int& nsteps) {
  // base = leftmost column of the block: lane t first becomes active at step >= t, i.e. after the
  // reference byte for its column has travelled down the shuffle chain from lane 0
  const int lo_key = rlen ? xlo : INT_MAX;
  const int hi_key = rlen ? xhi + lane : INT_MIN;
  int base = __reduce_min_sync(FULL, lo_key);
  const int send = __reduce_max_sync(FULL, hi_key);
  nsteps = 0;
  if (base != INT_MAX) nsteps = send - base; else base = 0;
  BlockGeom g;
  g.base = base;
  g.ngroups = (nsteps + 15) >> 4;
  return g;
}

struct TeamBest {
  float S;
  int x, y, firstX, firstY, status;
  unsigned long long cells;
};

// NW = warps that pipeline ONE problem (1: every warp of the CTA has its own problem; otherwise the whole
// CTA is the team: 4 warps for ordinary corridors, FILL_BIG_TEAM warps for the few huge matrices of a batch
// -- a 10^8-cell realignment matrix would otherwise keep one 4-warp team busy long after the rest of the
// grid has drained).
template <bool RAW, int NW, bool PERSIST>
__global__ void __launch_bounds__((NW == 1 ? FILL_WARPS_PER_CTA : NW) * 32,
                                  NW > FILL_WARPS_PER_CTA ? 1 : (NW == 1 ? FILL_CTAS_PER_SM : FILL_TEAM_CTAS_PER_SM))
convex_fill_kernel(const FillParams p) {
  constexpr int WARPS = NW == 1 ? FILL_WARPS_PER_CTA : NW;  // warps per CTA
  constexpr int CHUNK = NW == 1 ? 64 : FILL_TEAM_CHUNK;  // steps staged through shared memory at a time
  constexpr int GPC = CHUNK / 16;           // 16-step groups per chunk
  // staging records of a chunk, interleaved: s_io[w][2 * j] = what lane 0 consumes at step j of the chunk,
  // s_io[w][2 * j + 1] = what lane 31 produced at step j (one base register + immediates serve both)
  __shared__ uint4 s_io[WARPS][2 * (CHUNK + 1)];  // +1: lane 31 reads one record ahead
  __shared__ volatile unsigned long long s_prog[WARPS];
  __shared__ int s_work, s_slot;
  __shared__ TeamBest s_best[WARPS];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int tw = NW == 1 ? 0 : wib;  // warp index within the team
  // Boundary strips: a persistent launch (grid capped, every CTA loops over problems) owns strip blockIdx.x. A
  // launch of short-lived CTAs (one problem each, so that SM slots keep coming free for the other kernels of the
  // step) takes one of p.sm_slot_count strips of the SM it happens to run on and gives it back when it exits; a
  // CTA that finds them all taken (registers would let one more team in than the kernel runs best with: more
  // strips competing for L2) sleeps until one is returned.
  unsigned smid = 0, slot_bit = 0;
  if (!PERSIST) {
    if (threadIdx.x == 0) {
      asm("mov.u32 %0, %%smid;" : "=r"(smid));
      int got = -1;
      for (;;) {
        for (int b = 0; b < p.sm_slot_count && got < 0; ++b)
          if (!(atomicOr(p.sm_slots + smid, 1u << b) & (1u << b))) got = b;
        if (got >= 0) break;
        __nanosleep(2000);
      }
      slot_bit = 1u << got;
      s_slot = (int)smid * FILL_SM_SLOTS + got;
    }
    __syncthreads();
  }
  const int cta_slot = PERSIST ? (int)blockIdx.x : s_slot;
  const int team_global = NW == 1 ? cta_slot * WARPS + wib : cta_slot;
  uint4* const io_s = s_io[wib];
  // strip[x + STRIP_PAD] = {S, U, run, ref byte} of column x of the most recently finished bottom row
  uint4* const strip = reinterpret_cast<uint4*>(p.bnd) + (size_t)team_global * p.bnd_stride + STRIP_PAD;
  const Scoring sc = p.sc;
  const uint32_t empty_pack = RAW ? (DIR_STOP << 16) : 0u;  // scalar kernel: run 0.0f
  const uint4 EMPTY = make_uint4(0u, __float_as_uint(sc.open_read), empty_pack, 0u);
  const bool is0 = lane == 0, is31 = lane == 31;
  const int src_lane = (lane + 31) & 31;  // rotate: lane 0 receives what lane 31 staged for it
  const uint32_t my_prog = (uint32_t)__cvta_generic_to_shared(const_cast<unsigned long long*>(&s_prog[tw]));
  const uint32_t prev_prog =
      (uint32_t)__cvta_generic_to_shared(const_cast<unsigned long long*>(&s_prog[(tw + NW - 1) % NW]));

  // A short-lived CTA takes p.problems_per_cta (= 1) problems per warp / team. The bound is a kernel parameter on
  // purpose: with a literal 1 the compiler peels the loop away and allocates registers for straight-line code,
  // which runs slower than the loop form the kernel was tuned in.
  for (int taken = 0; PERSIST || taken < p.problems_per_cta; ++taken) {
    int w = 0;
    if (NW == 1) {
      if (is0) w = atomicAdd(p.work_counter, 1);
      w = __shfl_sync(FULL, w, 0);
    } else {
      __syncthreads();  // previous problem fully retired (strip, progress words, s_work)
      if (threadIdx.x == 0) s_work = atomicAdd(p.work_counter, 1);
      if (threadIdx.x < NW) s_prog[threadIdx.x] = 0ull;
      __syncthreads();
      w = s_work;
    }
    w += p.first;  // this launch works on order[first, last)
    if (w >= p.last) break;
    const int ai = p.order[w];
    const AlnDesc d = p.desc[ai];
    const uint8_t* __restrict__ ref = p.seq + d.ref_off;
    const uint8_t* __restrict__ qry = p.seq + d.qry_off;
    CorridorView cv;

    cv.bind(p.c_off, p.c_len, p.c_blkbase, p.c_delta, d);
    const int H = d.height, ref_len = d.ref_len;
    const int nblk = (H + 31) >> 5;

    // curr_max starts at -1 (:921), so the first visited cell is always recorded and only strictly
    // larger scores replace it: track improvements over 0 here and remember the first visited cell.
    float bestS = 0.0f;
    int bestX = 0, bestY = 0;
    int firstX = 0, firstY = -1;
    unsigned long long cells = 0;
    int status = ST_OK;

    // strip columns [wlo, whi) hold records of the row above the current block; everything else
    // reads as EMPTY. Nothing is written before block 0.
    int wlo = 0, whi = 0;

    // rows of this warp's next block are fetched one block ahead
    int n_off = 0, n_len = 0;
    uint32_t n_q = 0x100u;  // never equals a byte
    load_corridor_rows(cv, tw, lane, H, n_off, n_len);
    if ((tw << 5) + lane < H) n_q = qry[(tw << 5) + lane];

    for (int b = tw; b < nblk; b += NW) {
      const int y = (b << 5) + lane;
      const int off = n_off, len = n_len;
      const uint32_t q = n_q;
      {
        const int yn = y + 32 * NW;
        n_q = 0x100u;
        load_corridor_rows(cv, b + NW, lane, H, n_off, n_len);
        if (yn < H) n_q = qry[yn];
      }
      int xlo, xhi, nsteps;
      unsigned rlen;
      row_span(off, len, ref_len, xlo, xhi, rlen);
      const BlockGeom geo = block_geom(xlo, xhi, rlen, lane, nsteps);
      const int base = geo.base, ngroups = geo.ngroups;
      const int nchunks = (ngroups + GPC - 1) / GPC;
      cells += rlen;
      if (firstY < 0) {
        const unsigned any = __ballot_sync(FULL, rlen != 0);
        if (any) {
          const int l0 = __ffs(any) - 1;
          firstY = (b << 5) + l0;
          firstX = __shfl_sync(FULL, xlo, l0);
        }
      }
      if (NW > 1) {
        // what the warp filling block b-1 writes: recompute its geometry from its rows
        wlo = 0; whi = 0;
        if (b > 0) {
          int poff, plen, pxlo, pxhi, pn;
          unsigned prl;
          load_corridor_rows(cv, b - 1, lane, H, poff, plen);
          row_span(poff, plen, ref_len, pxlo, pxhi, prl);
          const BlockGeom pg = block_geom(pxlo, pxhi, prl, lane, pn);
          wlo = pg.base - 31;
          whi = wlo + (pg.ngroups << 4);
        }
      }
      // team mode: block until the producer of block b-1 has flushed strip columns < x_end
      auto wait_for = [&](int x_end) {
        if (NW > 1 && b > 0) {
          const int need = x_end < whi ? x_end : whi;
          if (need > wlo) {
            const unsigned long long key = prog_key(b - 1, need);
            if (is0) {
              while (ld_prog(prev_prog) < key) __nanosleep(64);
              team_fence();  // acquire on the polling lane, BEFORE the barrier that releases the others
            }
            __syncwarp();
          }
        }
      };
      auto strip_rec = [&](int x) -> uint4 { return (x >= wlo && x < whi) ? ld_strip(strip + x) : EMPTY; };

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
      bool lIsD = false;
      int lRun = 0;            // RAW kernel: raw indelRun / direction of (x-1, y)
      uint32_t lDir = DIR_STOP;
      float kS = bestS;
      int kStep = -1;

      // stage chunk 0 (+ the diagonal neighbour of lane 0's first cell)
      wait_for(base + CHUNK);
      if (is0) dS = __uint_as_float(strip_rec(base - 1).x);
      uint4 pa = EMPTY, pb = EMPTY;
      uint32_t ra = 0, rb = 0;
      {
        const int x0 = base + lane;
        if (lane < CHUNK) {
          pa = strip_rec(x0);
          ra = __ldg(ref + x0);
        }
        if (CHUNK > 32) {
          pb = strip_rec(x0 + 32);
          rb = __ldg(ref + x0 + 32);
        }
      }

      for (int c = 0; c < nchunks; ++c) {
        pa.w = ra;  // the reference byte of the column rides in the record
        pb.w = rb;
        if (lane < CHUNK) io_s[2 * lane] = pa;
        if (CHUNK > 32) io_s[2 * (lane + 32)] = pb;
        __syncwarp();
#if !FILL_LANE0_LDS
        if (is31) {  // lane 31's shuffle sources carry the strip record lane 0 needs next
          const uint4 t = io_s[0];
          oS = __uint_as_float(t.x);
          oU = __uint_as_float(t.y);
          oP = t.z;
          oC = t.w;
        }
#endif
        if (c + 1 < nchunks) {  // fetch the next chunk while this one is computed
          const int x0 = base + (c + 1) * CHUNK + lane;
          wait_for(base + (c + 2) * CHUNK);
          if (lane < CHUNK) {
            pa = strip_rec(x0);
            ra = __ldg(ref + x0);
          }
          if (CHUNK > 32) {
            pb = strip_rec(x0 + 32);
            rb = __ldg(ref + x0 + 32);
          }
        }
        const int g_end = min(ngroups, (c + 1) * GPC);
        // A 16-step group in which every lane stays inside its corridor row needs no masking at all
        // (the common case away from the block's leading and trailing wavefront).
        auto do_group = [&](auto all_active_tag, int g) {
          constexpr bool ALL_ACTIVE = decltype(all_active_tag)::value;
          uint4* iop = io_s + ((g % GPC) << 5);  // [0] lane 0's input of this step, [1] lane 31's output
          uint32_t dw = 0;
          {
#pragma unroll 2  // 2 keeps the per-step predicates in registers; 4 makes ptxas spill them to a bit mask
            for (int k = 0; k < 16; ++k) {
              const int s = (g << 4) + k;
              uint4 v;
#if FILL_LANE0_LDS
              // lanes 1..31 take their upper neighbour's cell from lane t-1, lane 0 from the staged strip record
              v.x = __float_as_uint(__shfl_up_sync(FULL, oS, 1));
              v.y = __float_as_uint(__shfl_up_sync(FULL, oU, 1));
              v.z = __shfl_up_sync(FULL, oP, 1);
              v.w = __shfl_up_sync(FULL, oC, 1);
              if (is0) v = iop[0];
#else
              v.x = __float_as_uint(__shfl_sync(FULL, oS, src_lane));
              v.y = __float_as_uint(__shfl_sync(FULL, oU, src_lane));
              v.z = __shfl_sync(FULL, oP, src_lane);
              v.w = __shfl_sync(FULL, oC, src_lane);
#endif
              const float nS = __uint_as_float(v.x), nU = __uint_as_float(v.y);
              const uint32_t r = v.w;
              const bool act = ALL_ACTIVE ? true : ((unsigned)rel < rlen);
              float dg = __fadd_rn(dS, sc.mis);
              if (r == q) dg = __fadd_rn(dS, sc.mat);
              dS = nS;
              // Outside the corridor the cell must degenerate to {0, 0, STOP}: a NaN maximum makes every
              // equality below false, and fmaxf(NaN, 0) = 0 gives the score.
              const float m0 = fmaxf(fmaxf(fmaxf(lL, 0.0f), dg), nU);
              const float m = (ALL_ACTIVE || act) ? m0 : __int_as_float(0x7fffffff);
              const bool eL = (m == lL), eU = (m == nU), eG = (m == dg);
              const float S = ALL_ACTIVE ? m : fmaxf(m, 0.0f);  // STOP implies m == 0
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
                const bool lr = lIsD, ur = upRunF > 0.0f;  // the left cell's run is > 0 iff it is a deletion
                const bool X = (eU & ur) | eG;  // bitwise on purpose: straight PLOP3s, no short-circuit
                const bool pD = eL & (lr | !X);
                const bool pI = (!pD) & eU & (ur | !eG);
                code = eG ? DIR_DIAG : DIR_STOP;
                if (pI) asm volatile("mad.lo.u32 %0, %1, 0, 1;" : "=r"(code) : "r"(code));
                if (pD) asm volatile("mad.lo.u32 %0, %1, 0, 2;" : "=r"(code) : "r"(code));
                // at most one of the two run counters is alive after this cell
                // "zero, then an addition under the predicate" instead of add + select: the selects, compares and
                // min/max of this loop all go through the half-rate ALU pipe, which is what binds the kernel; a
                // predicated FADD runs on the FMA pipe (same for dg above and U / L below)
                float newD = 0.0f, newI = 0.0f;
                if (pD) asm volatile("add.rn.f32 %0, %1, 0f3F800000;" : "=f"(newD) : "f"(lRunF));
                if (pI) asm volatile("add.rn.f32 %0, %1, 0f3F800000;" : "=f"(newI) : "f"(upRunF));
                const float runF = __fadd_rn(newD, newI);
                const float pen = fminf(sc.ext_min, __fadd_rn(sc.gap_ext, __fmul_rn(runF, sc.decay)));
                // e = (S == 0) ? 0 : S + pen  as one exact fused op: S + pen * [S != 0]
                float nz;
                asm("set.ne.f32.f32 %0, %1, 0f00000000;" : "=f"(nz) : "f"(S));
                U = __fadd_rn(S, sc.open_read);
                L = __fadd_rn(S, sc.open_ref);
                if (pI) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(U) : "f"(pen), "f"(nz), "f"(S));
                if (pD) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(L) : "f"(pen), "f"(nz), "f"(S));
                oP = __float_as_uint(newI);
                lRunF = newD;
                lIsD = pD;
              }
              oS = S;
              oU = U;
              oC = r;
              lL = L;
              if (S > kS) {  // strict: first maximum in row-major order (:1165-1170)
                // S + 0 == S bit for bit (S >= +0); as predicated FMA-pipe instructions instead of two selects
                asm volatile("add.rn.f32 %0, %1, 0f00000000;" : "=f"(kS) : "f"(S));
                asm volatile("mad.lo.s32 %0, %1, 1, 0;" : "=r"(kStep) : "r"(s));
              }
              dw = __funnelshift_r(dw, code, 2);
#if FILL_LANE0_LDS
              if (is31) iop[1] = make_uint4(__float_as_uint(oS), __float_as_uint(oU), oP, oC);
#else
              if (is31) {
                // the record is exactly the four shuffle sources (w is rewritten when the chunk is staged)
                iop[1] = make_uint4(__float_as_uint(oS), __float_as_uint(oU), oP, oC);
                const uint4 t = iop[2];
                oS = __uint_as_float(t.x);
                oU = __uint_as_float(t.y);
                oP = t.z;
                oC = t.w;
              }
#endif
              iop += 2;
              if (!ALL_ACTIVE || RAW) ++rel;  // the RAW kernel needs the column for its tail rule
            }
          }
          if (ALL_ACTIVE && !RAW) rel += 16;
          // a lane whose 16 steps all lie outside its row has nothing the traceback will ever read: the block's
          // leading and trailing wavefront stays out of HBM (whole 32-byte sectors, the idle lanes are neighbours)
          if (ALL_ACTIVE || (rel > 0 && rel - 16 < (int)rlen)) dwp[(size_t)g * 32] = dw;
        };
        for (int g = c * GPC; g < g_end; ++g) {
          const bool all_active = __all_sync(FULL, rel >= 0 && rel + 15 < (int)rlen);
          if (all_active) do_group(std::true_type{}, g);
          else do_group(std::false_type{}, g);
        }
        __syncwarp();
        // flush lane 31's records of this chunk: columns [base - 31 + CHUNK*c, ...)
        {
          const int done = (g_end - c * GPC) << 4;  // steps executed in this chunk
          const int xo = base - 31 + c * CHUNK;
          if (lane < done) st_strip(strip + xo + lane, io_s[2 * lane + 1]);
          if (CHUNK > 32 && lane + 32 < done) st_strip(strip + xo + lane + 32, io_s[2 * (lane + 32) + 1]);
          if (NW > 1) {
            team_fence();
            __syncwarp();
            if (is0) st_prog(my_prog, prog_key(b, xo + done));
          }
        }
        __syncwarp();
      }

      if (kStep >= 0) {
        bestS = kS;
        bestY = y;
        bestX = base + kStep - lane;
      }
      if (NW == 1) {
        wlo = base - 31;
        whi = base - 31 + (ngroups << 4);
      } else if (is0) {
        st_prog(my_prog, prog_key(b, (1 << 30)));  // block complete
      }
      __syncwarp();
    }
    if (NW > 1) {
      if (is0) st_prog(my_prog, ~0ull);  // also after an arena overflow: never leave a consumer waiting
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
    if (NW > 1) {
      if (is0) {
        TeamBest tb;
        tb.S = bestS; tb.x = bestX; tb.y = bestY; tb.firstX = firstX; tb.firstY = firstY;
        tb.status = status; tb.cells = cells;
        s_best[tw] = tb;
      }
      __syncthreads();
      if (wib == 0) {
        for (int t = 1; t < NW; ++t) {
          const TeamBest tb = s_best[t];
          const bool take = (tb.S > bestS) || (tb.S == bestS && (tb.y < bestY || (tb.y == bestY && tb.x < bestX)));
          if (take) { bestS = tb.S; bestY = tb.y; bestX = tb.x; }
          cells += tb.cells;
          if (tb.status != ST_OK) status = tb.status;
          if (tb.firstY >= 0 && (firstY < 0 || tb.firstY < firstY)) { firstY = tb.firstY; firstX = tb.firstX; }
        }
      }
    }
    if (bestS == 0.0f) {  // no positive score anywhere: the first visited cell stands (or nothing was visited)
      bestS = firstY >= 0 ? 0.0f : -1.0f;
      bestX = firstY >= 0 ? firstX : 0;
      bestY = firstY >= 0 ? firstY : 0;
    }
    if (is0 && (NW == 1 || wib == 0)) {
      FillOut o;
      o.best_score = bestS;
      o.best_x = bestX;
      o.best_y = bestY;
      o.status = status;
      o.cells = cells;
      p.out[ai] = o;
    }
  }
  if (!PERSIST) {
    __syncthreads();
    if (threadIdx.x == 0) atomicAnd(p.sm_slots + smid, ~slot_bit);
  }
}

}  // namespace

namespace {
template <bool RAW, int NW>
void launch_fill(const FillParams& p, int grid, cudaStream_t stream) {
  const int threads = (NW == 1 ? FILL_WARPS_PER_CTA : NW) * 32;
  if (p.sm_slots) convex_fill_kernel<RAW, NW, false><<<grid, threads, 0, stream>>>(p);
  else convex_fill_kernel<RAW, NW, true><<<grid, threads, 0, stream>>>(p);
}
}  // namespace

// p.sm_slots != nullptr selects the short-lived-CTA instantiation (one problem per warp / team, strips by SM slot)
cudaError_t launch_convex_fill(const FillParams& p, bool raw, bool team, int grid, cudaStream_t stream) {
  if (raw) {
    if (team) launch_fill<true, FILL_WARPS_PER_CTA>(p, grid, stream);
    else launch_fill<true, 1>(p, grid, stream);
  } else {
    if (team) launch_fill<false, FILL_WARPS_PER_CTA>(p, grid, stream);
    else launch_fill<false, 1>(p, grid, stream);
  }
  return cudaGetLastError();
}

cudaError_t launch_convex_fill_big(const FillParams& p, bool raw, int grid, cudaStream_t stream) {
  const int threads = FILL_BIG_TEAM * 32;
  if (raw) convex_fill_kernel<true, FILL_BIG_TEAM, true><<<grid, threads, 0, stream>>>(p);
  else convex_fill_kernel<false, FILL_BIG_TEAM, true><<<grid, threads, 0, stream>>>(p);
  return cudaGetLastError();
}

int fill_max_ctas_per_sm(bool raw, bool team) {
  int n = 0;
  const int threads = FILL_WARPS_PER_CTA * 32;
  if (raw) {
    if (team) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, convex_fill_kernel<true, FILL_WARPS_PER_CTA, true>, threads, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, convex_fill_kernel<true, 1, true>, threads, 0);
  } else {
    if (team) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, convex_fill_kernel<false, FILL_WARPS_PER_CTA, true>, threads, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, convex_fill_kernel<false, 1, true>, threads, 0);
  }
  return n;
}

}  // namespace nb
