// ngmlr_b200/csrc/sw_score.cu -- score-only local alignment of sub-read x candidate pairs (sm_100a).
//
// Replaces StrippedSW::BatchScore / SingleScore (src/StrippedSW.cpp:118-202) and the part of the
// vendored SSW library they reach: ssw_init -> qP_word, ssw_align(flag=0) -> sw_sse2_word
// (lib/Complete-Striped-Smith-Waterman-Library/src/ssw.c:964-989, 342-364, 997-1054, 366-538).
//
// Semantics kept bit-for-bit:
//   * both sequences are scored INCLUDING their terminating NUL (lengths are strlen+1); the NUL
//     and every non-ACGT byte map to code 4, which scores 0 against everything;
//   * +1 match / -1 mismatch; gap open = gap extend = 255 (int32 -1 narrowed to uint8_t);
//   * H = max(diag + s, E, F) with signed-saturating 16-bit adds; E/F updates use unsigned
//     saturating subtraction (clamp at 0); result = max H as uint16 -> float.
//
// Mapping: one warp per pair. Lane t owns SW_ROWS consecutive query rows; the warp sweeps the
// reference columns as a wavefront (lane t is one column behind lane t-1), passing the H and F of
// its last row down with __shfl_up. Queries longer than 32*SW_ROWS rows are processed in several
// passes; the bottom row of a pass is carried to the next through a per-warp strip in global
// memory. 32-bit integer lanes emulate the int16 saturation (values never exceed 32767).
#include <cuda_runtime.h>

#include "device_types.h"
#include "kernels.h"

namespace nb {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int SW_ROWS = 9;  // 32 * 9 = 288 >= 257 = 256-bp sub-read + NUL: one pass for the hot case
constexpr int SW_GAP = 255;
constexpr int SW_PAIRS = 5;  // packed path: 5 registers x 2 rows per lane = 320 rows
constexpr int SW_WARPS_PER_CTA = 4;

__device__ __forceinline__ int nt_code(uint32_t c) {
  // nt_table, src/StrippedSW.cpp:111-116
  c &= 0xdfu;  // fold case
  return c == 'A' ? 0 : (c == 'C' ? 1 : (c == 'G' ? 2 : (c == 'T' ? 3 : 4)));
}

// DecodeRefSequence(sequence, 0, position, bufferLength) as a random-access function
// (src/SequenceProvider.cpp:567-625): character i of the window, as an nt_table code.
struct GenomeWindow {
  const uint8_t* enc;
  unsigned long long pos, concat_len;
  int odd;        // position & 1: one extra leading base
  int pairs2;     // 2 * ceil(len / 2) characters decoded pairwise
  int x_at;       // index (after the odd base) that is overwritten by 'x' when len is odd, else -1
  int str_len;    // strlen of the decoded window
  bool invalid;   // position >= concat_len: the caller fills the buffer with 'N'
  __device__ void init(const uint8_t* e, unsigned long long p, unsigned long long cl, int buffer_len) {
    enc = e; pos = p; concat_len = cl;
    invalid = p >= cl;
    unsigned long long len = (unsigned long long)(buffer_len - 2), end = 0;
    if (!invalid && p + len > cl) { end = p + len - cl; len -= end; }
    odd = (int)(p & 1ull);
    pairs2 = (int)(((len + 1) / 2) * 2);
    x_at = (len & 1ull) ? pairs2 - 1 : -1;
    str_len = invalid ? buffer_len : odd + pairs2 + (int)end;
  }
  __device__ __forceinline__ int code(int i) const {
    if (invalid) return 4;                   // memset(buf, 'N', refMaxLen), src/ScoreBuffer.cpp:114
    if (i >= str_len) return 4;              // the terminating NUL (scored as N)
    const int k = i - odd;
    if (k >= pairs2 || k == x_at) return 4;  // 'x' padding
    const unsigned long long b = pos + (unsigned long long)i;  // base index in the concatenated genome
    const uint32_t byte = enc[b >> 1];
    const uint32_t c4 = (b & 1ull) ? (byte & 0xFu) : (byte >> 4);
    // enc4 A0 T1 G2 C3 N4 -> nt_table A0 C1 G2 T3 N4
    return c4 == 0 ? 0 : (c4 == 1 ? 3 : (c4 == 2 ? 2 : (c4 == 3 ? 1 : 4)));
  }
};

__device__ __forceinline__ uint32_t cpl(uint32_t c) {  // src/MappedRead.cpp:35-46
  return c == 'A' ? 'T' : (c == 'T' ? 'A' : (c == 'C' ? 'G' : (c == 'G' ? 'C' : c)));
}

template <bool GATHER>
__global__ void __launch_bounds__(SW_WARPS_PER_CTA * 32) sw_score_kernel(const SwParams p) {
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * SW_WARPS_PER_CTA + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * SW_WARPS_PER_CTA;
  int2* strip = reinterpret_cast<int2*>(p.scratch) + (size_t)warp_global * p.scratch_stride;

  for (int pair = warp_global; pair < p.n; pair += nwarps) {
    GenomeWindow gw;
    int rlen_g = 0;
    if (GATHER) {
      gw.init(p.enc, p.win_pos[pair], p.concat_len, p.win_len);
      rlen_g = gw.str_len + 1;  // strlen + 1: the NUL is scored
    }
    const bool q_rev = GATHER && p.rev[pair];
    const int qlen = p.qry_len[pair], rlen = GATHER ? rlen_g : p.ref_len[pair];
    if (qlen >= 100000 || rlen >= 100000) {  // maxSeqLen, src/StrippedSW.h:88
      if (lane == 0) p.out[pair] = -1.0f;
      continue;
    }
    const uint8_t* __restrict__ ref = GATHER ? p.seq : p.seq + p.ref_off[pair];
    const uint8_t* __restrict__ qry = p.seq + p.qry_off[pair];
    int best = 0;
    // ---- hot case: at most 320 query characters (256-bp sub-read + NUL) ----
    // A gap costs 255 per base, so a gapped path beats its best ungapped piece only if it gains
    // more than 255 before AND after the gap: impossible with fewer than 512 query characters.
    // Then E and F never influence the maximum and the recurrence is H = max(0, diag + s).
    // Two rows per register (16-bit halves) with Blackwell's packed integer ops:
    // VIMNMX.U16x2 (character mismatch), VIADDMNMX.S16x2.RELU (max(diag + s, 0)), VIMNMX.S16x2.
    if (qlen <= 32 * 2 * SW_PAIRS) {
      const int row0 = lane * 2 * SW_PAIRS;
      uint32_t qc2[SW_PAIRS], vq2[SW_PAIRS], H2[SW_PAIRS];
#pragma unroll
      for (int pq = 0; pq < SW_PAIRS; ++pq) {
        uint32_t codes[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int row = row0 + 2 * pq + h;
          int cq = 4;  // NUL / beyond the query: scores 0 against everything
          if (row < qlen - 1) cq = q_rev ? nt_code(cpl(qry[qlen - 2 - row])) : nt_code(qry[row]);
          codes[h] = (uint32_t)cq;
        }
        qc2[pq] = codes[0] | (codes[1] << 16);
        vq2[pq] = (codes[0] == 4u ? 0u : 0xffffu) | (codes[1] == 4u ? 0u : 0xffff0000u);
        H2[pq] = 0u;
      }
      uint32_t best2 = 0u, diag_top = 0u;
      const int nsteps = rlen + 31;
      for (int s = 0; s < nsteps; ++s) {
        const int c = s - lane;
        uint32_t up2 = __shfl_up_sync(FULL, H2[SW_PAIRS - 1], 1);
        if (lane == 0) up2 = 0u;
        if (c >= 0 && c < rlen) {
          const uint32_t rc = (uint32_t)(GATHER ? gw.code(c) : nt_code(ref[c]));
          const uint32_t rc2 = rc * 0x00010001u;
          const bool col_n = rc == 4u;
          uint32_t prev = diag_top;  // H of the row above pair 0 at column c-1 sits in its high half
#pragma unroll
          for (int pq = 0; pq < SW_PAIRS; ++pq) {
            const uint32_t old = H2[pq];
            const uint32_t diag = __byte_perm(prev, old, 0x5432);     // (H[2p-1], H[2p]) of column c-1
            const uint32_t ne = __vminu2(qc2[pq] ^ rc2, 0x00010001u);  // 1 per half where the bases differ
            const uint32_t mism = ne * 0xffffu;                        // 0xffff per differing half
            const uint32_t valid = col_n ? 0u : vq2[pq];
            const uint32_t sub = (mism | 0x00010001u) & valid;         // +1 / -1, 0 against N
            const uint32_t h = __viaddmax_s16x2_relu(diag, sub, 0u);
            H2[pq] = h;
            best2 = __vmaxs2(best2, h);
            prev = old;
          }
          diag_top = up2;
        }
      }
      best = max((int)(best2 & 0xffffu), (int)(best2 >> 16));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(FULL, best, o));
      if (lane == 0) p.out[pair] = (float)(best & 0xffff);
      continue;
    }
    const int rows_per_pass = 32 * SW_ROWS;
    for (int row_base = 0; row_base < qlen; row_base += rows_per_pass) {
      const bool first_pass = row_base == 0;
      const bool last_pass = row_base + rows_per_pass >= qlen;
      const int row0 = row_base + lane * SW_ROWS;
      int qc[SW_ROWS], H[SW_ROWS], E[SW_ROWS];
#pragma unroll
      for (int r = 0; r < SW_ROWS; ++r) {
        const int row = row0 + r;
        // the NUL at qlen-1 is part of the query and maps to 4; rows >= qlen do not exist (-1)
        if (q_rev) {  // RevSeq[j] = cpl(Seq[len - 1 - j]), then the NUL
          const int L = qlen - 1;
          qc[r] = row < L ? nt_code(cpl(qry[L - 1 - row])) : (row < qlen ? 4 : -1);
        } else {
          qc[r] = row < qlen ? nt_code(qry[row]) : -1;
        }
        H[r] = 0;
        E[r] = 0;
      }
      int outH = 0, outF = 0, diagTop = 0;
      const int nsteps = rlen + 31;
      for (int s = 0; s < nsteps; ++s) {
        const int c = s - lane;
        int upH = __shfl_up_sync(FULL, outH, 1);
        int upF = __shfl_up_sync(FULL, outF, 1);
        const bool in = c >= 0 && c < rlen;
        if (lane == 0) {
          upH = 0;
          upF = 0;
          if (!first_pass && in) {
            const int2 v = __ldcg(strip + c);
            upH = v.x;
            upF = v.y;
          }
        }
        if (in) {
          const int rc = GATHER ? gw.code(c) : nt_code(ref[c]);
          int diag = diagTop;
          diagTop = upH;
          int F = upF;
#pragma unroll
          for (int r = 0; r < SW_ROWS; ++r) {
            const int hOld = H[r];
            const int sub = ((rc | qc[r]) & 4) ? 0 : (rc == qc[r] ? 1 : -1);
            int h = min(diag + sub, 32767);      // _mm_adds_epi16
            h = max(h, E[r]);
            h = max(h, F);
            if (qc[r] < 0) h = 0;                // row beyond the query
            best = max(best, h);
            H[r] = h;
            const int hg = max(h - SW_GAP, 0);   // _mm_subs_epu16
            E[r] = max(max(E[r] - SW_GAP, 0), hg);
            F = max(max(F - SW_GAP, 0), hg);
            diag = hOld;
          }
          outH = H[SW_ROWS - 1];
          outF = F;
          if (lane == 31 && !last_pass) __stcg(strip + c, make_int2(outH, outF));
        }
      }
      __syncwarp();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(FULL, best, o));
    if (lane == 0) p.out[pair] = (float)(best & 0xffff);
  }
}

}  // namespace

cudaError_t launch_sw_score(const SwParams& p, int grid, cudaStream_t stream) {
  if (p.n <= 0) return cudaSuccess;
  sw_score_kernel<false><<<grid, SW_WARPS_PER_CTA * 32, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_sw_score_gather(const SwParams& p, int grid, cudaStream_t stream) {
  if (p.n <= 0) return cudaSuccess;
  sw_score_kernel<true><<<grid, SW_WARPS_PER_CTA * 32, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace nb
