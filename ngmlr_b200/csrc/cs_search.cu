// ngmlr_b200/csrc/cs_search.cu -- k-mer candidate search (stage 0) for sm_100a.
//
// Replaces, per (sub-)read, CS::RunRead's search (src/CS.cpp:324-398): CS::PrefixIteration
// (src/CSstatic.cpp:23-73) -> CS::PrefixSearch (src/CS.cpp:57-96) ->
// CompactPrefixTable::GetRefEntry (src/PrefixTable.cpp:476-532, revComp :70-88) ->
// CS::AddLocationStd (src/CS.cpp:98-149) -> CS::CollectResultsStd (src/CS.cpp:217-269).
//
// The vote is order dependent (the acceptance threshold 0.8 x max-so-far runs along with the
// hits, and candidates are emitted in the order in which their bin first crossed it), so each
// (sub-)read is processed by ONE thread that replays the reference's sequence of hits exactly;
// parallelism comes from the hundreds of thousands of independent sub-reads of a batch (a 2048-
// thread SM keeps 2048 of these latency-bound walks in flight). The index (5-byte Index records
// unpacked to tab/used arrays + uint32 position lists) and the open-addressing vote tables live in
// HBM (the used() flags as an L2-resident bitmap); a first pass counts each read's hits so that its table can be sized (the reference instead
// restarts with a larger table on overflow -- results do not depend on the table size).
//
// Assumes, like the reference's 1000-N leading spacer guarantees, position >= offset-in-read.
#include <cuda_runtime.h>

#include "device_types.h"
#include "kernels.h"

namespace nb {

namespace {

__device__ __forceinline__ uint32_t rev_comp(uint32_t prefix, int k, uint32_t mask) {
  // complement = xor 10b per base (A0 C1 T2 G3), then reverse the 2-bit groups
  uint32_t c = (prefix ^ 0xAAAAAAAAu) & mask;
  c = __brev(c);                                            // bit reversal also swaps bits in a pair
  c = ((c >> 1) & 0x55555555u) | ((c & 0x55555555u) << 1);  // swap them back
  return c >> (32 - 2 * k);
}

struct VoteEntry {
  uint32_t key;    // bin = (loc - correction) >> bin_shift   (CSTableEntry::m_Location is a uint)
  uint32_t state;  // bit 0: used, bit 1: already listed
  float f, r;
};

template <bool COUNT_ONLY>
__global__ void __launch_bounds__(128) cs_search_kernel(const CsParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const uint8_t* __restrict__ seq = p.seq + p.seq_off[i];
  const int len = p.seq_len[i];
  const int k = p.k;
  const uint32_t mask = (k == 16) ? 0xffffffffu : ((1u << (2 * k)) - 1u);

  VoteEntry* tab = nullptr;
  uint32_t cap_mask = 0;
  uint32_t* order = nullptr;
  int n_order = 0;
  float max_hits = 0.0f, thresh = 0.0f;
  unsigned long long hits = 0;
  if (!COUNT_ONLY) {
    tab = reinterpret_cast<VoteEntry*>(p.tables) + p.table_off[i];
    cap_mask = p.table_cap[i] - 1u;
    order = p.order + p.order_off[i];
  }

  auto vote = [&](uint32_t bin, bool reverse) {
    uint32_t s = (bin * 2654435761u) >> 7;
    s &= cap_mask;
    while ((tab[s].state & 1u) && tab[s].key != bin) s = (s + 1u) & cap_mask;
    VoteEntry e = tab[s];
    float score;
    if (!(e.state & 1u)) {
      e.key = bin;
      e.state = 1u;
      e.f = reverse ? 0.0f : 1.0f;
      e.r = reverse ? 1.0f : 0.0f;
      score = 1.0f;
    } else if (reverse) {
      e.r += 1.0f;
      score = e.r;
    } else {
      e.f += 1.0f;
      score = e.f;
    }
    if (score > max_hits) {  // (:136-141)
      max_hits = score;
      thresh = __fmul_rn(max_hits, p.sensitivity);
    }
    if (!(e.state & 2u) && score >= thresh) {  // (:143-147)
      e.state |= 2u;
      order[n_order++] = s;
    }
    tab[s] = e;
  };

  auto kmer = [&](uint32_t prefix, int pos) {
    // forward list, then the list of the reverse-complement k-mer (GetRefEntry)
    if ((p.used_bits[prefix >> 5] >> (prefix & 31u)) & 1u) {
      const uint32_t start = p.tab[prefix] - 1u, n = p.tab[prefix + 1] - 1u - start;
      if (COUNT_ONLY) {
        hits += n;
      } else {
        for (uint32_t j = 0; j < n; ++j) {
          const unsigned long long loc = (unsigned long long)p.pos[start + j] + p.unit_offset;
          vote((uint32_t)((loc - (unsigned long long)pos) >> p.bin_shift), false);
        }
      }
    }
    const uint32_t rc = rev_comp(prefix, k, mask);
    if ((p.used_bits[rc >> 5] >> (rc & 31u)) & 1u) {
      const uint32_t start = p.tab[rc] - 1u, n = p.tab[rc + 1] - 1u - start;
      if (COUNT_ONLY) {
        hits += n;
      } else {
        const unsigned long long corr = (unsigned long long)(len - (pos + k));
        for (uint32_t j = 0; j < n; ++j) {
          const unsigned long long loc = (unsigned long long)p.pos[start + j] + p.unit_offset;
          vote((uint32_t)((loc - corr) >> p.bin_shift), true);
        }
      }
    }
  };

  // ---- CS::PrefixIteration with prefixskip = 0, tail recursion as a loop ----
  {
    int cur = 0, length = len;
    for (;;) {
      if (length < k) break;
      if (seq[cur] == 'N') {
        int n_skip = 1;
        while (cur + n_skip < len && seq[cur + n_skip] == 'N') ++n_skip;
        cur += n_skip;
        if (n_skip >= length - k) break;
        length -= n_skip;
      }
      uint32_t prefix = 0;
      int j = 0;
      bool restart = false;
      for (; j < k - 1; ++j) {
        const uint32_t c = seq[cur + j];
        if (c == 'N') { restart = true; break; }
        prefix = (prefix << 2) | ((c >> 1) & 3u);
      }
      if (!restart) {
        for (j = k - 1; j < length; ++j) {
          const uint32_t c = seq[cur + j];
          if (c == 'N') { restart = true; break; }
          prefix = ((prefix << 2) | ((c >> 1) & 3u)) & mask;
          kmer(prefix, cur + j + 1 - k);
        }
      }
      if (!restart) break;
      cur += j + 1;
      length -= j + 1;
    }
  }

  if (COUNT_ONLY) {
    p.hits[i] = hits;
    return;
  }
  // ---- CS::CollectResultsStd ----
  const float thr = fmaxf(p.min_kmer_hits, thresh);
  CsCandidate* out = p.out + p.out_off[i];
  int n = 0;
  const unsigned long long half = p.bin_shift > 0 ? (1ull << (p.bin_shift - 1)) : 0ull;
  for (int j = 0; j < n_order; ++j) {
    const VoteEntry e = tab[order[j]];
    const unsigned long long loc = ((unsigned long long)e.key << p.bin_shift) + half;  // ResolveBin
    if (e.f >= thr) {
      out[n].loc = loc;
      out[n].score = e.f;
      out[n].reverse = 0;
      ++n;
    }
    if (e.r >= thr) {
      out[n].loc = loc;
      out[n].score = e.r;
      out[n].reverse = 1;
      ++n;
    }
  }
  p.out_count[i] = n;
  p.max_hits[i] = max_hits;
}

// Unpack the reference's 5-byte Index records {uint m_TabIndex; char m_RevCompIndex}
// (#pragma pack(1), src/PrefixTable.h:17-35) into an aligned uint32 array plus a bitmap of
// Index::used() (m_RevCompIndex != 0). The bitmap of a 13-mer index is 8 MB and stays in L2, so the
// ~75 % of lookups that hit an unused prefix never go to HBM.
__global__ void unpack_index_kernel(const uint8_t* __restrict__ packed, uint32_t n, uint32_t* __restrict__ tab,
                                    uint32_t* __restrict__ used_bits) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;  // blockDim is a multiple of 32
  bool used = false;
  if (i < n) {
    const uint8_t* r = packed + (size_t)i * 5;
    tab[i] = (uint32_t)r[0] | ((uint32_t)r[1] << 8) | ((uint32_t)r[2] << 16) | ((uint32_t)r[3] << 24);
    used = r[4] != 0;
  }
  const uint32_t word = __ballot_sync(0xffffffffu, used);
  if ((threadIdx.x & 31) == 0) used_bits[i >> 5] = word;
}

}  // namespace

cudaError_t launch_cs_search(const CsParams& p, bool count_only, cudaStream_t stream) {
  if (p.n <= 0) return cudaSuccess;
  const int grid = (p.n + 127) / 128;
  if (count_only)
    cs_search_kernel<true><<<grid, 128, 0, stream>>>(p);
  else
    cs_search_kernel<false><<<grid, 128, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_unpack_index(const uint8_t* packed, uint32_t n, uint32_t* tab, uint32_t* used_bits,
                                cudaStream_t stream) {
  if (!n) return cudaSuccess;
  unpack_index_kernel<<<(n + 255) / 256, 256, 0, stream>>>(packed, n, tab, used_bits);
  return cudaGetLastError();
}

}  // namespace nb
