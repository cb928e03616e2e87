// ngmlr_b200/csrc/ngm_files.cpp -- ngmlr's on-disk caches, byte-compatible, from C.
//
//   <ref>-enc.2.ngm             _SequenceProvider::writeEncRefToFile / readEncRefFromFile
//                               (src/SequenceProvider.cpp:207-272): uint cookie 0x74656, uint refCount,
//                               uloc binRefIndex (= 2 x used bytes), uloc encRefSize (allocated bytes),
//                               RefIdx[refCount] (128 bytes each, src/SequenceProvider.h:56-63), binRef bytes.
//   <ref>-ht-<k>-<skip>.2.ngm   CompactPrefixTable::saveToFile / readFromFile (src/PrefixTable.cpp:534-630):
//                               uint cookie 0x1701E, prefix length, ref skip, unit count, index size; per unit
//                               uint cRefTableLen, Index[index size] (5 bytes each), Location[cRefTableLen],
//                               uloc Offset; uint signature = sum of the five header words.
// The arrays are the ones ngmlr_b200_cs_get_index / cs_set_index / cs_set_reference exchange, so an index built on
// the device (ngmlr_b200_cs_build_index) becomes the cache file an unmodified ngmlr starts from, and the caches
// ngmlr already has on disk feed the device pipeline. Same bytes as ngmlr_b200/ngmfiles.py (the tests' writer) and,
// through it, as the files the unmodified reference writes (tests/test_cs_oracle.py, tests/test_gpu_index.py).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/ngmlr_b200.h"

namespace {

constexpr uint32_t REF_ENC_COOKIE = 0x74656;
constexpr uint32_t REF_TAB_COOKIE = 0x1701E;

#pragma pack(push, 1)
struct RefIdxRec {  // src/SequenceProvider.h:56-63 (natural alignment gives the same 128 bytes)
  uint32_t SeqId;
  uint32_t Flags;
  uint64_t SeqStart;
  uint32_t SeqLen;
  uint32_t NameLen;
  char name[100];
  char pad[4];
};
#pragma pack(pop)
static_assert(sizeof(RefIdxRec) == 128, "RefIdx is 128 bytes in the reference");

struct File {
  FILE* f;
  explicit File(const char* path, const char* mode) : f(fopen(path, mode)) {}
  ~File() {
    if (f) fclose(f);
  }
  bool put(const void* p, size_t n) { return n == 0 || fwrite(p, 1, n, f) == n; }
  bool get(void* p, size_t n) { return n == 0 || fread(p, 1, n, f) == n; }
};

}  // namespace

extern "C" {

int ngmlr_b200_ngm_write_index(const char* path, int k, int kmer_skip, const void* packed_index, uint32_t index_len,
                               const uint32_t* positions, uint32_t n_positions, uint64_t unit_offset) {
  if (!path || !packed_index || (n_positions && !positions) || k <= 0) return -1;
  File fp(path, "wb");
  if (!fp.f) return -2;
  const uint32_t head[5] = {REF_TAB_COOKIE, (uint32_t)k, (uint32_t)kmer_skip, 1u, index_len};
  const uint32_t sig = head[0] + head[1] + head[2] + head[3] + head[4];
  const bool ok = fp.put(head, sizeof head) && fp.put(&n_positions, 4) && fp.put(packed_index, (size_t)index_len * 5) &&
                  fp.put(positions, (size_t)n_positions * 4) && fp.put(&unit_offset, 8) && fp.put(&sig, 4);
  return ok ? 0 : -3;
}

int ngmlr_b200_ngm_read_index(const char* path, int32_t* k, int32_t* kmer_skip, uint32_t* index_len,
                              uint32_t* n_positions, uint64_t* unit_offset, void* packed_index, uint32_t* positions) {
  if (!path) return -1;
  File fp(path, "rb");
  if (!fp.f) return -2;
  uint32_t head[5], n_pos = 0;
  if (!fp.get(head, sizeof head) || head[0] != REF_TAB_COOKIE) return -3;
  if (head[3] != 1u) return -4;  // several table units (more than 4 G positions): the device search takes one
  if (!fp.get(&n_pos, 4)) return -3;
  // signature at the end of the file: the reference rebuilds the table when it does not match
  const long body = 24 + (long)head[4] * 5 + (long)n_pos * 4 + 8;
  uint32_t sig = 0;
  if (fseek(fp.f, body, SEEK_SET) != 0 || !fp.get(&sig, 4) || sig != head[0] + head[1] + head[2] + head[3] + head[4])
    return -5;
  if (k) *k = (int32_t)head[1];
  if (kmer_skip) *kmer_skip = (int32_t)head[2];
  if (index_len) *index_len = head[4];
  if (n_positions) *n_positions = n_pos;
  if (fseek(fp.f, 24, SEEK_SET) != 0) return -3;
  if (packed_index) {
    if (!fp.get(packed_index, (size_t)head[4] * 5)) return -3;
  } else if (fseek(fp.f, (long)head[4] * 5, SEEK_CUR) != 0) {
    return -3;
  }
  if (positions) {
    if (!fp.get(positions, (size_t)n_pos * 4)) return -3;
  } else if (fseek(fp.f, (long)n_pos * 4, SEEK_CUR) != 0) {
    return -3;
  }
  uint64_t off = 0;
  if (!fp.get(&off, 8)) return -3;
  if (unit_offset) *unit_offset = off;
  return 0;
}

int ngmlr_b200_ngm_write_reference(const char* path, const uint8_t* bin_ref, uint64_t used_bytes, uint64_t alloc_bytes,
                                   int n_refs, const uint64_t* seq_start, const uint32_t* seq_len,
                                   const char* const* names) {
  if (!path || (used_bytes && !bin_ref) || n_refs < 0 || (n_refs && (!seq_start || !seq_len))) return -1;
  if (alloc_bytes < used_bytes) alloc_bytes = used_bytes;
  File fp(path, "wb");
  if (!fp.f) return -2;
  const uint32_t cookie = REF_ENC_COOKIE, count = (uint32_t)n_refs;
  const uint64_t bin_ref_index = 2 * used_bytes;
  bool ok = fp.put(&cookie, 4) && fp.put(&count, 4) && fp.put(&bin_ref_index, 8) && fp.put(&alloc_bytes, 8);
  for (int i = 0; ok && i < n_refs; ++i) {
    RefIdxRec r;
    memset(&r, 0, sizeof r);
    r.SeqId = (uint32_t)i;
    r.SeqStart = seq_start[i];
    r.SeqLen = seq_len[i];
    char fallback[32];
    const char* nm = names ? names[i] : nullptr;
    if (!nm) {
      snprintf(fallback, sizeof fallback, "c%d", i);
      nm = fallback;
    }
    size_t nl = strlen(nm);
    if (nl > 100) nl = 100;
    memcpy(r.name, nm, nl);
    r.NameLen = (uint32_t)nl;
    ok = fp.put(&r, sizeof r);
  }
  ok = ok && fp.put(bin_ref, (size_t)used_bytes);
  // the reference writes its whole allocation (uninitialised behind the used part); zeros here
  std::vector<char> zeros((size_t)std::min<uint64_t>(alloc_bytes - used_bytes, 1u << 20), 0);
  for (uint64_t left = alloc_bytes - used_bytes; ok && left;) {
    const size_t n = (size_t)std::min<uint64_t>(left, zeros.size());
    ok = fp.put(zeros.data(), n);
    left -= n;
  }
  return ok ? 0 : -3;
}

int ngmlr_b200_ngm_read_reference(const char* path, uint32_t* n_refs, uint64_t* used_bytes, uint64_t* alloc_bytes,
                                  uint64_t* seq_start, uint32_t* seq_len, char* names, uint8_t* bin_ref) {
  if (!path) return -1;
  File fp(path, "rb");
  if (!fp.f) return -2;
  uint32_t cookie = 0, count = 0;
  uint64_t bin_ref_index = 0, enc_size = 0;
  if (!fp.get(&cookie, 4) || !fp.get(&count, 4) || !fp.get(&bin_ref_index, 8) || !fp.get(&enc_size, 8)) return -3;
  if (cookie != REF_ENC_COOKIE) return -3;
  if (bin_ref_index % 2 || bin_ref_index / 2 > enc_size) return -5;
  if (n_refs) *n_refs = count;
  if (used_bytes) *used_bytes = bin_ref_index / 2;
  if (alloc_bytes) *alloc_bytes = enc_size;
  for (uint32_t i = 0; i < count; ++i) {
    RefIdxRec r;
    if (!fp.get(&r, sizeof r)) return -3;
    if (seq_start) seq_start[i] = r.SeqStart;
    if (seq_len) seq_len[i] = r.SeqLen;
    if (names) {
      const uint32_t nl = r.NameLen > 100 ? 100 : r.NameLen;
      memcpy(names + (size_t)i * 101, r.name, nl);
      names[(size_t)i * 101 + nl] = 0;
    }
  }
  if (bin_ref && !fp.get(bin_ref, (size_t)(bin_ref_index / 2))) return -3;
  return 0;
}

}  // extern "C"
