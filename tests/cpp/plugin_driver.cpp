// tests/cpp/plugin_driver.cpp -- exercises libngmlr_b200.so exactly the way ngmlr would: dlopen,
// CreateAlignment(gpu_id), then calls through the IAlignment vtable (SingleAlign with
// CorridorLine[], BatchAlign with NgmlrB200BatchAlignArgs, BatchScore, SingleScore) with
// caller-allocated Align buffers (src/AlignmentBuffer.cpp:271-278). Problems are read from a
// simple text file written by tests/test_gpu_plugin.py, results are printed one line per problem
// for comparison with the oracle.
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "ngmlr_b200_ialignment.h"

struct Problem {
  std::string ref, qry;
  std::vector<CorridorLine> lines;
  int qs, qe;
};

static void print_align(const char* tag, int i, int ret, const Align& a) {
  unsigned bits, ibits;
  memcpy(&bits, &a.Score, 4);
  memcpy(&ibits, &a.Identity, 4);
  if (ret < 0) {
    printf("%s %d ret=-1 score_bits=%u\n", tag, i, bits);
    return;
  }
  printf("%s %d ret=%d score_bits=%u identity_bits=%u pos=%d qstart=%d qend=%d nm=%d alen=%d ops=%d sv=%d "
         "first=%d,%d last=%d,%d cigar=%s md=%s\n",
         tag, i, ret, bits, ibits, a.PositionOffset, a.QStart, a.QEnd, a.NM, a.alignmentLength,
         a.cigarOpCount, a.svType, a.firstPosition.refPosition, a.firstPosition.readPosition,
         a.lastPosition.refPosition, a.lastPosition.readPosition, a.pBuffer1, a.pBuffer2);
}

static void alloc_align(Align& a, size_t readLength) {
  a.maxBufferLength = (int)readLength * 4 + 8;
  a.maxMdBufferLength = (int)readLength * 4 + 8;
  a.pBuffer1 = new char[a.maxBufferLength];
  a.pBuffer2 = new char[a.maxMdBufferLength];
  a.pBuffer1[0] = a.pBuffer2[0] = '\0';
  a.nmPerPostionLength = ((int)readLength + 1) * 2;
  a.nmPerPosition = new PositionNM[a.nmPerPostionLength];
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  void* h = dlopen(argv[1], RTLD_NOW);
  if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 3; }
  pfCreateAlignment create = (pfCreateAlignment)dlsym(h, "CreateAlignment");
  pfDeleteAlignment destroy = (pfDeleteAlignment)dlsym(h, "DeleteAlignment");
  if (!create || !destroy) return 4;
  IAlignment* al = create(0);
  if (!al) { fprintf(stderr, "CreateAlignment returned NULL (no device?)\n"); return 5; }

  FILE* f = fopen(argv[2], "r");
  int n = 0;
  if (fscanf(f, "%d", &n) != 1) return 6;
  std::vector<Problem> ps(n);
  for (int i = 0; i < n; ++i) {
    int rl, ql;
    if (fscanf(f, "%d %d %d %d", &rl, &ql, &ps[i].qs, &ps[i].qe) != 4) return 6;
    std::vector<char> buf((size_t)std::max(rl, ql) + 2);
    if (fscanf(f, "%s", buf.data()) != 1) return 6;
    ps[i].ref = buf.data();
    if (fscanf(f, "%s", buf.data()) != 1) return 6;
    ps[i].qry = buf.data();
    ps[i].lines.resize(ql);
    for (int y = 0; y < ql; ++y)
      if (fscanf(f, "%d %d", &ps[i].lines[y].offset, &ps[i].lines[y].length) != 2) return 6;
  }
  fclose(f);

  printf("batchsize %d %d\n", al->GetScoreBatchSize(), al->GetAlignBatchSize());
  // 1) one blocking SingleAlign per problem, as computeAlignment does
  for (int i = 0; i < n; ++i) {
    Align a;
    alloc_align(a, ps[i].qry.size());
    a.svType = 1234;  // "read id" hack of the caller (:363)
    int ret = -2;
    try {
      ret = al->SingleAlign(i, ps[i].lines.data(), (int)ps[i].lines.size(), ps[i].ref.c_str(),
                            ps[i].qry.c_str(), a, ps[i].qs, ps[i].qe, 0);
    } catch (...) {
      ret = -3;
    }
    print_align("single", i, ret, a);
    unsigned long expect = 0;
    bool ok = true;
    for (size_t y = 0; y < ps[i].lines.size(); ++y) {  // prepare() publishes offsetInMatrix
      ok = ok && ps[i].lines[y].offsetInMatrix == expect;
      expect += (unsigned long)ps[i].lines[y].length;
    }
    printf("offsetInMatrix %d %s\n", i, ok ? "ok" : "BAD");
    delete[] a.pBuffer1; delete[] a.pBuffer2; delete[] a.nmPerPosition;
  }
  // 2) the same problems through BatchAlign
  {
    std::vector<Align> as(n);
    std::vector<const char*> refs(n), qrys(n);
    std::vector<NgmlrB200BatchAlignArgs> args(n);
    for (int i = 0; i < n; ++i) {
      alloc_align(as[i], ps[i].qry.size());
      refs[i] = ps[i].ref.c_str();
      qrys[i] = ps[i].qry.c_str();
      args[i] = {ps[i].lines.data(), (int)ps[i].lines.size(), ps[i].qs, ps[i].qe};
    }
    int rc = al->BatchAlign(0, n, refs.data(), qrys.data(), as.data(), args.data());
    printf("batchalign rc=%d\n", rc);
    for (int i = 0; i < n; ++i) {
      // BatchAlign reports a failed problem the way SingleAlign does: Score == -1
      print_align("batch", i, as[i].Score == -1.0f ? -1 : 0, as[i]);
      delete[] as[i].pBuffer1; delete[] as[i].pBuffer2; delete[] as[i].nmPerPosition;
    }
  }
  // 3) BatchScore / SingleScore on (ref prefix, read prefix) pairs
  {
    std::vector<std::string> r(n), q(n);
    std::vector<const char*> rp(n), qp(n);
    std::vector<float> out(n, -1.0f);
    for (int i = 0; i < n; ++i) {
      r[i] = ps[i].ref.substr(0, 306);
      q[i] = ps[i].qry.substr(0, 256);
      rp[i] = r[i].c_str();
      qp[i] = q[i].c_str();
    }
    int rc = al->BatchScore(0, n, rp.data(), qp.data(), out.data(), 0);
    printf("batchscore rc=%d\n", rc);
    for (int i = 0; i < n; ++i) {
      float s = -7.0f;
      int r1 = al->SingleScore(10, 0, rp[i], qp[i], s, 0);
      printf("score %d %g %g %d\n", i, out[i], s, r1);
    }
  }
  // 4) the unimplemented overload throws, like the reference
  {
    Align a;
    bool threw = false;
    try { al->SingleAlign(0, 40, "ACGT", "ACGT", a, 0); } catch (...) { threw = true; }
    printf("single_int_corridor_throws %d\n", threw ? 1 : 0);
  }
  destroy(al);
  return 0;
}
