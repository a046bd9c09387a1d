// tests/cpp/test_multi_gpu.cpp -- the multi-GPU entry points of include/rednoise.h driven by a plain C++ program: no Python,
// no torch, no CUDA headers; it links librednoise_b200.so only (what a caller such as Core::init, Core.cpp:58-65, would do).
//
//   test_multi_gpu [ndev] [depth] [expected_fnv64_hex]
//
// 1. the whole field of Terrain(depth) on device 0                     (rn_terrain2d, host destination)
// 2. the same field in row slabs on ndev devices                       (rn_terrain2d_sharded, device slabs)
// 3. assembled on every device by NCCL all-gather and by peer copies   (rn_gather)
// 4. and assembled straight into host memory by every device at once   (rn_terrain2d_sharded, host destination)
// Every assembly must equal (1) bit for bit; (1) must match the reference digest when one is given.
// With one visible device the slabs run as two parts on that device (NCCL is skipped: it needs distinct devices).
// Exit codes: 0 ok, 77 no usable device (skip), 1 mismatch / error.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../include/rednoise.h"

#define CHECK(call)                                                                                   \
  do {                                                                                                \
    int st__ = (call);                                                                                \
    if (st__ != RN_OK) {                                                                              \
      fprintf(stderr, "%s -> %d (%s): %s\n", #call, st__, rn_status_string(st__), rn_last_error_string()); \
      return 1;                                                                                       \
    }                                                                                                 \
  } while (0)

static uint64_t fnv64(const void* p, size_t n)
{
  const unsigned char* b = (const unsigned char*)p;
  uint64_t h = 0xcbf29ce484222325ull;
  for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 0x100000001b3ull; }
  return h;
}

int main(int argc, char** argv)
{
  int want = argc > 1 ? atoi(argv[1]) : 2;
  const int depth = argc > 2 ? atoi(argv[2]) : 12;
  const char* expect = argc > 3 ? argv[3] : nullptr;
  const int visible = rn_device_count();
  if (visible < 1) { fprintf(stderr, "no CUDA device\n"); return 77; }
  rn_params p;
  CHECK(rn_params_default(&p));
  p.depth = depth;
  const int side = rn_terrain2d_side(depth);
  const size_t nelem = (size_t)side * side, bytes = nelem * sizeof(rn_element);

  // (1) one device
  rn_context* ctx0 = nullptr;
  {
    int st = rn_create(0, &ctx0);
    if (st == RN_ERR_NO_DEVICE) { fprintf(stderr, "device 0 unusable: %s\n", rn_last_error_string()); return 77; }
    CHECK(st);
  }
  std::vector<rn_element> whole(nelem);
  CHECK(rn_terrain2d(ctx0, &p, nullptr, whole.data(), 0, 0, nullptr, nullptr));
  CHECK(rn_destroy(ctx0));
  const uint64_t digest = fnv64(whole.data(), bytes);
  printf("single device: depth %d, %d x %d samples, fnv64 %016llx\n", depth, side, side, (unsigned long long)digest);
  if (expect && strtoull(expect, nullptr, 16) != digest) { fprintf(stderr, "digest differs from the reference's %s\n", expect); return 1; }

  // (2)-(4) several devices
  int ndev = want < 1 ? 1 : want;
  int devices[RN_MAX_DEVICES];
  const bool distinct = visible >= ndev;
  for (int k = 0; k < ndev; k++) devices[k] = distinct ? k : k % visible;
  rn_multi* m = nullptr;
  CHECK(rn_create_multi(devices, ndev, &m));
  const int rows = rn_slab_rows(side, ndev);
  const size_t slab_elems = (size_t)rows * side;
  std::vector<rn_element*> slab(ndev), full(ndev);
  for (int k = 0; k < ndev; k++) {
    CHECK(rn_device_alloc(rn_multi_context(m, k), slab_elems * sizeof(rn_element), (void**)&slab[k]));
    CHECK(rn_device_alloc(rn_multi_context(m, k), slab_elems * ndev * sizeof(rn_element), (void**)&full[k]));
  }
  CHECK(rn_terrain2d_sharded(m, &p, slab.data(), 0, nullptr, 0));
  std::vector<rn_element> back(slab_elems * ndev);
  int bad = 0;
  for (int mode = 1; mode >= 0; mode--) {  // 1: NCCL all-gather, 0: peer copies
    if (mode == 1 && !distinct) { printf("NCCL all-gather skipped: %d device(s) visible for %d parts\n", visible, ndev); continue; }
    int st = rn_gather(m, &p, slab.data(), full.data(), mode);
    if (mode == 1 && st == RN_ERR_UNSUPPORTED) { printf("NCCL all-gather skipped: %s\n", rn_last_error_string()); continue; }
    CHECK(st);
    for (int k = 0; k < ndev; k++) {
      memset(back.data(), 0xab, back.size() * sizeof(rn_element));
      CHECK(rn_copy_to_host(rn_multi_context(m, k), back.data(), full[k], back.size() * sizeof(rn_element)));
      const bool ok = memcmp(back.data(), whole.data(), bytes) == 0;
      printf("%s assembly on device %d (part %d of %d): %s\n", mode ? "NCCL all-gather" : "peer-copy", devices[k], k, ndev, ok ? "identical" : "DIFFERS");
      bad |= !ok;
    }
  }
  // (4) every device sends its slab straight into one pinned host buffer
  rn_element* host = nullptr;
  CHECK(rn_host_alloc(bytes, (void**)&host));
  memset(host, 0xcd, bytes);
  CHECK(rn_terrain2d_sharded(m, &p, nullptr, 0, host, 0));
  {
    const bool ok = memcmp(host, whole.data(), bytes) == 0;
    printf("host assembly from %d parts: %s\n", ndev, ok ? "identical" : "DIFFERS");
    bad |= !ok;
  }
  // a batch of independent tiles, round-robin over the devices == the same tiles one by one
  {
    rn_params q = p;
    q.depth = 8;
    const int s8 = rn_terrain2d_side(8), ntiles = 5;
    const size_t tile = (size_t)s8 * s8;
    uint32_t seeds[5] = {12345u, 7u, 0xffffffffu, 99u, 2024u};
    std::vector<rn_element*> tdst(ndev);
    for (int k = 0; k < ndev; k++) CHECK(rn_device_alloc(rn_multi_context(m, k), tile * ntiles * sizeof(rn_element), (void**)&tdst[k]));
    CHECK(rn_terrain2d_batch_sharded(m, &q, seeds, ntiles, tdst.data(), 0));
    std::vector<rn_element> one(tile), got(tile);
    for (int t = 0; t < ntiles; t++) {
      rn_params r = q;
      r.seed = seeds[t];
      CHECK(rn_terrain2d(rn_multi_context(m, 0), &r, nullptr, one.data(), 0, 0, nullptr, nullptr));
      CHECK(rn_copy_to_host(rn_multi_context(m, t % ndev), got.data(), tdst[t % ndev] + (size_t)(t / ndev) * tile, tile * sizeof(rn_element)));
      const bool ok = memcmp(one.data(), got.data(), tile * sizeof(rn_element)) == 0;
      if (!ok) printf("batch tile %d DIFFERS\n", t);
      bad |= !ok;
    }
    printf("sharded batch of %d tiles: %s\n", ntiles, bad ? "see above" : "identical");
    for (int k = 0; k < ndev; k++) CHECK(rn_device_free(rn_multi_context(m, k), tdst[k]));
  }
  CHECK(rn_host_free(host));
  for (int k = 0; k < ndev; k++) {
    CHECK(rn_device_free(rn_multi_context(m, k), slab[k]));
    CHECK(rn_device_free(rn_multi_context(m, k), full[k]));
  }
  CHECK(rn_destroy_multi(m));
  printf(bad ? "FAILED\n" : "OK\n");
  return bad ? 1 : 0;
}
