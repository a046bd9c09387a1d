// rn_cells.cu -- the next stage of the reference constructor after the field: the attributes of the
// Cell quadtree, i.e. the arithmetic of Terrain::generateRecursive (Terrain.cpp:260-356) and
// root->radius *= 4 (:97).
//
// The reference recurses depth-first and threads the parent's world-space centre down the call
// (:351-354); but a node's attributes depend only on 5 samples of the finished field (its centre
// and 4 corners) and on its parent's centre, which is itself a function of 5 samples.  So every
// node is independent: one thread per node, the parent's centre recomputed, nodes written level by
// level in Morton order (x = low bit of every digit) -- exactly the order in which the reference's
// recursion fills each level (numEntries[level]++, :343).  IEEE binary32 in the reference's
// operand order (-fmad=false, IEEE division and square root), bit-exact against the oracle.
#include "rn_kernels.h"
#include "../../include/rednoise.h"

namespace rn {

namespace {

struct CellLaw {
  float feature_scale;   // Terrain.cpp:169
  float size_over_fs;    // size / featureScale, Terrain.cpp:325
  float radius[16];      // Cell::radius of a node of each level (Terrain.cpp:342,348; root x4, :97) -- a per-level constant
};

__device__ __forceinline__ V3 ldpos(const float4* g, size_t pitch, int i, int j, float* shade)
{
  const float4 v = __ldg(g + (size_t)i * pitch + (size_t)j);
  *shade = v.w;
  return V3{v.x, v.y, v.z};
}

// member Vector3::normalise(), Vector3.h:109-120
__device__ __forceinline__ V3 normalise_member(V3 v)
{
  const float f = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
  if (f < 1e-10f) return V3{0.0f, 0.0f, 1.0f};
  const float d = 1.0f / f;
  return V3{v.x * d, v.y * d, v.z * d};
}

// Terrain.cpp:281-282 and the changes to centre.y at :290-300; returns the centre in feature-scaled units
__device__ __forceinline__ V3 centre_fs(V3 pm, V3 p00, V3 p11, V3 p01, V3 p10, float feature_scale, bool* below_sea)
{
  V3 c = vmul(vadd(vadd(vadd(vadd(pm, p00), p11), p01), p10), 0.2f);
  c = vmul(c, feature_scale);
  *below_sea = c.y < 0.0f;
  if (*below_sea) c.y *= 0.3f;                          // :292
  if (c.y < -0.2f) c.y = (c.y + 0.2f) * 0.6f - 0.2f;    // :299
  return c;
}

// One node.  `S(i, j, &shade)` returns the position (and shade) of lattice sample (i, j).
template <class Sampler>
__device__ __forceinline__ void eval_cell(rn_cell* __restrict__ out, uint64_t k, int level, uint32_t bx, uint32_t by, int depth,
                                          const CellLaw& law, const Sampler& S)
{
  const int W = 1 << depth;
  const int w = W >> level;
  const int minX = (int)bx * w, minY = (int)by * w, maxX = minX + w, maxY = minY + w, midX = minX + w / 2, midY = minY + w / 2;
  float sm, s00, s11, s01, s10;
  const V3 pm = S(midX, midY, &sm), p00 = S(minX, minY, &s00), p11 = S(maxX, maxY, &s11);
  const V3 p01 = S(minX, maxY, &s01), p10 = S(maxX, minY, &s10);
  // Terrain.cpp:281-289
  V3 centre = vmul(vmul(vadd(vadd(vadd(vadd(pm, p00), p11), p01), p10), 0.2f), law.feature_scale);
  V3 normal = normalise_member(vcross(vsub(p11, p00), vsub(p10, p01)));
  float shade = (2 * sm + s00 + s11 + s01 + s10) / 6.0f;
  if (centre.y < 0.0f) {  // :290-296 change at sea level
    centre.y *= 0.3f;
    shade = (shade - 1.0f) * 0.3f + 1.0f;
    normal.y /= 0.3f;
    normal = normalise_member(normal);
  }
  V3 colour;
  if (centre.y < -0.2f) {  // :297-303 water
    const float dpt = -0.2f - centre.y;
    centre.y = (centre.y + 0.2f) * 0.6f - 0.2f;
    const float gch = 0.9f * (1.0f - 0.3f * dpt);
    colour = V3{0.0f, gch > 0.0f ? gch : 0.0f, 0.7f};
    const float div = 1.0f / (1.0f + 0.5f * dpt);
    colour = vmul(colour, div);
  } else if (centre.y < 0.0f) {  // :304-308 sand
    const float height = centre.y / 0.4f + 0.3f;
    colour = V3{0.95f - 0.25f * height, 0.9f - 0.4f * height, 0.25f};
  } else {  // :309-324 trees / snow
    const float flatness = normal.y - 1.0f;
    const float height = centre.y - 0.0f;
    const V3 tree{0.0f, 0.45f + 0.04f * height, 0.03f * height};
    const float snowHeight = centre.y - 3.0f;
    const V3 snow{0.65f + 0.04f * snowHeight, 0.65f + 0.04f * snowHeight, 0.5f + 0.07f * snowHeight};
    float dif = (flatness + 0.5f) * 0.75f;
    if (snowHeight > -0.5f) dif -= (snowHeight + 0.5f) * 0.1f;
    float blend = dif * 2.0f;
    blend = blend < 0.0f ? 0.0f : (blend > 1.0f ? 1.0f : blend);
    colour = vadd(vmul(snow, 1.0f - blend), vmul(tree, blend));
  }
  centre = vmul(centre, law.size_over_fs);  // :325
  V3 position = centre;
  if (level > 0) {  // :327-328: minus the parent's world-space centre (recomputed from the parent's 5 samples)
    const int pw = w * 2;
    const int pminX = (int)(bx >> 1) * pw, pminY = (int)(by >> 1) * pw, pmaxX = pminX + pw, pmaxY = pminY + pw;
    float dummy;
    bool below;
    const V3 pc = centre_fs(S(pminX + w, pminY + w, &dummy), S(pminX, pminY, &dummy), S(pmaxX, pmaxY, &dummy), S(pminX, pmaxY, &dummy),
                            S(pmaxX, pminY, &dummy), law.feature_scale, &below);
    position = vsub(position, vmul(pc, law.size_over_fs));
  }
  colour = vmul(colour, shade);  // :337
  const float cr = colour.x < 0.0f ? 0.0f : (colour.x > 1.0f ? 1.0f : colour.x);  // :338
  const float cg = colour.y < 0.0f ? 0.0f : (colour.y > 1.0f ? 1.0f : colour.y);
  const float cb = colour.z < 0.0f ? 0.0f : (colour.z > 1.0f ? 1.0f : colour.z);
  const float rad = law.radius[level];
  const int n0 = (int)(normal.x * 127.0f), n1 = (int)(normal.y * 127.0f), n2 = (int)(normal.z * 127.0f);  // :330-332, truncation
  const uint32_t packed = ((uint32_t)n0 & 0xffu) | (((uint32_t)n1 & 0xffu) << 8) | (((uint32_t)n2 & 0xffu) << 16) | ((uint32_t)level << 24);
  // one 32-byte rn_cell = two 128-bit stores: {position, colour.r} {colour.g, colour.b, radius, normal[3] + level}
  float4* dst = reinterpret_cast<float4*>(out + k);
  dst[0] = make_float4(position.x, position.y, position.z, cr);
  dst[1] = make_float4(cg, cb, rad, __uint_as_float(packed));
}

__device__ __forceinline__ void morton_decode(uint32_t e, uint32_t* bx_out, uint32_t* by_out)
{
  uint32_t bx = e & 0x55555555u, by = (e >> 1) & 0x55555555u;  // x = even bits, y = odd bits
  bx = (bx | (bx >> 1)) & 0x33333333u; by = (by | (by >> 1)) & 0x33333333u;
  bx = (bx | (bx >> 2)) & 0x0f0f0f0fu; by = (by | (by >> 2)) & 0x0f0f0f0fu;
  bx = (bx | (bx >> 4)) & 0x00ff00ffu; by = (by | (by >> 4)) & 0x00ff00ffu;
  bx = (bx | (bx >> 8)) & 0x0000ffffu; by = (by | (by >> 8)) & 0x0000ffffu;
  *bx_out = bx;
  *by_out = by;
}

struct GlobalSampler {
  const float4* g;
  size_t pitch;
  __device__ __forceinline__ V3 operator()(int i, int j, float* shade) const { return ldpos(g, pitch, i, j, shade); }
};

// Levels 0 .. nlevels-1 (the coarse part of the tree): one thread per node, samples gathered from global memory.
__global__ void k_cells(rn_cell* __restrict__ out, const float4* __restrict__ g, size_t pitch, int depth, uint64_t ncells, CellLaw law)
{
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ncells) return;
  // level: offset(l) = (4^l - 1) / 3 <= k  <=>  4^l <= 3k + 1
  const int level = (63 - __clzll((long long)(3 * k + 1))) >> 1;
  const uint64_t off = ((1ull << (2 * level)) - 1) / 3;
  uint32_t bx, by;
  morton_decode((uint32_t)(k - off), &bx, &by);
  eval_cell(out, k, level, bx, by, depth, law, GlobalSampler{g, pitch});
}

// The leaf level (3/4 of all nodes; node = 2x2 lattice cells): a CTA takes a Morton block of 16x16 leaves, i.e.
// a contiguous 8 KB run of the output, stages its 33x33 samples in shared memory with coalesced row loads and
// evaluates one leaf per thread (the 8x8 parents' samples are in the same tile).
constexpr int kLeafBlock = 16;
constexpr int kLeafTile = 2 * kLeafBlock + 1;
struct TileSampler {
  const float4* t;
  int i0, j0;
  __device__ __forceinline__ V3 operator()(int i, int j, float* shade) const
  {
    const float4 v = t[(i - i0) * kLeafTile + (j - j0)];
    *shade = v.w;
    return V3{v.x, v.y, v.z};
  }
};
__global__ void __launch_bounds__(kLeafBlock * kLeafBlock) k_cells_leaf(rn_cell* __restrict__ out, const float4* __restrict__ g, size_t pitch,
                                                                       int depth, CellLaw law)
{
  __shared__ float4 tile[kLeafTile * kLeafTile];
  const int level = depth - 1;
  const uint64_t off = ((1ull << (2 * level)) - 1) / 3;
  // block index = Morton index of the 16x16 block at level depth-1-4
  uint32_t Bx, By;
  morton_decode(blockIdx.x, &Bx, &By);
  const int i0 = (int)Bx * 2 * kLeafBlock, j0 = (int)By * 2 * kLeafBlock;
  for (int idx = threadIdx.x; idx < kLeafTile * kLeafTile; idx += blockDim.x) {
    const int r = idx / kLeafTile, c = idx - r * kLeafTile;
    tile[idx] = __ldg(g + (size_t)(i0 + r) * pitch + (size_t)(j0 + c));
  }
  __syncthreads();
  const TileSampler S{tile, i0, j0};
  uint32_t lx, ly;
  morton_decode(threadIdx.x, &lx, &ly);  // position inside the block, in the block's own Morton order
  eval_cell(out, off + (uint64_t)blockIdx.x * (kLeafBlock * kLeafBlock) + threadIdx.x, level, Bx * kLeafBlock + lx, By * kLeafBlock + ly,
            depth, law, S);
  // the 8x8, 4x4 and 2x2 ancestors of these leaves (levels depth-2 .. depth-4) need nothing outside the tile either:
  // their own 5 samples and their parents' (the parent of a depth-4 node is the whole 32x32 block)
  int t = (int)threadIdx.x;
#pragma unroll
  for (int up = 1; up <= 3; up++) {
    const int nb = kLeafBlock >> up;  // nodes per side of this level inside the block
    if (t < nb * nb && level - up >= 1) {
      const int lv = level - up;
      const uint64_t offl = ((1ull << (2 * lv)) - 1) / 3;
      morton_decode((uint32_t)t, &lx, &ly);
      eval_cell(out, offl + (uint64_t)blockIdx.x * (uint64_t)(nb * nb) + (uint64_t)t, lv, Bx * nb + lx, By * nb + ly, depth, law, S);
    }
    t -= nb * nb;
    if (t < 0) break;
  }
}

}  // namespace

cudaError_t cells_generate(void* cells, const Elem* field, size_t pitch, int depth, float feature_scale, float size, float amplitude,
                           cudaStream_t stream, uint64_t* launches)
{
  uint64_t n = 0;
  for (int l = 0; l < depth; l++) n += 1ull << (2 * l);
  CellLaw law;
  law.feature_scale = feature_scale;
  law.size_over_fs = size / feature_scale;
  {
    float radius = size;
    for (int l = 0; l < 16; l++) {
      law.radius[l] = sqrtf((2.0f + amplitude) * radius * radius);  // Terrain.cpp:342
      if (l == 0) law.radius[l] *= 4.0f;                            // Terrain.cpp:97
      radius /= 2.0f;                                               // Terrain.cpp:348
    }
  }
  const float4* g = reinterpret_cast<const float4*>(field);
  rn_cell* out = static_cast<rn_cell*>(cells);
  if (depth - 1 >= 4) {
    // the last 4 levels through the tiled kernel, the coarser levels through the gather kernel
    const int first_tiled = depth - 4 >= 1 ? depth - 4 : 1;
    const uint64_t ncoarse = ((1ull << (2 * first_tiled)) - 1) / 3;
    k_cells<<<(unsigned)((ncoarse + 255) / 256), 256, 0, stream>>>(out, g, pitch, depth, ncoarse, law);
    const unsigned nblocks = 1u << (2 * (depth - 1 - 4));
    k_cells_leaf<<<nblocks, kLeafBlock * kLeafBlock, 0, stream>>>(out, g, pitch, depth, law);
    if (launches) *launches += 2;
  } else {
    k_cells<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(out, g, pitch, depth, n, law);
    if (launches) *launches += 1;
  }
  return cudaGetLastError();
}

}  // namespace rn
