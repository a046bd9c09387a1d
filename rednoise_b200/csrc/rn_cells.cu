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
  float size;
  float amplitude;
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

__global__ void k_cells(rn_cell* __restrict__ out, const float4* __restrict__ g, size_t pitch, int depth, uint64_t ncells, CellLaw law)
{
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ncells) return;
  // level: offset(l) = (4^l - 1) / 3 <= k  <=>  4^l <= 3k + 1
  const int level = (63 - __clzll((long long)(3 * k + 1))) >> 1;
  const uint64_t off = ((1ull << (2 * level)) - 1) / 3;
  const uint32_t e = (uint32_t)(k - off);
  // Morton decode, x = even bits, y = odd bits
  uint32_t bx = e & 0x55555555u, by = (e >> 1) & 0x55555555u;
  bx = (bx | (bx >> 1)) & 0x33333333u; by = (by | (by >> 1)) & 0x33333333u;
  bx = (bx | (bx >> 2)) & 0x0f0f0f0fu; by = (by | (by >> 2)) & 0x0f0f0f0fu;
  bx = (bx | (bx >> 4)) & 0x00ff00ffu; by = (by | (by >> 4)) & 0x00ff00ffu;
  bx = (bx | (bx >> 8)) & 0x0000ffffu; by = (by | (by >> 8)) & 0x0000ffffu;
  const int W = 1 << depth;
  const int w = W >> level;
  const int minX = (int)bx * w, minY = (int)by * w, maxX = minX + w, maxY = minY + w, midX = minX + w / 2, midY = minY + w / 2;
  float sm, s00, s11, s01, s10;
  const V3 pm = ldpos(g, pitch, midX, midY, &sm), p00 = ldpos(g, pitch, minX, minY, &s00), p11 = ldpos(g, pitch, maxX, maxY, &s11);
  const V3 p01 = ldpos(g, pitch, minX, maxY, &s01), p10 = ldpos(g, pitch, maxX, minY, &s10);
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
    const V3 pc = centre_fs(ldpos(g, pitch, pminX + w, pminY + w, &dummy), ldpos(g, pitch, pminX, pminY, &dummy),
                            ldpos(g, pitch, pmaxX, pmaxY, &dummy), ldpos(g, pitch, pminX, pmaxY, &dummy),
                            ldpos(g, pitch, pmaxX, pminY, &dummy), law.feature_scale, &below);
    position = vsub(position, vmul(pc, law.size_over_fs));
  }
  colour = vmul(colour, shade);  // :337
  const float cr = colour.x < 0.0f ? 0.0f : (colour.x > 1.0f ? 1.0f : colour.x);  // :338
  const float cg = colour.y < 0.0f ? 0.0f : (colour.y > 1.0f ? 1.0f : colour.y);
  const float cb = colour.z < 0.0f ? 0.0f : (colour.z > 1.0f ? 1.0f : colour.z);
  float radius = law.size;
  for (int l = 0; l < level; l++) radius /= 2.0f;  // :348, exact halvings
  float rad = sqrtf((2.0f + law.amplitude) * radius * radius);  // :342
  if (level == 0) rad *= 4.0f;                                   // :97
  const int n0 = (int)(normal.x * 127.0f), n1 = (int)(normal.y * 127.0f), n2 = (int)(normal.z * 127.0f);  // :330-332, truncation
  const uint32_t packed = ((uint32_t)n0 & 0xffu) | (((uint32_t)n1 & 0xffu) << 8) | (((uint32_t)n2 & 0xffu) << 16) | ((uint32_t)level << 24);
  // one 32-byte rn_cell = two 128-bit stores: {position, colour.r} {colour.g, colour.b, radius, normal[3] + level}
  float4* dst = reinterpret_cast<float4*>(out + k);
  dst[0] = make_float4(position.x, position.y, position.z, cr);
  dst[1] = make_float4(cg, cb, rad, __uint_as_float(packed));
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
  law.size = size;
  law.amplitude = amplitude;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  k_cells<<<blocks, 256, 0, stream>>>(static_cast<rn_cell*>(cells), reinterpret_cast<const float4*>(field), pitch, depth, n, law);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

}  // namespace rn
