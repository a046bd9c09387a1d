// rn_rock.cu -- the Rock generator: Rock::generateBaseRecursive on the six faces of a
// cube that share their edges (Rock.cpp:64-67,127-218) and the Cell attributes of Rock::generateRecursive (:220-271).
//
// The reference recurses depth-first and lets a `height == 0.0f` test decide whether an edge midpoint still has to be
// displaced, so the rand() call that displaces a sample is defined by the visiting order.  Here every call is
// addressed by a closed-form index instead:
//   * faces are visited 0..5; inside a face the quadtree in pre-order, children in Morton order (x = low bit); a node
//     draws for its left, right, bottom, top midpoint if that sample is still unset (Rock.cpp:181-212);
//   * nodes of one level are visited in Morton order, which is monotone in x and in y: the left and bottom midpoints
//     of a node have been set by its neighbours, unless the node lies on the face border, where the midpoint is set
//     iff a lower-numbered face shares that cube edge;
//   * hence a subtree of width 2^k draws E(k) - (2^k - 1) * (sides already set) times, E(k) = (2^k+1)^2 - 4 - (4^k-1)/3,
//     and the index of a node's slot is a sum of such terms along the path from the root.
// With the index known, the LCG state is reached by skip-ahead and the levels are evaluated breadth-first, one launch
// per level, each face on its own: a sample on a cube edge is computed by both faces that see it, from the same call
// index, so the faces never exchange data.  A displaced height of exactly 0.0f would make the reference draw again
// (value-dependent); the kernels flag it and rn_rock / rn_cloud then regenerate the input with the literal recursion on one
// device thread (k_rock_sequential below): slower, but the samples and the rand() call count are the reference's.
//
// KIND 1 is the Cloud generator (Cloud.cpp), the two-face variant: a lower and an upper sheet that share their whole border
// (Cloud::element, Cloud.cpp:116-121: face 1 finds all four sides set by face 0), ten preset control heights (:64-73) that the
// `height == 0.0f` test on centres (:131-132) preserves, flat boxPos/getPos (:106-113,167-173), grey colour, normal flipped on
// the upper sheet (:199-200).
#include "rn_kernels.h"
#include "rn_tile.cuh"
#include <atomic>

namespace rn {

namespace {

struct RockTopo {
  int pre[6][4];       // side (L, R, B, T) of face f belongs to a lower-numbered face
  uint32_t base[6];    // call index of the face's first draw (mod 2^32: only the LCG state is derived from it)
};

constexpr int kRock = 0, kCloud = 1;
constexpr int kRockSmallMaxDepth = 7;  // rocks up to this depth fit one CTA per face in shared memory (k_rock_small)
constexpr int kRockNodeLevel = 7;  // levels from this one on use the node-major kernel (measured: depth 7 77 vs 92 us, depth 11 0.89 vs 2.1 ms)

struct RockParams {
  float2* faces;       // nfaces x (W+1) x (W+1) {height, shade}: face f = the reference's currentFace, (x, y) as element() sees it
  int depth;
  uint32_t state;      // generator state when the rock is entered
  float amplitude;
  RockTopo topo;
  int* collision;      // set if a displaced height is exactly 0.0f
};

// Rock::element (Rock.cpp:127-151): face-local (x, y) -> cube coordinates
__host__ __device__ inline void cube_of(int W, int f, int x, int y, int& X, int& Y, int& Z)
{
  const int a = f >> 1, fixed = (f & 1) ? 0 : W;
  if (a == 0) { X = fixed; Y = x; Z = y; }
  else if (a == 1) { X = x; Y = fixed; Z = y; }
  else { X = x; Y = y; Z = fixed; }
}
__host__ __device__ inline bool face_local(int W, int g, int X, int Y, int Z, int& x, int& y)
{
  const int a = g >> 1, fixed = (g & 1) ? 0 : W;
  if (a == 0) { if (X != fixed) return false; x = Y; y = Z; }
  else if (a == 1) { if (Y != fixed) return false; x = X; y = Z; }
  else { if (Z != fixed) return false; x = X; y = Y; }
  return true;
}
__host__ __device__ inline uint32_t rock_E(int k)  // edge midpoints in the closed square of a node of width 2^k (mod 2^32)
{
  const uint32_t w = 1u << k;
  return (w + 1u) * (w + 1u) - 4u - (uint32_t)((((uint64_t)1 << (2 * k)) - 1u) / 3u);
}
__host__ __device__ inline uint32_t node_draws(const int* pre, int j, int qx, int qy, int upto)
{
  const int n = 1 << j;
  uint32_t d = 0;
  if (upto > 0 && qx == 0 && !pre[0]) d++;
  if (upto > 1 && !(qx == n - 1 && pre[1])) d++;
  if (upto > 2 && qy == 0 && !pre[2]) d++;
  if (upto > 3 && !(qy == n - 1 && pre[3])) d++;
  return d;
}
__host__ __device__ inline uint32_t subtree_draws(const int* pre, int depth, int j, int qx, int qy)
{
  const int n = 1 << j, k = depth - j;
  const int preset = ((qx > 0) || pre[0]) + ((qx == n - 1) && pre[1]) + ((qy > 0) || pre[2]) + ((qy == n - 1) && pre[3]);
  return rock_E(k) - ((1u << k) - 1u) * (uint32_t)preset;
}
// call index of slot `slot` of node (level l, qx, qy) of face f
__device__ inline uint32_t rock_call_index(const RockTopo& t, int depth, int f, int l, int qx, int qy, int slot)
{
  const int* pre = t.pre[f];
  uint32_t idx = t.base[f];
  for (int j = 0; j < l; j++) {
    const int ax = qx >> (l - j), ay = qy >> (l - j);
    idx += node_draws(pre, j, ax, ay, 4);
    const int cx = qx >> (l - j - 1), cy = qy >> (l - j - 1);
    const int ci = ((cy & 1) << 1) | (cx & 1);
    for (int s = 0; s < ci; s++) idx += subtree_draws(pre, depth, j + 1, (ax << 1) | (s & 1), (ay << 1) | (s >> 1));
  }
  return idx + node_draws(pre, l, qx, qy, slot);
}
// call index of the edge midpoint (x, y) of face f at level l, resolved on the face that draws it
template <int KIND>
__device__ inline uint32_t rock_point_call_index(const RockTopo& t, int depth, int f, int l, int x, int y)
{
  const int W = 1 << depth, sh = depth - l;  // node width w = 1 << sh
  int X = 0, Y = 0, Z = 0;
  if (KIND == kRock) cube_of(W, f, x, y, X, Y, Z);
  for (int g = 0; g < (KIND == kRock ? 6 : 2); g++) {
    int gx, gy;
    if (KIND == kRock) {
      if (!face_local(W, g, X, Y, Z, gx, gy)) continue;
    } else {  // Cloud::element: the border of the upper sheet IS the border of the lower one
      const bool border = x == 0 || y == 0 || x == W || y == W;
      if (g != (border ? 0 : f)) continue;
      gx = x; gy = y;
    }
    if ((gx & ((1 << sh) - 1)) == 0) {  // on a vertical node side: right slot of the node on its left, left slot in column 0
      const int qy = gy >> sh;
      if (gx == 0) return rock_call_index(t, depth, g, l, 0, qy, 0);
      return rock_call_index(t, depth, g, l, (gx >> sh) - 1, qy, 1);
    }
    const int qx = gx >> sh;
    if (gy == 0) return rock_call_index(t, depth, g, l, qx, 0, 2);
    return rock_call_index(t, depth, g, l, qx, (gy >> sh) - 1, 3);
  }
  return 0;
}

// One level: the centres (Rock.cpp:179) and edge midpoints (:181-212) of all nodes of width W >> l, on the
// half-step lattice of that level ((2n+1)^2 points per face; those with two even coordinates are older).
// One level, one thread per SAMPLE (used for the coarse levels, where a level is a handful of warps and latency counts): the centres (Rock.cpp:179) and edge midpoints (:181-212) of all nodes of width W >> l, on the
// half-step lattice of that level ((2n+1)^2 points per face; those with two even coordinates are older).
// the sample (px, py) of the half-step lattice of level l on face f, whose (W+1)^2 samples start at g (global or shared memory)
template <int KIND>
__device__ __forceinline__ void rock_point(const RockParams& rp, float2* g, int f, int l, int px, int py)
{
  const int depth = rp.depth, W = 1 << depth, P = W + 1;
  const int ox = px & 1, oy = py & 1;
  if (!ox && !oy) return;
  const int hsh = depth - l - 1, h = 1 << hsh;  // half step
  const int x = px << hsh, y = py << hsh;
  float2 e;
  if (ox && oy) {
    if (KIND == kCloud && g[(size_t)x * P + y].x != 0.0f) return;  // a preset control height (Cloud.cpp:64-73,131)
    const float2 a = g[(size_t)(x - h) * P + (y - h)], b = g[(size_t)(x + h) * P + (y - h)], c = g[(size_t)(x - h) * P + (y + h)],
                 d = g[(size_t)(x + h) * P + (y + h)];
    e.x = (a.x + b.x + c.x + d.x) * 0.25f;  // Rock.h:25-29
    e.y = (a.y + b.y + c.y + d.y) * 0.25f;
  } else {
    const float2 a = ox ? g[(size_t)(x - h) * P + y] : g[(size_t)x * P + (y - h)], b = ox ? g[(size_t)(x + h) * P + y] : g[(size_t)x * P + (y + h)];
    const uint32_t k = rock_point_call_index<KIND>(rp.topo, depth, f, l, x, y);
    const float r = random_from_draw(lcg_draw_from_state(lcg_skip(rp.state, k + 1u)));
    const float offset = rp.amplitude * r * (float)(2 * h) / (float)W;  // Rock.cpp:185: amplitude * random() * extent / maxWidth
    e.x = (a.x + b.x) * 0.5f + offset;                                   // Rock.h:30-34, Rock.cpp:186
    e.y = (a.y + b.y) * 0.5f + offset * 1.5f;                            // radiosityStrength, Rock.cpp:171,187
    if (e.x == 0.0f) atomicExch(rp.collision, 1);
  }
  g[(size_t)x * P + y] = e;
}

template <int KIND>
__global__ void k_rock_level_points(const __grid_constant__ RockParams rp, int l)
{
  const int P = (1 << rp.depth) + 1;
  const int n2 = (2 << l) + 1;  // points per side of the half-step lattice
  const int f = blockIdx.z;
  const int px = blockIdx.y * blockDim.y + threadIdx.y, py = blockIdx.x * blockDim.x + threadIdx.x;
  if (px >= n2 || py >= n2) return;
  rock_point<KIND>(rp, rp.faces + (size_t)f * P * P, f, l, px, py);
}

// One level, one thread per quadtree NODE (f, qx, qy) of width W >> l: its centre (Rock.cpp:179) and the edge midpoints it is
// responsible for -- right and top always, left / bottom only in column / row 0 -- so every sample of the face is written once
// and the walk from the root that yields the call index is done once per node instead of once per sample.  A midpoint on a
// face side that an earlier face has drawn takes its index from that face (rock_point_call_index).
template <int KIND>
__device__ __forceinline__ void rock_edge(const RockParams& rp, float2* g, int P, int W, int f, int l, int x, int y, bool along_x, int h, bool mine,
                                          uint32_t k_mine)
{
  const float2 a = along_x ? g[(size_t)(x - h) * P + y] : g[(size_t)x * P + (y - h)], b = along_x ? g[(size_t)(x + h) * P + y] : g[(size_t)x * P + (y + h)];
  const uint32_t k = mine ? k_mine : rock_point_call_index<KIND>(rp.topo, rp.depth, f, l, x, y);
  const float r = random_from_draw(lcg_draw_from_state(lcg_skip(rp.state, k + 1u)));
  const float offset = rp.amplitude * r * (float)(2 * h) / (float)W;  // Rock.cpp:185: amplitude * random() * extent / maxWidth
  float2 e;
  e.x = (a.x + b.x) * 0.5f + offset;                                   // Rock.h:30-34, Rock.cpp:186
  e.y = (a.y + b.y) * 0.5f + offset * 1.5f;                            // radiosityStrength, Rock.cpp:171,187
  if (e.x == 0.0f) atomicExch(rp.collision, 1);
  g[(size_t)x * P + y] = e;
}

template <int KIND>
__global__ void k_rock_level(const __grid_constant__ RockParams rp, int l)
{
  const int depth = rp.depth, W = 1 << depth, P = W + 1;
  const int n = 1 << l;
  const int f = blockIdx.z;
  const int qx = blockIdx.y * blockDim.y + threadIdx.y, qy = blockIdx.x * blockDim.x + threadIdx.x;
  if (qx >= n || qy >= n) return;
  const int w = W >> l, h = w >> 1;
  const int x0 = qx * w, y0 = qy * w, x1 = x0 + w, y1 = y0 + w, mx = x0 + h, my = y0 + h;
  float2* g = rp.faces + (size_t)f * P * P;
  if (!(KIND == kCloud && g[(size_t)mx * P + my].x != 0.0f)) {  // (a preset control height of a cloud stays: Cloud.cpp:64-73,131)
    const float2 a = g[(size_t)x0 * P + y0], b = g[(size_t)x1 * P + y0], c = g[(size_t)x0 * P + y1], d = g[(size_t)x1 * P + y1];
    g[(size_t)mx * P + my] = make_float2((a.x + b.x + c.x + d.x) * 0.25f, (a.y + b.y + c.y + d.y) * 0.25f);  // Rock.h:25-29
  }
  const int* pre = rp.topo.pre[f];
  const uint32_t k0 = rock_call_index(rp.topo, depth, f, l, qx, qy, 0);  // call index of the node's first draw
  // which of the four slots (left, right, bottom, top) this node draws itself (Rock.cpp:181-212 and the visiting order)
  const bool d0 = qx == 0 && !pre[0], d1 = !(qx == n - 1 && pre[1]), d2 = qy == 0 && !pre[2], d3 = !(qy == n - 1 && pre[3]);
  if (qx == 0) rock_edge<KIND>(rp, g, P, W, f, l, x0, my, false, h, d0, k0);
  rock_edge<KIND>(rp, g, P, W, f, l, x1, my, false, h, d1, k0 + (uint32_t)d0);
  if (qy == 0) rock_edge<KIND>(rp, g, P, W, f, l, mx, y0, true, h, d2, k0 + (uint32_t)d0 + (uint32_t)d1);
  rock_edge<KIND>(rp, g, P, W, f, l, mx, y1, true, h, d3, k0 + (uint32_t)d0 + (uint32_t)d1 + (uint32_t)d2);
}

__global__ void k_rock_init(float2* faces, size_t n, int cloud_W)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float2 v = make_float2(0.0f, 1.0f);  // Rock.cpp:57-60, Cloud.cpp:59-60
  if (cloud_W) {                       // Cloud.cpp:64-73: centre and quarter points of both sheets
    const size_t P = (size_t)cloud_W + 1, per = P * P;
    const int f = (int)(i / per), x = (int)((i % per) / P), y = (int)(i % P), q = cloud_W / 4;
    const bool preset = (x == 2 * q && y == 2 * q) || ((x == q || x == 3 * q) && (y == q || y == 3 * q));
    if (preset) v.x = f ? 0.5f : -0.5f;
  }
  faces[i] = v;
}

// ---- The literal algorithm, for the rare inputs whose draws are value-dependent ----
// When a displaced height is exactly 0.0f the reference's next visitor of that sample draws again (Rock.cpp:181-212), the
// subtree visited in between has used the first value, and every later call index moves: there is no closed form.  Those
// inputs (about 1 rock in 500) are generated by ONE thread that runs the reference's recursion itself -- same visiting order,
// same tests -- on the storage Rock::element / Cloud::element resolve to (here: the lowest-numbered face that contains the
// sample), after which k_rock_resolve copies the shared samples into every face's view.  Slow (tens of ms at depth 7), exact.
template <int KIND>
__device__ inline float2* rock_canon(float2* faces, int W, int f, int x, int y)
{
  const size_t P = (size_t)W + 1;
  if (KIND == kCloud) {
    const bool border = x == 0 || y == 0 || x == W || y == W;
    return faces + ((size_t)(border ? 0 : f) * P + (size_t)x) * P + (size_t)y;
  }
  int X, Y, Z;
  cube_of(W, f, x, y, X, Y, Z);
  for (int g = 0; g < 6; g++) {
    int gx, gy;
    if (face_local(W, g, X, Y, Z, gx, gy)) return faces + ((size_t)g * P + (size_t)gx) * P + (size_t)gy;
  }
  return faces;
}

template <int KIND>
__device__ inline void rock_seq_edge(const RockParams& rp, int W, int f, int x, int y, int ax, int ay, int bx, int by, int extent, uint32_t& state,
                                     unsigned long long& draws)
{
  float2* el = rock_canon<KIND>(rp.faces, W, f, x, y);
  if (el->x == 0.0f) {
    const float2 a = *rock_canon<KIND>(rp.faces, W, f, ax, ay), b = *rock_canon<KIND>(rp.faces, W, f, bx, by);
    state = state * kLcgA + kLcgC;
    draws++;
    const float offset = rp.amplitude * random_from_draw(lcg_draw_from_state(state)) * (float)extent / (float)W;
    *el = make_float2((a.x + b.x) * 0.5f + offset, (a.y + b.y) * 0.5f + offset * 1.5f);
  }
}

template <int KIND>
__global__ void k_rock_sequential(const __grid_constant__ RockParams rp, unsigned long long* draws_out)
{
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const int W = 1 << rp.depth;
  uint32_t state = rp.state;
  unsigned long long draws = 0;
  for (int f = 0; f < (KIND == kCloud ? 2 : 6); f++) {
    int sx[48], sy[48], sw[48];  // pre-order stack: at most 3 pending siblings per level
    int sp = 0;
    sx[0] = 0; sy[0] = 0; sw[0] = W; sp = 1;
    while (sp > 0) {
      sp--;
      const int x0 = sx[sp], y0 = sy[sp], w = sw[sp];
      if (w <= 1) continue;
      const int h = w >> 1, x1 = x0 + w, y1 = y0 + w, mx = x0 + h, my = y0 + h;
      float2* m = rock_canon<KIND>(rp.faces, W, f, mx, my);
      if (KIND == kRock || m->x == 0.0f) {  // Rock.cpp:179 / Cloud.cpp:131-132
        const float2 a = *rock_canon<KIND>(rp.faces, W, f, x0, y0), b = *rock_canon<KIND>(rp.faces, W, f, x1, y0),
                     c = *rock_canon<KIND>(rp.faces, W, f, x0, y1), d = *rock_canon<KIND>(rp.faces, W, f, x1, y1);
        *m = make_float2((a.x + b.x + c.x + d.x) * 0.25f, (a.y + b.y + c.y + d.y) * 0.25f);
      }
      rock_seq_edge<KIND>(rp, W, f, x0, my, x0, y0, x0, y1, w, state, draws);
      rock_seq_edge<KIND>(rp, W, f, x1, my, x1, y0, x1, y1, w, state, draws);
      rock_seq_edge<KIND>(rp, W, f, mx, y0, x1, y0, x0, y0, w, state, draws);
      rock_seq_edge<KIND>(rp, W, f, mx, y1, x1, y1, x0, y1, w, state, draws);
      // children in the order of Rock.cpp:213-216 (pushed in reverse)
      sx[sp] = mx; sy[sp] = my; sw[sp] = h; sp++;
      sx[sp] = x0; sy[sp] = my; sw[sp] = h; sp++;
      sx[sp] = mx; sy[sp] = y0; sw[sp] = h; sp++;
      sx[sp] = x0; sy[sp] = y0; sw[sp] = h; sp++;
    }
  }
  *draws_out = draws;
}

template <int KIND>
__global__ void k_rock_resolve(float2* faces, int W, size_t n)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t P = (size_t)W + 1, per = P * P;
  const int f = (int)(i / per), x = (int)((i % per) / P), y = (int)(i % P);
  const float2* src = rock_canon<KIND>(faces, W, f, x, y);
  if (src != faces + i) faces[i] = *src;
}

// ---- Cell attributes: Rock::generateRecursive (Rock.cpp:228-271) ----
__device__ inline V3 rock_box_pos(int W, int f, int x, int y)  // Rock.cpp:100-124
{
  const float fw = (float)W;
  const int a = f >> 1;
  const float fixed = (f & 1) ? 0.f : 1.f;
  const float u = __fdiv_rn((float)x, fw), v = __fdiv_rn((float)y, fw);
  float X, Y, Z;
  if (a == 0) { X = fixed; Y = u; Z = v; }
  else if (a == 1) { X = u; Y = fixed; Z = v; }
  else { X = u; Y = v; Z = fixed; }
  return V3{X - 0.5f, Y - 0.5f, Z - 0.5f};
}
__device__ inline V3 rock_normalise(V3 v)  // member normalise(), Vector3.h:109-120
{
  const float f = __fsqrt_rn(v.x * v.x + v.y * v.y + v.z * v.z);
  if (f < 1e-10f) return V3{0.0f, 0.0f, 1.0f};
  const float d = __frcp_rn(f);
  return V3{v.x * d, v.y * d, v.z * d};
}
template <int KIND>
__device__ inline V3 rock_get_pos(const float2* g, int P, int W, int f, int x, int y)
{
  if (KIND == kCloud) {  // Cloud.cpp:106-113,167-173
    const float fw = (float)W;
    return V3{-1.f + __fdiv_rn(2.0f * (float)x, fw), 0.0f + g[(size_t)x * P + y].x, -1.f + __fdiv_rn(2.0f * (float)y, fw)};
  }
  const V3 r = rock_normalise(rock_box_pos(W, f, x, y));  // Rock.cpp:220-227
  const float s = 1.0f + g[(size_t)x * P + y].x;
  return V3{r.x * s, r.y * s, r.z * s};
}
__device__ inline float clampedf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }  // basics.h:12-15

struct RockCell {  // rn_cell (include/rednoise.h)
  float position[3];
  float colour[3];
  float radius;
  int8_t normal[3];
  uint8_t level;
};

// node e_all (level-major) of face f, whose samples start at g (global or shared memory), into out[e_all]
template <int KIND>
__device__ __forceinline__ void rock_cell(RockCell* __restrict__ out, const float2* g, int depth, int f, uint32_t e_all, float radius0)
{
  const int W = 1 << depth, P = W + 1;
  // level-major index -> (level, entry); entry = Morton code of the node, x in the low bit (children order of :266-269)
  int l = 0;
  uint32_t off = 0;
  while (off + (1u << (2 * l)) <= e_all) { off += 1u << (2 * l); l++; }
  const uint32_t m = e_all - off;
  int qx = 0, qy = 0;
  for (int b = 0; b < l; b++) { qx |= ((m >> (2 * b)) & 1u) << b; qy |= ((m >> (2 * b + 1)) & 1u) << b; }
  const int w = W >> l;
  const int minX = qx * w, minY = qy * w, maxX = minX + w, maxY = minY + w, midX = minX + w / 2, midY = minY + w / 2;
  V3 centre = rock_get_pos<KIND>(g, P, W, f, midX, midY);
  centre = vmul(centre, 1.0f);  // `* size` with the constructor's size argument 1.0f (Rock.cpp:77,236)
  V3 position = centre;
  {  // cell->position -= parentPosWorldSpace: the parent's centre, (0,0,0) for the root PhysicalCell
    V3 parent{0.0f, 0.0f, 0.0f};
    if (l > 0) {
      const int pw = w * 2, pX = (qx >> 1) * pw + pw / 2, pY = (qy >> 1) * pw + pw / 2;
      parent = vmul(rock_get_pos<KIND>(g, P, W, f, pX, pY), 1.0f);
    }
    position = vsub(position, parent);
  }
  const float shade = g[(size_t)midX * P + midY].y;
  RockCell c;
  c.colour[0] = clampedf((KIND == kCloud ? 0.5f : 0.6f) * shade, 0.0f, 1.0f);   // Rock.cpp:240-242, Cloud.cpp:186-188
  c.colour[1] = clampedf((KIND == kCloud ? 0.5f : 0.54f) * shade, 0.0f, 1.0f);
  c.colour[2] = clampedf((KIND == kCloud ? 0.5f : 0.48f) * shade, 0.0f, 1.0f);
  c.position[0] = position.x; c.position[1] = position.y; c.position[2] = position.z;
  const V3 xc = vsub(rock_get_pos<KIND>(g, P, W, f, minX, minY), rock_get_pos<KIND>(g, P, W, f, maxX, maxY));
  const V3 yc = vsub(rock_get_pos<KIND>(g, P, W, f, maxX, minY), rock_get_pos<KIND>(g, P, W, f, minX, maxY));
  V3 normal = vcross(xc, yc);
  const bool flip = KIND == kCloud ? f == 1 : normal.x * centre.x + normal.y * centre.y + normal.z * centre.z < 0.0f;  // Cloud.cpp:199, Rock.cpp:250
  if (flip) normal = V3{-normal.x, -normal.y, -normal.z};
  normal = rock_normalise(normal);
  c.normal[0] = (int8_t)(int)(normal.x * 127.0f);
  c.normal[1] = (int8_t)(int)(normal.y * 127.0f);
  c.normal[2] = (int8_t)(int)(normal.z * 127.0f);
  float radius = radius0;
  for (int b = 0; b < l; b++) radius = __fdiv_rn(radius, 2.0f);  // radius /= 2.0f per level (:264)
  c.radius = radius;
  c.level = (uint8_t)l;
  const float4* src = reinterpret_cast<const float4*>(&c);
  float4* dst = reinterpret_cast<float4*>(out + e_all);
  dst[0] = src[0];
  dst[1] = src[1];
}

template <int KIND>
__global__ void k_rock_cells(RockCell* __restrict__ out, const float2* __restrict__ faces, int depth, uint32_t count, float radius0)
{
  const uint32_t e_all = blockIdx.x * blockDim.x + threadIdx.x;
  if (e_all >= count) return;
  const int f = blockIdx.y;
  const size_t P = ((size_t)1 << depth) + 1;
  rock_cell<KIND>(out + (size_t)f * count, faces + (size_t)f * P * P, depth, f, e_all, radius0);
}

// ---- small rocks, ONE launch for a whole batch (the app builds 20 rocks of depth 2..7 back to back, Core.cpp:69-79) ----
// A CTA owns one face of one rock: the face lives in shared memory ((W+1)^2 x 8 B <= 133 KB at depth 7) through all levels
// (one barrier per level: a level's centres and edge midpoints read older samples only), is then written out, and the Cell
// attributes of its quadtree are evaluated from the same shared image.  Same per-sample code as the level kernels.
struct RockBatchItem {
  RockParams rp;     // rp.faces: device destination of the rock's faces, or nullptr
  RockCell* cells;   // device destination of the rock's nfaces x count cells, or nullptr
  float radius0;
};

template <int KIND>
__global__ void __launch_bounds__(256) k_rock_small(const RockBatchItem* __restrict__ items)
{
  extern __shared__ __align__(16) float2 gs[];
  const RockBatchItem& it = items[blockIdx.y];
  const RockParams rp = it.rp;
  const int f = blockIdx.x, depth = rp.depth, W = 1 << depth, P = W + 1, tid = threadIdx.x;
  const int q = W / 4;
  for (int i = tid; i < P * P; i += 256) {
    float2 v = make_float2(0.0f, 1.0f);  // Rock.cpp:57-60, Cloud.cpp:59-60
    if (KIND == kCloud) {                // Cloud.cpp:64-73: centre and quarter points of both sheets
      const int x = i / P, y = i - x * P;
      const bool preset = (x == 2 * q && y == 2 * q) || ((x == q || x == 3 * q) && (y == q || y == 3 * q));
      if (preset) v.x = f ? 0.5f : -0.5f;
    }
    gs[i] = v;
  }
  __syncthreads();
  for (int l = 0; l < depth; l++) {
    const int n2 = (2 << l) + 1;
    for (int i = tid; i < n2 * n2; i += 256) {
      const int px = i / n2, py = i - px * n2;
      rock_point<KIND>(rp, gs, f, l, px, py);
    }
    __syncthreads();
  }
  if (rp.faces) {
    float2* dst = rp.faces + (size_t)f * P * P;
    for (int i = tid; i < P * P; i += 256) dst[i] = gs[i];
  }
  if (it.cells) {
    uint32_t count = 0;
    for (int l = 0; l < depth; l++) count += 1u << (2 * l);
    for (uint32_t e = tid; e < count; e += 256) rock_cell<KIND>(it.cells + (size_t)f * count, gs, depth, f, e, it.radius0);
  }
}


}  // namespace

// which sides of a face were drawn by an earlier face (cube topology of Rock::element / shared border of Cloud::element), and
// the faces' first call
static void rock_topo(int kind, int depth, RockTopo& topo)
{
  const int W = 1 << depth;
  const int sx[4] = {0, W, W / 2, W / 2}, sy[4] = {W / 2, W / 2, 0, W};
  uint32_t acc = 0;
  for (int f = 0; f < 6; f++) {
    for (int s = 0; s < 4; s++) {
      topo.pre[f][s] = 0;
      if (kind == kCloud) { topo.pre[f][s] = f == 1; continue; }
      int X, Y, Z, x, y;
      cube_of(W, f, sx[s], sy[s], X, Y, Z);
      for (int g = 0; g < f; g++)
        if (face_local(W, g, X, Y, Z, x, y)) topo.pre[f][s] = 1;
    }
    topo.base[f] = acc;
    acc += rock_E(depth) - (uint32_t)(W - 1) * (uint32_t)(topo.pre[f][0] + topo.pre[f][1] + topo.pre[f][2] + topo.pre[f][3]);
  }
}

uint64_t rock_draw_count(int depth)
{
  const uint64_t W = (uint64_t)1 << depth;
  return 4 * W * W - 4;
}

// faces: device, nfaces x (W+1)^2 float2 (6 for a rock, 2 for a cloud); cells: device or nullptr, nfaces x (4^depth - 1)/3 rn_cell;
// flag: device int (collision)
// sequential != 0: the literal one-thread path (k_rock_sequential); flag then receives the rand() call count as an unsigned long long
cudaError_t rock_generate(void* faces_v, void* cells_v, int* flag, int kind, int depth, uint32_t state, float amplitude, int sequential,
                          cudaStream_t stream, uint64_t* launches)
{
  float2* faces = static_cast<float2*>(faces_v);
  const int W = 1 << depth, nfaces = kind == kCloud ? 2 : 6;
  const size_t P = (size_t)W + 1, nall = (size_t)nfaces * P * P;
  RockParams rp;
  rp.faces = faces; rp.depth = depth; rp.state = state; rp.amplitude = amplitude; rp.collision = flag;
  rock_topo(kind, depth, rp.topo);
  cudaError_t e = cudaMemsetAsync(flag, 0, 2 * sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return e;
  k_rock_init<<<(unsigned)((nall + 255) / 256), 256, 0, stream>>>(faces, nall, kind == kCloud ? W : 0);
  uint64_t nl = 1;
  if (sequential) {
    unsigned long long* draws = reinterpret_cast<unsigned long long*>(flag);
    if (kind == kCloud) {
      k_rock_sequential<kCloud><<<1, 32, 0, stream>>>(rp, draws);
      k_rock_resolve<kCloud><<<(unsigned)((nall + 255) / 256), 256, 0, stream>>>(faces, W, nall);
    } else {
      k_rock_sequential<kRock><<<1, 32, 0, stream>>>(rp, draws);
      k_rock_resolve<kRock><<<(unsigned)((nall + 255) / 256), 256, 0, stream>>>(faces, W, nall);
    }
    nl += 2;
  }
  for (int l = 0; l < (sequential ? 0 : depth); l++) {
    if (l < kRockNodeLevel) {  // few samples: one thread per sample (shortest critical path)
      const int n2 = (2 << l) + 1;
      const dim3 block(32, 8), grid((n2 + 31) / 32, (n2 + 7) / 8, nfaces);
      if (kind == kCloud) k_rock_level_points<kCloud><<<grid, block, 0, stream>>>(rp, l);
      else k_rock_level_points<kRock><<<grid, block, 0, stream>>>(rp, l);
    } else {                   // many samples: one thread per node (one index walk for its 3-5 samples)
      const int nn = 1 << l;
      const dim3 block(32, 8), grid((nn + 31) / 32, (nn + 7) / 8, nfaces);
      if (kind == kCloud) k_rock_level<kCloud><<<grid, block, 0, stream>>>(rp, l);
      else k_rock_level<kRock><<<grid, block, 0, stream>>>(rp, l);
    }
    nl++;
  }
  if (cells_v) {
    uint32_t count = 0;
    for (int l = 0; l < depth; l++) count += 1u << (2 * l);
    const float radius0 = (kind == kCloud ? 1.8f : 1.4f) * sqrtf((1.0f + 0.5f * amplitude) * 0.5f);  // Rock.cpp:76, Cloud.cpp:84
    const dim3 grid((count + 127) / 128, nfaces);
    if (kind == kCloud) k_rock_cells<kCloud><<<grid, 128, 0, stream>>>(static_cast<RockCell*>(cells_v), faces, depth, count, radius0);
    else k_rock_cells<kRock><<<grid, 128, 0, stream>>>(static_cast<RockCell*>(cells_v), faces, depth, count, radius0);
    nl++;
  }
  if (launches) *launches += nl;
  return cudaGetLastError();
}

// `n` rocks of depth <= kRockSmallMaxDepth in ONE launch (k_rock_small): states[i] = generator state when rock i is entered,
// faces[i] / cells[i] = device destinations (or nullptr), flags = n device ints (collision per rock), items_dev = device scratch of
// rock_small_item_bytes() * n bytes, items_host = pinned or pageable host scratch of the same size.
size_t rock_small_item_bytes() { return sizeof(RockBatchItem); }
cudaError_t rock_small_batch(int n, const int* depths, const uint32_t* states, float amplitude, void* const* faces, void* const* cells, int* flags,
                             void* items_dev, void* items_host, cudaStream_t stream, uint64_t* launches)
{
  RockBatchItem* h = static_cast<RockBatchItem*>(items_host);
  int maxd = 1;
  for (int i = 0; i < n; i++) {
    if (depths[i] < 1 || depths[i] > kRockSmallMaxDepth) return cudaErrorInvalidValue;
    RockBatchItem& it = h[i];
    it.rp.faces = static_cast<float2*>(faces ? faces[i] : nullptr);
    it.rp.depth = depths[i];
    it.rp.state = states[i];
    it.rp.amplitude = amplitude;
    it.rp.collision = flags + i;
    rock_topo(kRock, depths[i], it.rp.topo);
    it.cells = static_cast<RockCell*>(cells ? cells[i] : nullptr);
    it.radius0 = 1.4f * sqrtf((1.0f + 0.5f * amplitude) * 0.5f);  // Rock.cpp:76
    if (depths[i] > maxd) maxd = depths[i];
  }
  cudaError_t e = cudaMemsetAsync(flags, 0, sizeof(int) * (size_t)n, stream);
  if (e != cudaSuccess) return e;
  e = cudaMemcpyAsync(items_dev, h, sizeof(RockBatchItem) * (size_t)n, cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) return e;
  const size_t P = ((size_t)1 << maxd) + 1, bytes = P * P * sizeof(float2);
  static std::atomic<size_t> configured[64] = {};
  int dev = 0;
  e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 64 || configured[dev].load(std::memory_order_acquire) < bytes) {
    e = cudaFuncSetAttribute(k_rock_small<kRock>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(((size_t)1 << kRockSmallMaxDepth) + 1) * (int)(((size_t)1 << kRockSmallMaxDepth) + 1) * (int)sizeof(float2));
    if (e != cudaSuccess) return e;
    if (dev < 64) configured[dev].store((size_t)1 << 30, std::memory_order_release);
  }
  k_rock_small<kRock><<<dim3(6, (unsigned)n), 256, bytes, stream>>>(static_cast<const RockBatchItem*>(items_dev));
  if (launches) *launches += 1;
  return cudaGetLastError();
}

uint64_t cloud_draw_count(int depth)  // the lower sheet draws every edge midpoint of its closed square, the upper one all but the border
{
  const uint64_t W = (uint64_t)1 << depth;
  const uint64_t E = (W + 1) * (W + 1) - 4 - ((((uint64_t)1 << (2 * depth)) - 1) / 3);
  return 2 * E - 4 * (W - 1);
}

}  // namespace rn
