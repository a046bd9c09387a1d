/* oracle/rn_oracle_cloud.c -- TEST INFRASTRUCTURE ONLY (see rn_oracle.h).
 *
 * Plain-C restatement of the Cloud generator (renderEngine/source/Cloud.cpp, include/Cloud.h): the two-face
 * variant of Rock -- a lower and an upper sheet that share their border (element, Cloud.cpp:116-121), ten preset
 * control heights (:64-73), centres averaged only while still 0.0f (:131-132), edge midpoints displaced once
 * (:134-165) -- and the Cell attributes of Cloud::generateRecursive (:174-222; getPos :167-173; boxPos :106-113).
 * Recursion, visiting order and the `height == 0.0f` tests are kept literally.  The app never instantiates a Cloud
 * (Core.cpp:80-82); the class is restated for completeness of SURVEY section 8(f).
 * Pinned bit for bit against the reference's own functions (oracle/_ref, ref_cloud).
 */
#include "rn_oracle.h"
#include <math.h>
#include <stdlib.h>

typedef struct { float height, shade; } cloud_el; /* Cloud.h:21-35 */
typedef struct { float x, y, z; } k3;

typedef struct
{
  int depth, W, face;
  float amplitude;
  uint32_t state;
  uint64_t draws;
  cloud_el* base[2];
  rn_cell* out;
  uint64_t entries[32];
} cloud_ctx;

static float cloud_random(cloud_ctx* c)
{
  c->state = rno_lcg_next_state(c->state);
  c->draws++;
  return rno_random_from_draw((int)((c->state >> 16) & 0x7fffu));
}
/* Cloud.cpp:116-121 */
static cloud_el* element(cloud_ctx* c, int x, int y)
{
  const size_t P = (size_t)c->W + 1;
  if (c->face == 0 || x == 0 || y == 0 || x == c->W || y == c->W) return &c->base[0][(size_t)x * P + y];
  return &c->base[1][(size_t)x * P + y];
}
static const float kRadiosityStrength = 1.5f; /* Cloud.cpp:123 */

static void edge_mid(cloud_ctx* c, int x, int y, int ax, int ay, int bx, int by, int extent)
{
  cloud_el* el = element(c, x, y);
  if (el->height == 0.0f)
  {
    const cloud_el *a = element(c, ax, ay), *b = element(c, bx, by);
    el->height = (a->height + b->height) * 0.5f; /* Cloud.h:30-34 */
    el->shade = (a->shade + b->shade) * 0.5f;
    float offset = c->amplitude * cloud_random(c) * (float)extent / (float)c->W; /* Cloud.cpp:138 */
    el->height += offset;
    el->shade += offset * kRadiosityStrength;
  }
}
/* Cloud.cpp:124-171 */
static void generate_base_recursive(cloud_ctx* c, int minX, int minY, int maxX, int maxY)
{
  if ((maxX - minX) <= 1 && (maxY - minY) <= 1) return;
  const int midX = (minX + maxX) / 2, midY = (minY + maxY) / 2;
  cloud_el* m = element(c, midX, midY);
  if (m->height == 0.0f)
  {
    const cloud_el *a = element(c, minX, minY), *b = element(c, maxX, minY), *cc = element(c, minX, maxY), *d = element(c, maxX, maxY);
    m->height = (a->height + b->height + cc->height + d->height) * 0.25f; /* Cloud.h:25-29 */
    m->shade = (a->shade + b->shade + cc->shade + d->shade) * 0.25f;
  }
  edge_mid(c, minX, midY, minX, minY, minX, maxY, maxY - minY);
  edge_mid(c, maxX, midY, maxX, minY, maxX, maxY, maxY - minY);
  edge_mid(c, midX, minY, maxX, minY, minX, minY, maxX - minX);
  edge_mid(c, midX, maxY, maxX, maxY, minX, maxY, maxX - minX);
  generate_base_recursive(c, minX, minY, midX, midY);
  generate_base_recursive(c, midX, minY, maxX, midY);
  generate_base_recursive(c, minX, midY, midX, maxY);
  generate_base_recursive(c, midX, midY, maxX, maxY);
}
/* Cloud.cpp:106-113, :167-173 */
static k3 get_pos(cloud_ctx* c, int x, int y)
{
  k3 r;
  r.x = -1.f + 2.0f * (float)x / (float)c->W;
  r.y = 0.0f;
  r.z = -1.f + 2.0f * (float)y / (float)c->W;
  r.y += element(c, x, y)->height;
  return r;
}
static k3 k3_normalise(k3 v) /* member normalise(), Vector3.h:109-120 */
{
  float f = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
  if (f < 1e-10f) { k3 r = {0.0f, 0.0f, 1.0f}; return r; }
  float d = 1.0f / f;
  k3 r = {v.x * d, v.y * d, v.z * d};
  return r;
}
static float clampedf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* Cloud.cpp:174-222 */
static void generate_recursive(cloud_ctx* c, int has_parent, k3 parent_pos, int minX, int minY, int maxX, int maxY, int level, float size, float radius)
{
  uint64_t off = 0;
  for (int l = 0; l < level; l++) off += (uint64_t)1 << (2 * l);
  rn_cell* cell = &c->out[off + c->entries[level]];
  const int midX = (minX + maxX) / 2, midY = (minY + maxY) / 2;
  k3 centre = get_pos(c, midX, midY);
  centre.x *= size; centre.y *= size; centre.z *= size;
  const float shade = element(c, midX, midY)->shade;
  cell->colour[0] = clampedf(0.5f * shade, 0.0f, 1.0f);
  cell->colour[1] = clampedf(0.5f * shade, 0.0f, 1.0f);
  cell->colour[2] = clampedf(0.5f * shade, 0.0f, 1.0f);
  k3 position = centre;
  if (has_parent) { position.x -= parent_pos.x; position.y -= parent_pos.y; position.z -= parent_pos.z; }
  cell->position[0] = position.x; cell->position[1] = position.y; cell->position[2] = position.z;
  const k3 p00 = get_pos(c, minX, minY), p11 = get_pos(c, maxX, maxY), p10 = get_pos(c, maxX, minY), p01 = get_pos(c, minX, maxY);
  const k3 xc = {p00.x - p11.x, p00.y - p11.y, p00.z - p11.z};
  const k3 yc = {p10.x - p01.x, p10.y - p01.y, p10.z - p01.z};
  k3 normal = {(xc.y * yc.z) - (xc.z * yc.y), (xc.z * yc.x) - (xc.x * yc.z), (xc.x * yc.y) - (xc.y * yc.x)};
  if (c->face == 1) { normal.x = -normal.x; normal.y = -normal.y; normal.z = -normal.z; }
  normal = k3_normalise(normal);
  cell->normal[0] = (int8_t)(char)(normal.x * 127.0f);
  cell->normal[1] = (int8_t)(char)(normal.y * 127.0f);
  cell->normal[2] = (int8_t)(char)(normal.z * 127.0f);
  cell->radius = radius;
  cell->level = (uint8_t)level;
  c->entries[level]++;
  if (level == c->depth - 1) return;
  level++;
  radius /= 2.0f;
  /* visiting order of :217-220 (the child POINTERS are stored in decreasing slots there: children[3], [2], [1], [0]) */
  generate_recursive(c, 1, centre, minX, minY, midX, midY, level, size, radius);
  generate_recursive(c, 1, centre, midX, minY, maxX, midY, level, size, radius);
  generate_recursive(c, 1, centre, minX, midY, midX, maxY, level, size, radius);
  generate_recursive(c, 1, centre, midX, midY, maxX, maxY, level, size, radius);
}

/* faces: 2 x (W+1)^2 x {height, shade} as seen through Cloud::element; cells (optional): 2 x rno_cell_count(depth).
 * depth >= 2 (the presets of Cloud.cpp:64-73 need maxWidth/4 >= 1).  Returns the rand() calls consumed. */
long long rno_cloud(uint32_t state_in, int depth, float amplitude, float* faces, rn_cell* cells, uint32_t* state_out)
{
  if (depth < 2 || depth > 12) return -1;
  cloud_ctx c;
  c.depth = depth; c.W = 1 << depth; c.amplitude = amplitude; c.state = state_in; c.draws = 0; c.out = NULL;
  const int W = c.W;
  const size_t P = (size_t)W + 1, n = P * P;
  for (int f = 0; f < 2; f++)
  {
    c.base[f] = (cloud_el*)malloc(n * sizeof(cloud_el));
    if (!c.base[f]) return -2;
    for (size_t i = 0; i < n; i++) { c.base[f][i].height = 0.0f; c.base[f][i].shade = 1.0f; }
    const float v = f ? 0.5f : -0.5f; /* Cloud.cpp:64-73 */
    c.base[f][(size_t)(W / 2) * P + W / 2].height = v;
    c.base[f][(size_t)(W / 4) * P + W / 4].height = v;
    c.base[f][(size_t)(3 * W / 4) * P + W / 4].height = v;
    c.base[f][(size_t)(W / 4) * P + 3 * W / 4].height = v;
    c.base[f][(size_t)(3 * W / 4) * P + 3 * W / 4].height = v;
  }
  for (c.face = 0; c.face < 2; c.face++) generate_base_recursive(&c, 0, 0, W, W);
  if (faces)
    for (c.face = 0; c.face < 2; c.face++)
      for (int x = 0; x <= W; x++)
        for (int y = 0; y <= W; y++)
        {
          const cloud_el* e = element(&c, x, y);
          float* o = faces + (((size_t)c.face * P + (size_t)x) * P + (size_t)y) * 2;
          o[0] = e->height; o[1] = e->shade;
        }
  if (cells)
  {
    const uint64_t count = rno_cell_count(depth);
    for (c.face = 0; c.face < 2; c.face++)
    {
      for (int l = 0; l < 32; l++) c.entries[l] = 0;
      c.out = cells + (size_t)c.face * count;
      const float radius = 1.8f * sqrtf((1.0f + 0.5f * amplitude) * 0.5f); /* Cloud.cpp:84 */
      const k3 zero = {0.0f, 0.0f, 0.0f};
      generate_recursive(&c, 1, zero, 0, 0, W, W, 0, 1.0f, radius);
    }
  }
  for (int f = 0; f < 2; f++) free(c.base[f]);
  if (state_out) *state_out = c.state;
  return (long long)c.draws;
}
