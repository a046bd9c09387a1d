/* oracle/rn_oracle_cells.c -- TEST INFRASTRUCTURE ONLY (see rn_oracle.h).
 *
 * Plain-C restatement of Terrain::generateRecursive (renderEngine/source/Terrain.cpp:260-356) and
 * root->radius *= 4 (:97): the Cell-quadtree attributes computed from the finished field.  The
 * recursion is kept (depth-first, children in the order of :351-354) so that the visiting order
 * -- which defines where a node lands inside its level -- is the reference's own.  IEEE binary32
 * in the operand order of the cited lines, no FMA contraction.
 * Pinned bit for bit against the reference's own generateRecursive (oracle/_ref, ref_terrain_cells).
 */
#include "rn_oracle.h"
#include <math.h>

typedef struct { float x, y, z; } c3;

static c3 c3_add(c3 a, c3 b) { c3 r = {a.x + b.x, a.y + b.y, a.z + b.z}; return r; }      /* Vector3.h:61-64 */
static c3 c3_sub(c3 a, c3 b) { c3 r = {a.x - b.x, a.y - b.y, a.z - b.z}; return r; }      /* Vector3.h:75-78 */
static c3 c3_mul(c3 a, float f) { c3 r = {a.x * f, a.y * f, a.z * f}; return r; }         /* Vector3.h:31-34,39-44 */
static c3 c3_cross(c3 a, c3 b)                                                             /* Vector3.h:130-135 */
{
  c3 r = {(a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x)};
  return r;
}
/* member normalise(), Vector3.h:109-120: f = sqrtf(|v|^2); f < 1e-10 -> (0,0,1); v *= 1.0f/f */
static c3 c3_normalise(c3 v)
{
  float f = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
  if (f < 1e-10f) { c3 r = {0.0f, 0.0f, 1.0f}; return r; }
  float d = 1.0f / f;
  c3 r = {v.x * d, v.y * d, v.z * d};
  return r;
}
static float clampedf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); } /* basics.h:12-15 */

typedef struct
{
  const rn_params* p;
  const rn_element* g;
  size_t pitch;
  int depth;
  rn_cell* out;
  uint64_t entries[32]; /* numEntries[level], Terrain.cpp:343 */
} cell_ctx;

static c3 pos_at(const cell_ctx* c, int i, int j)
{
  const rn_element* e = &c->g[(size_t)i * c->pitch + (size_t)j];
  c3 r = {e->x, e->y, e->z};
  return r;
}
static float shade_at(const cell_ctx* c, int i, int j) { return c->g[(size_t)i * c->pitch + (size_t)j].shade; }

/* Terrain.cpp:260-356 */
static void generate_recursive(cell_ctx* c, int has_parent, c3 parent_pos, int minX, int minY, int maxX, int maxY, int level, float size,
                               float radius)
{
  const rn_params* p = c->p;
  uint64_t off = 0;
  for (int l = 0; l < level; l++) off += (uint64_t)1 << (2 * l);
  rn_cell* cell = &c->out[off + c->entries[level]];
  const int midX = (minX + maxX) / 2, midY = (minY + maxY) / 2;
  c3 centre = c3_mul(c3_add(c3_add(c3_add(c3_add(pos_at(c, midX, midY), pos_at(c, minX, minY)), pos_at(c, maxX, maxY)), pos_at(c, minX, maxY)),
                            pos_at(c, maxX, minY)),
                     0.2f);
  centre = c3_mul(centre, p->feature_scale);
  c3 dir1 = c3_sub(pos_at(c, maxX, maxY), pos_at(c, minX, minY));
  c3 dir2 = c3_sub(pos_at(c, maxX, minY), pos_at(c, minX, maxY));
  c3 normal = c3_normalise(c3_cross(dir1, dir2));
  float shade = (2 * shade_at(c, midX, midY) + shade_at(c, minX, minY) + shade_at(c, maxX, maxY) + shade_at(c, minX, maxY) + shade_at(c, maxX, minY)) / 6.0f;
  if (centre.y < 0.0f)
  {
    centre.y *= 0.3f;
    shade = (shade - 1.0f) * 0.3f + 1.0f;
    normal.y /= 0.3f;
    normal = c3_normalise(normal);
  }
  c3 colour;
  if (centre.y < -0.2f)
  {
    float depth = -0.2f - centre.y;
    centre.y = (centre.y + 0.2f) * 0.6f - 0.2f;
    float g = 0.9f * (1.0f - 0.3f * depth);
    colour.x = 0.0f;
    colour.y = g > 0.0f ? g : 0.0f; /* max(0.0f, g), basics.h:24-27 */
    colour.z = 0.7f;
    float div = 1.0f / (1.0f + 0.5f * depth); /* operator/=, Vector3.h:55-60 */
    colour = c3_mul(colour, div);
  }
  else if (centre.y < 0.0f)
  {
    float height = centre.y / 0.4f + 0.3f;
    colour.x = 0.95f - 0.25f * height;
    colour.y = 0.9f - 0.4f * height;
    colour.z = 0.25f;
  }
  else
  {
    float flatness = normal.y - 1.0f;
    float height = centre.y - 0.0f;
    c3 tree = {0.0f, 0.45f + 0.04f * height, 0.03f * height};
    float snowHeight = centre.y - 3.0f;
    c3 snow = {0.65f + 0.04f * snowHeight, 0.65f + 0.04f * snowHeight, 0.5f + 0.07f * snowHeight};
    float dif = (flatness + 0.5f) * 0.75f;
    if (snowHeight > -0.5f) dif -= (snowHeight + 0.5f) * 0.1f;
    float blend = clampedf(dif * 2.0f, 0.0f, 1.0f);
    colour = c3_add(c3_mul(snow, 1.0f - blend), c3_mul(tree, blend));
  }
  centre = c3_mul(centre, size / p->feature_scale);
  c3 position = centre;
  if (has_parent) position = c3_sub(position, parent_pos);
  cell->position[0] = position.x; cell->position[1] = position.y; cell->position[2] = position.z;
  cell->normal[0] = (int8_t)(char)(normal.x * 127.0f);
  cell->normal[1] = (int8_t)(char)(normal.y * 127.0f);
  cell->normal[2] = (int8_t)(char)(normal.z * 127.0f);
  colour = c3_mul(colour, shade);
  cell->colour[0] = clampedf(colour.x, 0.0f, 1.0f);
  cell->colour[1] = clampedf(colour.y, 0.0f, 1.0f);
  cell->colour[2] = clampedf(colour.z, 0.0f, 1.0f);
  cell->radius = sqrtf((2.0f + p->amplitude) * radius * radius);
  cell->level = (uint8_t)level;
  c->entries[level]++;
  if (level == c->depth - 1) return;
  level++;
  radius /= 2.0f;
  generate_recursive(c, 1, centre, minX, minY, midX, midY, level, size, radius);
  generate_recursive(c, 1, centre, midX, minY, maxX, midY, level, size, radius);
  generate_recursive(c, 1, centre, minX, midY, midX, maxY, level, size, radius);
  generate_recursive(c, 1, centre, midX, midY, maxX, maxY, level, size, radius);
}

uint64_t rno_cell_count(int depth)
{
  uint64_t n = 0;
  for (int l = 0; l < depth; l++) n += (uint64_t)1 << (2 * l);
  return n;
}

/* field: the finished (post-volcano) field of p, dense (W+1)^2; cells: rno_cell_count(depth) entries */
int rno_terrain2d_cells(const rn_params* p, const rn_element* field, rn_cell* cells)
{
  if (!p || !field || !cells || p->depth < RN_MIN_DEPTH || p->depth > RN_MAX_DEPTH) return RN_ERR_BAD_PARAM;
  cell_ctx c;
  c.p = p;
  c.g = field;
  c.depth = p->depth;
  c.pitch = ((size_t)1 << p->depth) + 1;
  c.out = cells;
  for (int l = 0; l < 32; l++) c.entries[l] = 0;
  const int W = 1 << p->depth;
  c3 zero = {0.0f, 0.0f, 0.0f};
  generate_recursive(&c, 0, zero, 0, 0, W, W, 0, p->size, p->size);
  cells[0].radius *= 4.0f; /* Terrain.cpp:97 */
  return RN_OK;
}
