/* oracle/rn_oracle_rock.c -- TEST INFRASTRUCTURE ONLY (see rn_oracle.h).
 *
 * Plain-C restatement of the Rock generator (renderEngine/source/Rock.cpp, include/Rock.h): the
 * recursive edge-midpoint displacement of a {height, shade} field on the six faces of a cube that
 * share their edges (generateBaseRecursive, Rock.cpp:173-218; element/boxElement, :127-170), and
 * the Cell tree computed from it (generateRecursive, :228-271; getPos, :220-227; boxPos, :100-124).
 * The recursion, its visiting order and the `height == 0.0f` "not set yet" test are kept literally:
 * they define which rand() call displaces which sample.  IEEE binary32 in the operand order of the
 * cited lines, no FMA contraction.  rand() is the MSVC CRT generator continued from `state_in`
 * (the reference does not re-seed for rocks: Rock.cpp:63 is commented out, Core.cpp:72,76).
 * Pinned bit for bit against the reference's own functions (oracle/_ref, ref_rock).
 */
#include "rn_oracle.h"
#include <math.h>
#include <stdlib.h>

typedef struct { float height, shade; } rock_el; /* Rock.h:21-35 */
typedef struct { float x, y, z; } r3;

typedef struct
{
  int depth, W, face;
  float amplitude;
  uint32_t state;
  uint64_t draws;
  rock_el* base[6]; /* baseArray[6][W+1][W+1], Rock.cpp:50-63 */
  /* cell tree */
  rn_cell* out;
  uint64_t entries[32];
  float size;
} rock_ctx;

static float rock_random(rock_ctx* c) /* basics.h:16-19 on the continued stream */
{
  c->state = rno_lcg_next_state(c->state);
  c->draws++;
  return rno_random_from_draw((int)((c->state >> 16) & 0x7fffu));
}

/* Rock.cpp:154-170 */
static rock_el* box_element(rock_ctx* c, int x, int y, int z)
{
  const int W = c->W, P = W + 1;
  if (x == 0) return &c->base[0][(size_t)y * P + z];
  else if (x == W) return &c->base[1][(size_t)y * P + z];
  else if (y == 0) return &c->base[2][(size_t)x * P + z];
  else if (y == W) return &c->base[3][(size_t)x * P + z];
  else if (z == 0) return &c->base[4][(size_t)x * P + y];
  else if (z == W) return &c->base[5][(size_t)x * P + y];
  return &c->base[0][0];
}
/* Rock.cpp:127-151 */
static rock_el* element(rock_ctx* c, int x, int y)
{
  int X = 0, Y = 0, Z = 0;
  const int face = c->face / 2, side = c->face % 2;
  if (face == 0) { X = side ? 0 : c->W; Y = x; Z = y; }
  else if (face == 1) { X = x; Y = side ? 0 : c->W; Z = y; }
  else if (face == 2) { X = x; Y = y; Z = side ? 0 : c->W; }
  return box_element(c, X, Y, Z);
}
static void average4(rock_el* e, const rock_el* a, const rock_el* b, const rock_el* cc, const rock_el* d) /* Rock.h:25-29 */
{
  e->height = (a->height + b->height + cc->height + d->height) * 0.25f;
  e->shade = (a->shade + b->shade + cc->shade + d->shade) * 0.25f;
}
static void average2(rock_el* e, const rock_el* a, const rock_el* b) /* Rock.h:30-34 */
{
  e->height = (a->height + b->height) * 0.5f;
  e->shade = (a->shade + b->shade) * 0.5f;
}
static const float kRadiosityStrength = 1.5f; /* Rock.cpp:171 */

static void edge_mid(rock_ctx* c, int x, int y, int ax, int ay, int bx, int by, int extent)
{
  rock_el* el = element(c, x, y);
  if (el->height == 0.0f)
  {
    average2(el, element(c, ax, ay), element(c, bx, by));
    float offset = c->amplitude * rock_random(c) * (float)extent / (float)c->W; /* Rock.cpp:185 */
    el->height += offset;
    el->shade += offset * kRadiosityStrength;
  }
}

/* Rock.cpp:173-218 */
static void generate_base_recursive(rock_ctx* c, int minX, int minY, int maxX, int maxY)
{
  if ((maxX - minX) <= 1 && (maxY - minY) <= 1) return;
  const int midX = (minX + maxX) / 2, midY = (minY + maxY) / 2;
  average4(element(c, midX, midY), element(c, minX, minY), element(c, maxX, minY), element(c, minX, maxY), element(c, maxX, maxY));
  edge_mid(c, minX, midY, minX, minY, minX, maxY, maxY - minY);
  edge_mid(c, maxX, midY, maxX, minY, maxX, maxY, maxY - minY);
  edge_mid(c, midX, minY, maxX, minY, minX, minY, maxX - minX);
  edge_mid(c, midX, maxY, maxX, maxY, minX, maxY, maxX - minX);
  generate_base_recursive(c, minX, minY, midX, midY);
  generate_base_recursive(c, midX, minY, maxX, midY);
  generate_base_recursive(c, minX, midY, midX, maxY);
  generate_base_recursive(c, midX, midY, maxX, maxY);
}

/* Rock.cpp:100-124 */
static r3 box_pos(const rock_ctx* c, int x, int y)
{
  float X = 0, Y = 0, Z = 0;
  const int face = c->face / 2, side = c->face % 2;
  const float fw = (float)c->W;
  if (face == 0) { X = side ? 0.f : 1.f; Y = (float)x / fw; Z = (float)y / fw; }
  else if (face == 1) { X = (float)x / fw; Y = side ? 0.f : 1.f; Z = (float)y / fw; }
  else if (face == 2) { X = (float)x / fw; Y = (float)y / fw; Z = side ? 0.f : 1.f; }
  r3 r = {X - 0.5f, Y - 0.5f, Z - 0.5f};
  return r;
}
/* member normalise(), Vector3.h:109-120 */
static r3 r3_normalise(r3 v)
{
  float f = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
  if (f < 1e-10f) { r3 r = {0.0f, 0.0f, 1.0f}; return r; }
  float d = 1.0f / f;
  r3 r = {v.x * d, v.y * d, v.z * d};
  return r;
}
/* Rock.cpp:220-227 */
static r3 get_pos(rock_ctx* c, int x, int y)
{
  r3 r = r3_normalise(box_pos(c, x, y));
  const float height = element(c, x, y)->height;
  const float s = 1.0f + height;
  r.x *= s; r.y *= s; r.z *= s;
  return r;
}
static float clampedf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); } /* basics.h:12-15 */

/* Rock.cpp:228-271 */
static void generate_recursive(rock_ctx* c, int has_parent, r3 parent_pos, int minX, int minY, int maxX, int maxY, int level, float size, float radius)
{
  uint64_t off = 0;
  for (int l = 0; l < level; l++) off += (uint64_t)1 << (2 * l);
  rn_cell* cell = &c->out[off + c->entries[level]];
  const int midX = (minX + maxX) / 2, midY = (minY + maxY) / 2;
  r3 centre = get_pos(c, midX, midY);
  centre.x *= size; centre.y *= size; centre.z *= size;
  r3 colour = {0.6f, 0.54f, 0.48f};
  const float shade = element(c, midX, midY)->shade;
  colour.x *= shade; colour.y *= shade; colour.z *= shade;
  cell->colour[0] = clampedf(colour.x, 0.0f, 1.0f);
  cell->colour[1] = clampedf(colour.y, 0.0f, 1.0f);
  cell->colour[2] = clampedf(colour.z, 0.0f, 1.0f);
  r3 position = centre;
  if (has_parent) { position.x -= parent_pos.x; position.y -= parent_pos.y; position.z -= parent_pos.z; }
  cell->position[0] = position.x; cell->position[1] = position.y; cell->position[2] = position.z;
  const r3 p00 = get_pos(c, minX, minY), p11 = get_pos(c, maxX, maxY), p10 = get_pos(c, maxX, minY), p01 = get_pos(c, minX, maxY);
  const r3 xc = {p00.x - p11.x, p00.y - p11.y, p00.z - p11.z};
  const r3 yc = {p10.x - p01.x, p10.y - p01.y, p10.z - p01.z};
  r3 normal = {(xc.y * yc.z) - (xc.z * yc.y), (xc.z * yc.x) - (xc.x * yc.z), (xc.x * yc.y) - (xc.y * yc.x)};
  if (normal.x * centre.x + normal.y * centre.y + normal.z * centre.z < 0.0f) { normal.x = -normal.x; normal.y = -normal.y; normal.z = -normal.z; }
  normal = r3_normalise(normal);
  cell->normal[0] = (int8_t)(char)(normal.x * 127.0f);
  cell->normal[1] = (int8_t)(char)(normal.y * 127.0f);
  cell->normal[2] = (int8_t)(char)(normal.z * 127.0f);
  cell->radius = radius;
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

uint64_t rno_rock_draw_count(int depth) /* every edge-midpoint of the cube surface once: 4W^2 - 4 */
{
  const uint64_t W = (uint64_t)1 << depth;
  return 4 * W * W - 4;
}

/* faces: 6 x (W+1) x (W+1) x {height, shade}: face f is the reference's currentFace, element (x, y) the value
 * Rock::element(x, y) sees (shared cube edges appear on both faces).  cells (optional): 6 x rno_cell_count(depth),
 * face-major, level-major inside a face, in the recursion's visiting order.  The root PhysicalCell is scene graph.
 * Returns the number of rand() calls consumed; *state_out = generator state afterwards. */
long long rno_rock(uint32_t state_in, int depth, float amplitude, float size, float* faces, rn_cell* cells, uint32_t* state_out)
{
  if (depth < 1 || depth > 12) return -1;
  rock_ctx c;
  c.depth = depth; c.W = 1 << depth; c.amplitude = amplitude; c.state = state_in; c.draws = 0; c.size = size; c.out = NULL;
  const size_t P = (size_t)c.W + 1, n = P * P;
  for (int f = 0; f < 6; f++)
  {
    c.base[f] = (rock_el*)malloc(n * sizeof(rock_el));
    if (!c.base[f]) return -2;
    for (size_t i = 0; i < n; i++) { c.base[f][i].height = 0.0f; c.base[f][i].shade = 1.0f; } /* Rock.cpp:57-60 */
  }
  for (c.face = 0; c.face < 6; c.face++) generate_base_recursive(&c, 0, 0, c.W, c.W); /* Rock.cpp:64-67 */
  if (faces)
    for (c.face = 0; c.face < 6; c.face++)
      for (int x = 0; x <= c.W; x++)
        for (int y = 0; y <= c.W; y++)
        {
          const rock_el* e = element(&c, x, y);
          float* o = faces + (((size_t)c.face * P + (size_t)x) * P + (size_t)y) * 2;
          o[0] = e->height; o[1] = e->shade;
        }
  if (cells)
  {
    const uint64_t count = rno_cell_count(depth);
    for (c.face = 0; c.face < 6; c.face++)
    {
      for (int l = 0; l < 32; l++) c.entries[l] = 0;
      c.out = cells + (size_t)c.face * count;
      const float radius = 1.4f * sqrtf((1.0f + 0.5f * amplitude) * 0.5f); /* Rock.cpp:76 */
      const r3 zero = {0.0f, 0.0f, 0.0f};
      generate_recursive(&c, 1, zero, 0, 0, c.W, c.W, 0, 1.0f, radius); /* Rock.cpp:77: parent = root, size argument 1.0f */
    }
  }
  for (int f = 0; f < 6; f++) free(c.base[f]);
  if (state_out) *state_out = c.state;
  return (long long)c.draws;
}

/* ------------------------------------------------------------------------------------------------
 * Index-addressed restatement: the same rock with every rand() call addressed by a closed-form call
 * index and the levels evaluated breadth-first, each face on its own (shared cube edges are computed
 * by every face that sees them, from the same call index).  This is the formulation the CUDA kernels
 * use; it must equal rno_rock bit for bit whenever no displaced height is exactly 0.0f (the
 * reference's "unset" sentinel would then draw again; returns -3 in that case).
 *
 * Visiting order (Rock.cpp:64-67,173-218): faces 0..5; inside a face the quadtree in pre-order, children
 * in Morton order (x is the low bit); at a node the four edge midpoints left, right, bottom, top, each
 * drawn only if still unset.  Same-level nodes are visited in Morton order, which is monotone in x and in
 * y, so a node finds its left and bottom midpoints already set by its neighbours unless it lies on the
 * face border, where the midpoint is set iff an earlier face shares that cube edge.
 * ------------------------------------------------------------------------------------------------ */
typedef struct { int pre[6][4]; uint64_t base[6]; } rock_topo; /* pre[f][side]: side (L, R, B, T) of face f belongs to an earlier face */

static void cube_of(int W, int f, int x, int y, int* X, int* Y, int* Z)
{
  const int a = f / 2, fixed = (f % 2) ? 0 : W;
  if (a == 0) { *X = fixed; *Y = x; *Z = y; }
  else if (a == 1) { *X = x; *Y = fixed; *Z = y; }
  else { *X = x; *Y = y; *Z = fixed; }
}
/* local coordinates of a cube-surface point on face g, if it lies on g */
static int face_local(int W, int g, int X, int Y, int Z, int* x, int* y)
{
  const int a = g / 2, fixed = (g % 2) ? 0 : W;
  if (a == 0) { if (X != fixed) return 0; *x = Y; *y = Z; }
  else if (a == 1) { if (Y != fixed) return 0; *x = X; *y = Z; }
  else { if (Z != fixed) return 0; *x = X; *y = Y; }
  return 1;
}
static uint64_t rock_E(int k) /* edge-midpoint samples in the closed square of a node of width 2^k */
{
  const uint64_t w = (uint64_t)1 << k;
  return (w + 1) * (w + 1) - 4 - ((((uint64_t)1 << (2 * k)) - 1) / 3);
}
static void rock_topology(int depth, rock_topo* t)
{
  const int W = 1 << depth;
  const int sx[4] = {0, W, W / 2, W / 2}, sy[4] = {W / 2, W / 2, 0, W};
  for (int f = 0; f < 6; f++)
    for (int s = 0; s < 4; s++)
    {
      int X, Y, Z, x, y;
      cube_of(W, f, sx[s], sy[s], &X, &Y, &Z);
      t->pre[f][s] = 0;
      for (int g = 0; g < f; g++)
        if (face_local(W, g, X, Y, Z, &x, &y)) t->pre[f][s] = 1;
    }
  uint64_t acc = 0;
  for (int f = 0; f < 6; f++)
  {
    t->base[f] = acc;
    acc += rock_E(depth) - (uint64_t)(W - 1) * (uint64_t)(t->pre[f][0] + t->pre[f][1] + t->pre[f][2] + t->pre[f][3]);
  }
}
/* draws of the node (level j, coordinates qx, qy) in slots < upto */
static int node_draws(const int* pre, int j, int qx, int qy, int upto)
{
  const int n = 1 << j;
  int d = 0;
  if (upto > 0 && qx == 0 && !pre[0]) d++;
  if (upto > 1 && !(qx == n - 1 && pre[1])) d++;
  if (upto > 2 && qy == 0 && !pre[2]) d++;
  if (upto > 3 && !(qy == n - 1 && pre[3])) d++;
  return d;
}
static uint64_t subtree_draws(const int* pre, int depth, int j, int qx, int qy)
{
  const int n = 1 << j, k = depth - j;
  const int preset = ((qx > 0) || pre[0]) + ((qx == n - 1) && pre[1]) + ((qy > 0) || pre[2]) + ((qy == n - 1) && pre[3]);
  return rock_E(k) - (((uint64_t)1 << k) - 1) * (uint64_t)preset;
}
/* call index (0-based, over the whole rock) of slot `slot` of node (level l, qx, qy) of face f */
static uint64_t rock_call_index(const rock_topo* t, int depth, int f, int l, int qx, int qy, int slot)
{
  const int* pre = t->pre[f];
  uint64_t idx = t->base[f];
  for (int j = 0; j < l; j++)
  {
    const int ax = qx >> (l - j), ay = qy >> (l - j); /* ancestor at level j */
    idx += (uint64_t)node_draws(pre, j, ax, ay, 4);
    const int cx = qx >> (l - j - 1), cy = qy >> (l - j - 1); /* path child at level j+1 */
    const int ci = ((cy & 1) << 1) | (cx & 1);
    for (int s = 0; s < ci; s++) idx += subtree_draws(pre, depth, j + 1, (ax << 1) | (s & 1), (ay << 1) | (s >> 1));
  }
  return idx + (uint64_t)node_draws(pre, l, qx, qy, slot);
}
/* call index of the edge-midpoint sample (x, y) of face f, whose level is l (step w = W >> l): resolved on the face
 * that draws it (the lowest-numbered face containing the point) */
static uint64_t rock_point_call_index(const rock_topo* t, int depth, int f, int l, int x, int y)
{
  const int W = 1 << depth, w = W >> l, n = 1 << l;
  int X, Y, Z;
  cube_of(W, f, x, y, &X, &Y, &Z);
  for (int g = 0; g < 6; g++)
  {
    int gx, gy;
    if (!face_local(W, g, X, Y, Z, &gx, &gy)) continue;
    if (gx % w == 0)
    { /* on a vertical node side: right slot of the node on its left, left slot of column 0 */
      const int qy = gy / w;
      if (gx == 0) return rock_call_index(t, depth, g, l, 0, qy, 0);
      return rock_call_index(t, depth, g, l, gx / w - 1, qy, 1);
    }
    const int qx = gx / w;
    if (gy == 0) return rock_call_index(t, depth, g, l, qx, 0, 2);
    return rock_call_index(t, depth, g, l, qx, gy / w - 1, 3);
  }
  (void)n;
  return 0;
}

long long rno_rock_indexed(uint32_t state_in, int depth, float amplitude, float* faces, uint32_t* state_out)
{
  if (depth < 1 || depth > 12 || !faces) return -1;
  const int W = 1 << depth;
  const size_t P = (size_t)W + 1;
  rock_topo t;
  rock_topology(depth, &t);
  int collision = 0;
  for (int f = 0; f < 6; f++)
  {
    float* g = faces + (size_t)f * P * P * 2;
#define EL(x, y) (g + ((size_t)(x) * P + (size_t)(y)) * 2)
    for (size_t i = 0; i < P * P; i++) { g[2 * i] = 0.0f; g[2 * i + 1] = 1.0f; }
    for (int l = 0; l < depth; l++)
    {
      const int w = W >> l, h = w / 2;
      for (int x = 0; x <= W; x += h)
        for (int y = 0; y <= W; y += h)
        {
          const int ox = (x / h) & 1, oy = (y / h) & 1;
          if (!ox && !oy) continue; /* set at an earlier level */
          float* e = EL(x, y);
          if (ox && oy)
          { /* centre: Rock.cpp:179 */
            const float *a = EL(x - h, y - h), *b = EL(x + h, y - h), *c = EL(x - h, y + h), *d = EL(x + h, y + h);
            e[0] = (a[0] + b[0] + c[0] + d[0]) * 0.25f;
            e[1] = (a[1] + b[1] + c[1] + d[1]) * 0.25f;
            continue;
          }
          const float *a = ox ? EL(x - h, y) : EL(x, y - h), *b = ox ? EL(x + h, y) : EL(x, y + h);
          const uint64_t k = rock_point_call_index(&t, depth, f, l, x, y);
          const float r = rno_random_from_draw((int)rno_lcg_draw(state_in, k));
          const float offset = amplitude * r * (float)w / (float)W;
          /* average(a, b) of Rock.h:30-34: the two-operand sum is commutative, so the order of the ends is immaterial */
          e[0] = (a[0] + b[0]) * 0.5f + offset;
          e[1] = (a[1] + b[1]) * 0.5f + offset * kRadiosityStrength;
          if (e[0] == 0.0f) collision = 1;
        }
    }
#undef EL
  }
  if (state_out) *state_out = rno_lcg_state_after(state_in, rno_rock_draw_count(depth));
  return collision ? -3 : (long long)rno_rock_draw_count(depth);
}
