// oracle/adapter_main.cpp -- TEST INFRASTRUCTURE ONLY: the drop-in claim, compiled and run.
//
// oracle/make_adapter.py applies the adapter block of INTEGRATION.md section 2 to a scratch copy of the reference's Terrain.cpp
// (lines 44-89 and 112-115 replaced, everything else untouched) and builds this program from it, the reference's other
// unmodified sources and librednoise_b200.so.  The program runs the REAL constructor Terrain::Terrain(depth, 500, 0.2f)
// (Core.cpp:65) -- whose field now comes from the GPU -- walks the Cell tree it built and compares every node with what the
// reference's own CPU passes + generateRecursive give for the same seed (ref_terrain_cells, ref_harness.cpp), plus volcanoPos
// and the position of the rand() stream afterwards (rocks continue it, Core.cpp:72,76).
//
//   adapter_check depth seed      exit 0 identical, 1 different, 77 no usable device
#include "ref_harness.cpp"
#include "../include/rednoise.h"
#include <vector>

static unsigned int g_adapter_seed = 12345u;
extern "C" time_t rn_adapter_time(time_t*) { return (time_t)g_adapter_seed; }

struct Rec { float pos[3], col[3], radius; signed char n[3]; unsigned char level; };

int main(int argc, char** argv)
{
  const int depth = argc > 1 ? atoi(argv[1]) : 8;
  g_adapter_seed = argc > 2 ? (unsigned)strtoul(argv[2], 0, 10) : 12345u;
  if (rn_device_count() < 1) { fprintf(stderr, "no CUDA device\n"); return 77; }
  const float size = 500.0f, amplitude = 0.2f;
  const int W = 1 << depth;
  // expected: the reference's own CPU code
  const size_t ncell = (size_t)rn_terrain2d_cell_count(depth);
  std::vector<Rec> want(ncell), got;
  std::vector<float> field((size_t)(W + 1) * (W + 1) * 4);
  if (ref_terrain_cells(depth, amplitude, size, g_adapter_seed, field.data(), want.data()) != (long long)ncell) { fprintf(stderr, "reference run failed\n"); return 1; }
  const float* c = &field[((size_t)(W / 2) * (W + 1) + (W / 2)) * 4];
  const float want_vpos[3] = {c[0] * size, c[1] * size, c[2] * size};  // Terrain.cpp:140 then :89
  // the adapted constructor
  g_seed_forced = 0;
  g_calls = 0;
  Terrain* t = new Terrain(depth, size, amplitude);
  const unsigned int state_after = g_state;
  // breadth-first from the root = level by level in the recursion's visiting order (children 4e .. 4e+3 of node e)
  std::vector<const Cell*> level(1, (const Cell*)t->root), next;
  for (int l = 0; l < depth; l++) {
    next.clear();
    for (const Cell* cell : level) {
      Rec r;
      r.pos[0] = cell->position.x; r.pos[1] = cell->position.y; r.pos[2] = cell->position.z;
      r.col[0] = cell->colour.x; r.col[1] = cell->colour.y; r.col[2] = cell->colour.z;
      r.radius = cell->radius;
      r.n[0] = cell->normal[0]; r.n[1] = cell->normal[1]; r.n[2] = cell->normal[2];
      r.level = (unsigned char)l;
      got.push_back(r);
      if (l + 1 < depth)
        for (int k = 0; k < 4; k++) next.push_back(cell->children[k]);
    }
    level.swap(next);
  }
  int bad = 0;
  if (got.size() != ncell) { printf("node count %zu, expected %zu\n", got.size(), ncell); bad = 1; }
  else if (memcmp(got.data(), want.data(), ncell * sizeof(Rec)) != 0) {
    size_t k = 0;
    while (k < ncell && memcmp(&got[k], &want[k], sizeof(Rec)) == 0) k++;
    printf("Cell tree DIFFERS, first at node %zu (level %d)\n", k, (int)want[k].level);
    bad = 1;
  }
  if (memcmp(&t->volcanoPos, want_vpos, 12) != 0) { printf("volcanoPos DIFFERS: %g %g %g vs %g %g %g\n", t->volcanoPos.x, t->volcanoPos.y, t->volcanoPos.z, want_vpos[0], want_vpos[1], want_vpos[2]); bad = 1; }
  const unsigned int want_state = rn_lcg_state_after(g_adapter_seed, rn_terrain2d_draw_count(depth));
  if (state_after != want_state) { printf("rand() stream position DIFFERS: state %08x, expected %08x\n", state_after, want_state); bad = 1; }
  printf("adapter: Terrain(%d, 500, 0.2f) seed %u: %zu cells, volcanoPos (%g, %g, %g), rand() state %08x -> %s\n", depth, g_adapter_seed, got.size(),
         t->volcanoPos.x, t->volcanoPos.y, t->volcanoPos.z, state_after, bad ? "DIFFERENT" : "IDENTICAL to the reference's CPU constructor");
  return bad;
}
