/* oracle/rn_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the reference's field-synthesis path, used as the
 * parity checker for the CUDA path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may call into it.  The product
 * (rednoise_b200/) never links or loads anything under oracle/.
 *
 * Parity status: PINNED for 2D -- the restatement is checked bit-for-bit against
 * the reference's own unmodified sources compiled into oracle/_ref/ (see
 * ref_harness.cpp, tests/test_oracle_vs_ref.py) and against the committed
 * digests in tests/golden/ that were produced by that build.
 * 3D (rno_volume3d*) has NO reference counterpart: parity unpinned by the
 * reference, pinned only against this file's own golden digests.
 */
#ifndef RN_ORACLE_H
#define RN_ORACLE_H
#include <stdint.h>
#include <stddef.h>
#include "../include/rednoise.h"

#ifdef __cplusplus
extern "C" {
#endif

/* MSVC CRT LCG (not in the reference tree; VS2008 runtime, renderEngine.vcproj:4). */
uint32_t rno_lcg_next_state(uint32_t state);
uint32_t rno_lcg_state_after(uint32_t seed, uint64_t k); /* skip-ahead */
uint32_t rno_lcg_draw(uint32_t seed, uint64_t k);        /* k-th 15-bit draw, 0-based */
float rno_random_from_draw(int r);                       /* basics.h:16-19 */

uint64_t rno_draw_count(int depth);
uint64_t rno_call_index(int depth, int i, int j);

/* Sequential restatement: reference loop order, sequential LCG (Terrain.cpp:44-82). */
int rno_terrain2d_seq(const rn_params* p, rn_element* grid);
/* Index-addressed restatement: RNG by call index + skip-ahead, shade by
 * ordered gather (A.3), samples of each stage visited in reversed order when
 * `reverse` != 0.  Must equal rno_terrain2d_seq bit for bit. */
int rno_terrain2d_indexed(const rn_params* p, rn_element* grid, int reverse);
/* Terrain.cpp:121-141; volcano_pos (optional float[3]) = volcanoPos after Terrain.cpp:89. */
void rno_add_volcano(const rn_params* p, rn_element* grid, float* volcano_pos);
/* seq + optional volcano (p->apply_volcano). */
int rno_terrain2d(const rn_params* p, rn_element* grid, float* volcano_pos);
/* ntiles independent fields, one per seed, on up to `threads` host threads (OpenMP). */
int rno_terrain2d_batch(const rn_params* p, const uint32_t* seeds, int ntiles, rn_element* grids, int threads);

/* Terrain::generateRecursive (Terrain.cpp:260-356) + root radius (:97): Cell-quadtree attributes from the finished field. */
uint64_t rno_cell_count(int depth);
int rno_terrain2d_cells(const rn_params* p, const rn_element* field, rn_cell* cells);

/* Rock generator (rn_oracle_rock.c): Rock::generateBaseRecursive on the six cube faces + Rock::generateRecursive. */
uint64_t rno_rock_draw_count(int depth);
long long rno_rock(uint32_t state_in, int depth, float amplitude, float size, float* faces, rn_cell* cells, uint32_t* state_out);
/* index-addressed, breadth-first restatement (the CUDA formulation); -3 if a displaced height is exactly 0.0f */
long long rno_rock_indexed(uint32_t state_in, int depth, float amplitude, float* faces, uint32_t* state_out);

/* Cloud generator (rn_oracle_cloud.c): the two-face variant of Rock (Cloud.cpp). */
long long rno_cloud(uint32_t state_in, int depth, float amplitude, float* faces, rn_cell* cells, uint32_t* state_out);

/* 3D density-volume extension (rn_oracle3d.c): NO reference counterpart, parity unpinned by the reference. */
uint32_t rno_volume_draw(uint32_t seed, uint32_t i, uint32_t j, uint32_t k); /* the integer lattice hash: 15-bit draw of lattice point (i,j,k) */
/* planes [i0, i0+ni) of the (W+1)^3 volume into `slab` (ni x (W+1) x (W+1) floats, dense) */
int rno_volume3d(uint32_t seed, int depth, float amplitude, int i0, int ni, float* slab);

uint64_t rno_fnv64(const void* data, size_t nbytes);

#ifdef __cplusplus
}
#endif
#endif
