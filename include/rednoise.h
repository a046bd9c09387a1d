/* rednoise.h -- C ABI of rednoise_b200, the B200-native (sm_100a) drop-in for the
 * red-noise field-synthesis hot path of TGlad/redNoise.
 *
 * The reference has no plugin/FFI layer; the de-facto boundary this library
 * replaces is the field-synthesis half of the constructor
 *     Terrain::Terrain(int depth, Length size, Scale amplitude)
 *         renderEngine/include/Terrain.h:20, renderEngine/source/Terrain.cpp:15
 * i.e. grid initialisation (Terrain.cpp:44-56), seeding (:58-59), the
 * diamond-square level loop (:60-82 -> averageSquare :216-222, generateSquare
 * :224-258, averageDiamond :144-168, generateDiamond :171-214) and the
 * addVolcano epilogue (:88, :121-141).  INTEGRATION.md shows the adapter a
 * maintainer would add inside Terrain::Terrain.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, POD structs, int status codes
 *     (0 = ok).  Nothing throws across the boundary.  No CPU fallback: every
 *     compute entry point fails with RN_ERR_NO_DEVICE / RN_ERR_CUDA when no
 *     sm_100 device is usable.
 *   - A field is a row-major array of rn_element (16 B = the reference's
 *     Terrain::TerrainElement, Terrain.h:15-19), i-major: element (i,j) of the
 *     reference's baseArray[i][j] lives at dst[(i-i0)*pitch_elems + (j-j0)].
 *   - The caller owns dst; the library owns only its context and scratch.
 *   - Calls are blocking unless a stream is supplied with dst_is_device != 0,
 *     in which case work is enqueued on that stream and the call returns.
 *   - libc rand()/srand() state is never touched (the reference shares its
 *     rand() stream with later callers, Core.cpp:72,76; rn_terrain2d_draw_count
 *     tells an adapter how far to advance the host stream).
 */
#ifndef REDNOISE_H
#define REDNOISE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RN_VERSION 100 /* 0.1.0 */

/* status codes */
enum {
  RN_OK = 0,
  RN_ERR_BAD_PARAM = 1,   /* NULL pointer, depth out of range, bad window, ... */
  RN_ERR_UNSUPPORTED = 2, /* lacunarity != 2, ... */
  RN_ERR_NO_DEVICE = 3,   /* no CUDA device / not sm_100 */
  RN_ERR_CUDA = 4,        /* a CUDA runtime call failed (see rn_last_error_string) */
  RN_ERR_OOM = 5          /* device scratch allocation failed */
};

#define RN_MIN_DEPTH 3  /* Terrain.cpp:60: step = W/4 must be >= 2 for a level to exist */
#define RN_MAX_DEPTH 14 /* (2^14+1)^2 * 16 B = 4.3 GB; draw count stays < 2^32 */

/* One finished sample == Terrain::TerrainElement {Vector3 pos; float shade}
 * (renderEngine/include/Terrain.h:15-19), 16 bytes, written with one 128-bit store. */
typedef struct rn_element {
  float x, y, z, shade;
} rn_element;

/* The reference's generator call surface.  Defaults (rn_params_default) are the
 * literals the reference hard-codes, so a default call == the reference. */
typedef struct rn_params {
  uint32_t seed;              /* srand() argument, Terrain.cpp:58-59 (reference: time(NULL)) */
  int32_t depth;              /* "octave count": W = 1<<depth, depth-2 displacement levels (Terrain.h:20, Terrain.cpp:60-82) */
  float amplitude;            /* Terrain.h:20; 0.2f at Core.cpp:65 */
  float size;                 /* Terrain.h:20; 500 at Core.cpp:65; only scales volcano_pos (Terrain.cpp:89) */
  float persistence;          /* per-level gain; the reference fixes 0.5 via step/maxWidth (Terrain.cpp:186,231) */
  int32_t lacunarity;         /* per-level frequency ratio; the reference fixes 2 (Terrain.cpp:81). Only 2 is supported. */
  float sideways_scale;       /* 0.5f, Terrain.cpp:142 */
  float radiosity_strength;   /* 0.5f, Terrain.cpp:143 */
  float feature_scale;        /* 80.0f, Terrain.cpp:169 */
  float amplitude_variation;  /* 0.5f, Terrain.cpp:170 */
  int32_t bulgy_rocks;        /* 1: #define BULGY_ROCKS, Terrain.cpp:189 */
  int32_t apply_volcano;      /* 1: addVolcano runs, Terrain.cpp:88 */
} rn_params;

/* "grid extent and origin": a sub-window of the one global (W+1)x(W+1) lattice.
 * The window is seam-exact: its samples are bit-identical to the same indices
 * of the full field. */
typedef struct rn_window {
  int32_t i0, j0; /* first row (x index) and column (z index) */
  int32_t ni, nj; /* rows and columns */
} rn_window;

typedef struct rn_context rn_context;

/* ---- host-only helpers (no GPU needed) ---------------------------------- */
int rn_version(void);
const char* rn_last_error_string(void);
const char* rn_status_string(int status);
int rn_params_default(rn_params* p);
/* number of samples per side: (1<<depth)+1 */
int rn_terrain2d_side(int depth);
/* rand() calls the reference consumes for one Terrain of this depth:
 * sum over levels of 3n^2+2n. */
uint64_t rn_terrain2d_draw_count(int depth);
/* 0-based index of the rand() call that displaces lattice point (i,j) of a
 * depth-`depth` field, or UINT64_MAX for the 25 undisplaced points at
 * multiples of W/4 / out-of-range arguments. */
uint64_t rn_terrain2d_call_index(int depth, int i, int j);
/* MSVC CRT LCG (the reference's rand(), basics.h:18): state after `k` calls
 * from srand(seed), by skip-ahead; and the k-th (0-based) 15-bit draw. */
uint32_t rn_lcg_state_after(uint32_t seed, uint64_t k);
uint32_t rn_lcg_draw(uint32_t seed, uint64_t k);

/* ---- context ------------------------------------------------------------ */
/* One context per CUDA device (one process per GPU is the intended
 * deployment).  device < 0 selects the current device. */
int rn_create(int device, rn_context** out_ctx);
int rn_destroy(rn_context* ctx);
/* Kernels launched by this context since creation (for the bench's claim). */
uint64_t rn_launch_count(const rn_context* ctx);
/* Select the synthesis algorithm: 0 = auto (= 2), 1 = level-synchronous global-memory kernels (one
 * launch set per level), 2 = fused tile kernels for the coarse levels + the warp-synchronous march
 * kernel for the last two levels of fields of depth >= 10, 3 = as 2 with the round-1 block streaming
 * kernel for those levels.  1 and 3 are kept as independent cross-checks; all give the same bits. */
int rn_set_algorithm(rn_context* ctx, int algo);

/* Call pipelining (algorithm 2, on by default): the small coarse stages of a call run on the
 * context's high-priority side stream into one of two scratch slots, so back-to-back calls
 * overlap them with the output stage of the previous call.  Results are unaffected. */
int rn_set_pipelining(rn_context* ctx, int enable);

/* Stage profiling (algorithm 2): when enabled, every launch of the stage chain is bracketed by
 * CUDA events on the launching stream; rn_last_stage_times() then synchronizes and returns the
 * device time in ms of each launch of the most recent rn_terrain2d / rn_terrain2d_batch call
 * (ms[0..n-1], coarse to fine; the last one is the kernel that writes the output).  Returns the
 * number of launches, or a negative status.  Meant for roofline reporting, not for timed runs. */
int rn_set_profiling(rn_context* ctx, int enable);
int rn_last_stage_times(rn_context* ctx, float* ms, int cap);

/* ---- the hot path --------------------------------------------------------- */
/* Synthesize the field of Terrain(depth,size,amplitude) with srand(p->seed).
 *   win          NULL = whole (W+1)x(W+1) grid
 *   dst          ni*pitch_elems rn_element, device memory if dst_is_device else host
 *   pitch_elems  row pitch in elements (>= nj); 0 means nj
 *   stream       cudaStream_t (as void*) or NULL for the context's own stream.
 *   volcano_pos  optional float[3] (host): Terrain::volcanoPos after Terrain.cpp:89
 *                (only meaningful when apply_volcano; requires a blocking call) */
int rn_terrain2d(rn_context* ctx, const rn_params* p, const rn_window* win, rn_element* dst,
                 size_t pitch_elems, int dst_is_device, void* stream, float* volcano_pos);

/* `ntiles` independent full fields that share `p` except for the seed
 * (seeds[t]); tile t is written at dst + t*tile_stride_elems (0 = side*side). */
int rn_terrain2d_batch(rn_context* ctx, const rn_params* p, const uint32_t* seeds, int ntiles,
                       rn_element* dst, size_t tile_stride_elems, int dst_is_device, void* stream);

/* A batch is generated in chunks of whole fields (one launch chain per chunk; bounded intermediate grids, and a field count for
 * which the (strip, segment) units of the output level fill the GPU's warp slots evenly): the number of fields per chunk that
 * rn_terrain2d_batch uses for a device destination, 0 on bad arguments.  Informational (scratch sizing, profiling). */
int rn_terrain2d_batch_chunk(rn_context* ctx, const rn_params* p, int ntiles);

/* The integer part of the path, exposed for bit-exact parity tests: writes the
 * 15-bit draws k = first .. first+n-1 of srand(seed) as uint16. */
int rn_lcg_draws_device(rn_context* ctx, uint32_t seed, uint64_t first, uint64_t n, uint16_t* dst,
                        int dst_is_device, void* stream);

/* ---- next stage of the reference constructor: field -> Cell quadtree attributes -------------
 * Replaces the arithmetic of Terrain::generateRecursive (renderEngine/source/Terrain.cpp:260-356,
 * called at :96) + root->radius *= 4 (:97): for every node of the quadtree over the field (level l
 * has 4^l nodes, l = 0 .. depth-1; node (l, bx, by) covers lattice [bx*w, (bx+1)*w] x [by*w, (by+1)*w],
 * w = W >> l) the position relative to its parent, colour, quantised normal and radius.
 * Nodes are stored level by level; inside a level in the reference's visiting order (depth-first
 * with children ordered x-then-y, :351-354), i.e. Morton order with the x bit as the low bit of
 * every digit.  The scene-graph pointers (Cell::parent/children, Cell.h:12-19) are implied by
 * that order: child k of node e at level l is node 4e+k of level l+1. */
typedef struct rn_cell {
  float position[3]; /* Cell::position, w.r.t. the parent's centre (Terrain.cpp:325-328) */
  float colour[3];   /* Cell::colour (Terrain.cpp:290-324,337-338) */
  float radius;      /* Cell::radius (Terrain.cpp:342; root x4, :97) */
  int8_t normal[3];  /* Cell::normal, quantised (Terrain.cpp:330-332) */
  uint8_t level;
} rn_cell;

/* nodes of all levels: (4^depth - 1) / 3; first node of `level` */
uint64_t rn_terrain2d_cell_count(int depth);
uint64_t rn_terrain2d_cell_level_offset(int level);
/* field: the finished (W+1)x(W+1) field of `p` (row pitch field_pitch elements, 0 = W+1), device or host memory,
 * or NULL to synthesize it first (rn_terrain2d with the whole window).  cells: rn_terrain2d_cell_count(p->depth)
 * entries, device or host memory. */
int rn_terrain2d_cells(rn_context* ctx, const rn_params* p, const rn_element* field, size_t field_pitch, int field_is_device,
                       rn_cell* cells, int cells_is_device, void* stream);

/* ---- Rock generator ---------------------------------------------
 * Replaces the field synthesis and the per-node arithmetic of Rock::Rock (renderEngine/source/Rock.cpp:14-99):
 * generateBaseRecursive on the six cube faces (:64-67,173-218) and generateRecursive (:70-79,228-271).  The
 * reference does not seed for rocks (Rock.cpp:63 is commented out): they continue the CRT rand() stream, so the
 * caller passes the generator state at entry (srand(x): state x; after k calls: rn_lcg_state_after(x, k)) and gets
 * the state afterwards back (rn_rock_draw_count(depth) = 4 W^2 - 4 calls later).
 *   faces  6 x (W+1) x (W+1) x {height, shade}, W = 1 << depth: face f is the reference's currentFace, element
 *          (x, y) at faces[(f*(W+1) + x)*(W+1) + y] is what Rock::element(x, y) refers to (Rock.cpp:127-170; a sample
 *          on a cube edge appears on both faces that share it).  May be NULL.
 *   cells  6 x rn_terrain2d_cell_count(depth) rn_cell: per face the nodes level by level, inside a level in the
 *          recursion's visiting order (children 4e..4e+3 of node e), with position relative to the parent, colour,
 *          quantised normal and radius of Rock.cpp:236-262.  The root PhysicalCell (:69,82-89) is scene graph and
 *          stays with the caller.  May be NULL.
 * The reference decides with `height == 0.0f` whether an edge midpoint still has to be displaced (:181-212).  This
 * implementation addresses every rand() call by a closed-form index and evaluates the levels in parallel; when a
 * displaced height comes out exactly 0.0f (about 1 rock in 500) the reference's next visitor draws again and the call
 * sequence becomes value-dependent, so the rock is regenerated by the literal recursion on one device thread (slower,
 * same samples as the reference, more rand() calls than rn_rock_draw_count).  The call synchronises `stream`. */
#define RN_ROCK_MIN_DEPTH 1
#define RN_ROCK_MAX_DEPTH 12
typedef struct rn_rock_params {
  uint32_t lcg_state; /* CRT generator state when Rock::Rock is entered */
  int32_t depth;      /* Rock.h:10: detail; Core.cpp:76 uses 2..7 */
  float amplitude;    /* Core.cpp:76: 0.3f */
} rn_rock_params;
typedef struct rn_rock_element {
  float height, shade; /* Rock.h:21-24 */
} rn_rock_element;
uint64_t rn_rock_draw_count(int depth);
int rn_rock(rn_context* ctx, const rn_rock_params* p, rn_rock_element* faces, rn_cell* cells, int dst_is_device, void* stream,
            uint32_t* lcg_state_after);

/* The app's rock loop in one call (Core.cpp:69-79): `nrocks` rocks continue ONE rand() stream that starts in state `lcg_state`.
 * With depths == NULL the loop is the reference's: before each rock one rand() call gives
 * exponent = random(-5.0f, 0.0f) (basics.h:20-23) and detail = (int)(exponent + 7) (Core.cpp:72-75), then Rock(detail, amplitude, .)
 * consumes its calls; with depths != NULL rock i has depth depths[i] and no call is made in between.  Rocks of depth <= 7 (all of the
 * app's) are generated by ONE kernel launch for the whole batch (a CTA per face, the face in shared memory through all levels);
 * deeper ones and the rare value-dependent inputs (see rn_rock) go through rn_rock's path one by one.  Blocking.
 *   faces / cells   nrocks pointers (device or host memory by dst_is_device), each sized like rn_rock's; either array may be NULL
 *   depths_out, exponents_out, states_out   optional, nrocks entries: detail, exponent (rockRadius = 10 * powf(2, exponent),
 *                   Core.cpp:73-74) and the generator state at the entry of every rock
 *   lcg_state_after the state after the last rock */
#define RN_ROCK_BATCH_MAX 4096
int rn_rock_batch(rn_context* ctx, uint32_t lcg_state, int nrocks, const int32_t* depths, float amplitude, rn_rock_element* const* faces,
                  rn_cell* const* cells, int dst_is_device, void* stream, int32_t* depths_out, float* exponents_out, uint32_t* states_out,
                  uint32_t* lcg_state_after);

/* ---- Cloud generator --------------------------------------------
 * Cloud::Cloud (renderEngine/source/Cloud.cpp:14-105; the app never instantiates one, Core.cpp:80-82): the two-face
 * variant of Rock.  Same calling convention as rn_rock with 2 faces instead of 6: faces[(f*(W+1) + x)*(W+1) + y] is what
 * Cloud::element(x, y) refers to for currentFace f (the two sheets share their border, Cloud.cpp:116-121), cells =
 * 2 x rn_terrain2d_cell_count(depth) in the recursion's visiting order -- note that Cloud stores the child POINTERS in
 * decreasing slots (children[3 - k] = node 4e + k, Cloud.cpp:217-220).  depth >= 2 (the preset control heights of
 * :64-73 sit at maxWidth/4).  rn_cloud_draw_count(depth) rand() calls are consumed. */
#define RN_CLOUD_MIN_DEPTH 2
uint64_t rn_cloud_draw_count(int depth);
int rn_cloud(rn_context* ctx, const rn_rock_params* p, rn_rock_element* faces, rn_cell* cells, int dst_is_device, void* stream,
             uint32_t* lcg_state_after);

/* ---- 3D density volume: an EXTENSION, not a drop-in -------------------------------------------
 * The reference has no volume generator: there is nothing to be bit-exact with, so the
 * definition is this repository's own CPU specification oracle/rn_oracle3d.c -- 3D midpoint displacement in the
 * reference's style (cube centres, face centres, edge midpoints per level; amplitude law of Rock.cpp:185,
 * random() of basics.h:16-19; the draw of lattice point (i,j,k) is a 32-bit integer hash of (seed, i, j, k);
 * depth-1 levels on a (W+1)^3 lattice, W = 1 << depth).  The CUDA path reproduces that specification bit for bit.
 * The call produces planes [i0, i0+ni) of the volume into `slab`: DEVICE memory, ni x (W+1) x (W+1) floats, dense,
 * element (i,j,k) at slab[((i-i0)*(W+1)+j)*(W+1)+k] (rn_volume3d_slab_elems floats).  Slabs of one volume are
 * independent (what a slab needs of the coarser levels is recomputed on its device, in context-owned scratch), which
 * is how a volume shards across GPUs without a collective and why a 2049^3 volume (34.4 GB) never has to exist on
 * one device. */
#define RN_VOLUME_MIN_DEPTH 2
#define RN_VOLUME_MAX_DEPTH 11 /* 2049^3 floats = 34.4 GB */
typedef struct rn_volume_params {
  uint32_t seed;
  int32_t depth;
  float amplitude;
} rn_volume_params;
int rn_volume3d(rn_context* ctx, const rn_volume_params* p, int i0, int ni, float* slab, void* stream);
/* The same planes into HOST memory (pinned or pageable; same layout).  Blocking.  The slab is generated in chunks of planes
 * into two context-owned staging halves, so that the generation of chunk k+1 overlaps the device-to-host copy of chunk k
 * (the copy is the bottleneck: 4 B per sample over PCIe). */
int rn_volume3d_host(rn_context* ctx, const rn_volume_params* p, int i0, int ni, float* host_slab);
uint64_t rn_volume3d_slab_elems(int depth, int ni);

/* ---- memory helpers ---------------------------------------------------------------------------
 * So that a C/C++ caller (the adapter inside Terrain::Terrain) needs no CUDA runtime of its own:
 * device memory on a context's device, page-locked host memory (a pinned destination lets rn_terrain2d's device-to-host
 * copies run at full PCIe rate and asynchronously; `new TerrainElement[]`, Terrain.cpp:45-49, is pageable), and a blocking
 * device-to-host copy. */
int rn_device_count(void); /* CUDA devices visible (0 without a driver) */
int rn_device_alloc(rn_context* ctx, size_t bytes, void** out);
int rn_device_free(rn_context* ctx, void* ptr);
int rn_host_alloc(size_t bytes, void** out);
int rn_host_free(void* ptr);
int rn_copy_to_host(rn_context* ctx, void* dst_host, const void* src_device, size_t bytes);

/* ---- several GPUs from one host thread ----------------------------------------------------------
 * The path shards with no exchange step: every sample is a pure function of (seed, params, lattice
 * index), and what a slab needs beyond its own rows (its dependency cone through the coarser levels) is RECOMPUTED on its
 * device, never communicated.  An rn_multi owns one rn_context (own streams, own scratch) per listed device; the calls below
 * enqueue the work of every device from the calling thread and then wait for all of them.  The reference's one caller
 * (Core::init -> new Terrain, Core.cpp:58-65) is a single host thread, which is why this is not one-process-per-GPU only.
 *   rn_slab_rows / rn_slab_window   n equal row slabs of a (side x side) field: slab k = rows [k*rows, min(side, (k+1)*rows));
 *                                   trailing slabs may be shorter or empty (ni == 0)
 *   rn_terrain2d_sharded            ONE field: slab k is synthesized on device k.
 *       slab_dst   n device pointers (slab_dst[k] on device k, row pitch slab_pitch_elems, 0 = side), or NULL
 *       host_dst   the whole field assembled in HOST memory (row pitch host_pitch_elems, 0 = side): every device sends its
 *                  slab straight into its rows, all devices concurrently; or NULL
 *   rn_terrain2d_batch_sharded      `ntiles` independent fields, tile t on device t % n (round-robin): tile t is written at
 *                                   tile_dst[t % n] + (t / n) * tile_stride_elems (device memory of that device)
 *   rn_gather                       the optional final assembly (the only collective of the path): every device ends up with
 *                                   all n slabs, slab j at full_dst[k] + j * rows * side.  slab_src[k] holds rn_slab_rows x side
 *                                   elements (dense); full_dst[k] n * rn_slab_rows x side.  use_nccl != 0: one ncclAllGather per
 *                                   device inside a group (NCCL over NVLink / NVSwitch; the library binds libnccl.so.2 at run
 *                                   time, RN_ERR_UNSUPPORTED if it is absent or a device is listed twice); 0: peer copies
 *                                   (cudaMemcpyPeerAsync), which also work without NCCL. */
#define RN_MAX_DEVICES 16
typedef struct rn_multi rn_multi;
int rn_create_multi(const int* devices, int n, rn_multi** out_multi);
int rn_destroy_multi(rn_multi* m);
int rn_multi_size(const rn_multi* m);
rn_context* rn_multi_context(rn_multi* m, int k);
int rn_slab_rows(int side, int nparts);
int rn_slab_window(int side, int part, int nparts, rn_window* out);
int rn_terrain2d_sharded(rn_multi* m, const rn_params* p, rn_element* const* slab_dst, size_t slab_pitch_elems, rn_element* host_dst,
                         size_t host_pitch_elems);
int rn_terrain2d_batch_sharded(rn_multi* m, const rn_params* p, const uint32_t* seeds, int ntiles, rn_element* const* tile_dst,
                               size_t tile_stride_elems);
int rn_gather(rn_multi* m, const rn_params* p, const rn_element* const* slab_src, rn_element* const* full_dst, int use_nccl);

/* Device self-test of the arithmetic shortcuts: the normal-range sqrt / reciprocal sequences
 * used by the fused kernels are compared with the IEEE-rounded intrinsics over every binary32
 * value in 2^-100 .. 2^100.  mismatches[0..2] = sqrt, reciprocal, 1/sqrt (all must be 0). */
int rn_selftest_math(rn_context* ctx, uint64_t* mismatches3);

#ifdef __cplusplus
}
#endif
#endif /* REDNOISE_H */
