// rn_kernels.h -- internal interface between the C ABI (rn_abi.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "rn_math.cuh"

namespace rn {

// Algorithm 1 (rn_levelsync.cu): levels [0,nlevels) of a depth-`depth` field in global memory.
cudaError_t levelsync_generate(Elem* g, float* off, size_t pitch, int depth, int nlevels, const Law& law, float persistence,
                               uint32_t seed, cudaStream_t stream, uint64_t* launches);

// Algorithm 2 (rn_fused.cu): fused tile kernels + streamed last levels, 2-6 launches per field; writes only the requested window.
// march: streamed levels run the warp-synchronous kernel (rn_march.cuh), else the round-1 block kernel (rn_stream.cuh).
size_t fused_scratch_bytes(int depth, int nfields);
int fused_batch_chunk(int depth, int cmax, int sm_count);  // fields per chunk of a batch (rn_plan.h march_batch_chunk)
cudaError_t fused_generate(Elem* dst, size_t pitch, size_t field_stride, int depth, int i0, int j0, int ni, int nj, const Law& law,
                           float persistence, uint32_t seed, const uint32_t* seeds, int nfields, int volcano, float amplitude,
                           void* scratch, cudaStream_t stream, uint64_t* launches, cudaEvent_t* events = nullptr, int* nstages = nullptr,
                           cudaStream_t coarse_stream = nullptr, cudaEvent_t coarse_done = nullptr, cudaStream_t tile_stream = nullptr,
                           cudaEvent_t tiles_done = nullptr, int sm_count = 0, int march = 1);

// Cell-quadtree attributes of a finished field (rn_cells.cu; Terrain.cpp:260-356): `cells` = rn_cell[(4^depth-1)/3].
cudaError_t cells_generate(void* cells, const Elem* field, size_t pitch, int depth, float feature_scale, float size, float amplitude,
                           cudaStream_t stream, uint64_t* launches);

// Rock generator (rn_rock.cu; Rock.cpp:64-67,173-271): faces = 6 x (W+1)^2 float2 {height, shade}, cells = 6 x (4^depth-1)/3 rn_cell or
// nullptr, flag = device int set when a displaced height is exactly 0.0f (the reference's "unset" sentinel).
// kind 0 = Rock (6 faces), 1 = Cloud (Cloud.cpp: 2 sheets sharing their border, preset control heights).
uint64_t rock_draw_count(int depth);
uint64_t cloud_draw_count(int depth);
// flag: 16 bytes of device memory.  sequential = 0: the index-addressed parallel path, flag[0] (int) is set when a displaced height is exactly
// 0.0f (then the result is not the reference's); sequential = 1: the literal one-thread recursion, *(uint64*)flag = rand() calls made.
cudaError_t rock_generate(void* faces, void* cells, int* flag, int kind, int depth, uint32_t state, float amplitude, int sequential,
                          cudaStream_t stream, uint64_t* launches);

// n rocks of depth <= 7 in one launch (one CTA per face, the face in shared memory through all levels): see rn_rock.cu
size_t rock_small_item_bytes();
cudaError_t rock_small_batch(int n, const int* depths, const uint32_t* states, float amplitude, void* const* faces, void* const* cells, int* flags,
                             void* items_dev, void* items_host, cudaStream_t stream, uint64_t* launches);

// 3D density-volume extension (rn_volume.cu; defined by the CPU specification of DESIGN.md section 5.8, no reference counterpart):
// planes [i0, i0+ni) of the (W+1)^3 float volume into `slab` (ni x (W+1) x (W+1), dense); scratch: volume_scratch_bytes() bytes.
size_t volume_scratch_bytes(int depth, int i0, int ni);
cudaError_t volume_generate(float* slab, void* scratch, int depth, uint32_t seed, float amplitude, int i0, int ni, cudaStream_t stream,
                            uint64_t* launches, cudaEvent_t* events = nullptr, int* nstages = nullptr, cudaStream_t coarse = nullptr,
                            cudaEvent_t coarse_done = nullptr);

// Epilogues / utilities (rn_util.cu).
// addVolcano (Terrain.cpp:121-141) over the window [i0,i0+ni) x [j0,j0+nj) of a depth-`depth` field
// stored at g with row pitch `pitch` (g points at window element (i0,j0)).
cudaError_t volcano_apply(Elem* g, size_t pitch, int depth, int i0, int j0, int ni, int nj, float amplitude, cudaStream_t stream,
                          uint64_t* launches);
// Exhaustive device check of sqrt_rn_normal / rcp_rn_normal against the IEEE intrinsics (3 mismatch counters).
cudaError_t selftest_math(unsigned long long* dev_counts3, cudaStream_t stream, uint64_t* launches);
// 15-bit draws first..first+n-1 of srand(seed) as uint16.
cudaError_t lcg_draws(uint16_t* dst, uint32_t seed, uint64_t first, uint64_t n, cudaStream_t stream, uint64_t* launches);

}  // namespace rn
