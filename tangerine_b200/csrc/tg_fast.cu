// FAST build of the brick kernel (TG_MESH_FAST): tg_bricks.cuh compiled once more, with FMA contraction, approximate
// sqrt / division (-use_fast_math) and float in place of the reference's double promotions (tg_sdf.h TG_FAST_MATH).
// Opt-in: the default path (tg_engine.cu) stays bit-identical to the reference; this one stays inside the north-star's
// tolerance (1e-5 relative / 4 ULP on raw samples, BASELINE.json) and is checked against it in tests/test_gpu_fast.py.
//
// Everything in this translation unit lives in namespace tg_fast, so that the two builds of every inline device function
// cannot meet at link time.
#define TG_FAST_MATH 1
#define tg tg_fast
#include "tg_bricks.cuh"
#undef tg

namespace tg
{

// Launchers for tg_engine.cu: MeshParams / DeviceModel / DeviceGrid have the same layout in both namespaces.
int LaunchMeshBricksFast(const void* mesh_params, unsigned blocks, void* stream)
{
	tg_fast::MeshBricksKernel<<<blocks, tg_fast::kBrickThreads, 0, static_cast<cudaStream_t>(stream)>>>(*static_cast<const tg_fast::MeshParams*>(mesh_params));
	return int(cudaGetLastError());
}

int LaunchLatticeFast(const void* device_model, const void* device_grid, float* out, unsigned tiles_x, unsigned tiles_y, unsigned tile_count, unsigned long long* counters, unsigned live, void* stream)
{
	tg_fast::LatticeKernel<<<(tile_count + tg_fast::kBrickWarps - 1) / tg_fast::kBrickWarps, tg_fast::kBrickThreads, 0, static_cast<cudaStream_t>(stream)>>>(
		*static_cast<const tg_fast::DeviceModel*>(device_model), *static_cast<const tg_fast::DeviceGrid*>(device_grid), out, tiles_x, tiles_y, tile_count, counters, live);
	return int(cudaGetLastError());
}

int FastBrickBlocksPerSm()
{
	int per_sm = 1;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tg_fast::MeshBricksKernel, tg_fast::kBrickThreads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
	return per_sm;
}

} // namespace tg
