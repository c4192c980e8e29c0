// Host-visible interface of the CUDA engine (implemented in tg_engine.cu).  The C ABI in tg_api.cpp is a
// thin layer over these classes; nothing here is exported from the shared library directly.
#pragma once

#include <atomic>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/tangerine_b200.h"
#include "tg_octree.h"

namespace tg
{

struct PinnedBlock
{
	void* ptr = nullptr;
	size_t bytes = 0;
	bool in_use = false;
};

class Context
{
public:
	int device = 0;
	void* stream = nullptr;      // cudaStream_t
	void* stream2 = nullptr;     // cudaStream_t: second compute lane of the slab pipeline (odd slabs)
	void* copy_stream = nullptr; // cudaStream_t: device -> host copies that overlap compute
	void* copy_events[1] = { nullptr };
	void* timer_events[2] = { nullptr, nullptr };
	std::vector<PinnedBlock> pinned; // grow-only pool of page-locked host buffers for results
	std::vector<PinnedBlock> device_blocks; // grow-only cache of device buffers for results (no allocator call in steady state)
	std::vector<void*> mailboxes;    // free list of small page-locked blocks the device delivers its counts into
	void* index_base = nullptr;      // device: running vertex total of the slabs already emitted (pipelined export)
	// last whole-grid export on this context, so that the next identical one can size its host arrays exactly
	struct { const void* model = nullptr; uint64_t sx = 0, sy = 0, sz = 0; uint32_t flags = 0; uint64_t vertices = 0, triangles = 0; } last_export;
	void* arena = nullptr;       // grow-only device scratch reused by every call on this context
	size_t arena_bytes = 0;
	void* arena2 = nullptr;      // same for the second lane
	size_t arena2_bytes = 0;
	int sm_count = 0;
	int brick_blocks_per_sm = 1; // resident MeshBricksKernel blocks per SM (persistent grid size)

	// Export progress mirrors the reference's atomics (export.cpp:48-57).
	std::atomic<int> stage{ 0 };
	std::atomic<bool> active{ true };
	// Results handed to the caller (tg_mesh) keep blocks of this context's pinned / device caches: the context is only
	// torn down once the last of them was freed (tg_context_destroy on a context with live meshes defers to that moment).
	std::atomic<int> live_results{ 0 };
	std::atomic<bool> orphaned{ false };
	std::atomic<uint64_t> progress_done[4];
	std::atomic<uint64_t> progress_total[4];

	static Context* Create(int device, std::string& error);
	~Context();

	// vertex / quad counts of recent exports, so that a repeated export sizes its arrays exactly
	struct ExportHint { const void* model; uint64_t sx, sy, sz, k_begin, k_end; uint32_t flags; uint64_t vertices, quads; };
	std::vector<ExportHint> hints;

	void* AcquirePinned(size_t bytes, std::string& error);
	void ReleasePinned(void* ptr);
	// Result buffers.  Released only after every stream that touched them was synchronised, so a released block can
	// be handed out again at once.  (cudaMallocAsync in the export path queued behind NVML / driver locks and showed
	// up as milliseconds of idle GPU in short exports.)
	void* AcquireDevice(size_t bytes, std::string& error);
	void ReleaseDevice(void* ptr);
	void* AcquireMailbox(std::string& error);
	void ReleaseMailbox(void* ptr);
	bool Cancelled() const { return !active.load(); }
};

class Model
{
public:
	Context* context = nullptr;
	FlatModel flat;          // host copy of the tables (kept for stats and debugging)
	void* d_nodes = nullptr;
	void* d_interp = nullptr;
	void* d_tree = nullptr;
	void* d_materials = nullptr;
	void* d_regions = nullptr;
	void* d_node_rank = nullptr;
	void* staging[6] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr }; // page-locked host copies of the six tables
	uint64_t device_bytes = 0;
	double upload_seconds = 0.0;
	int leaf_count = 0;

	static Model* Create(Context* context, const Tree& tree, float target_size, int threads, std::string& error);
	~Model();
};

struct MeshResultDevice; // opaque device-side result kept alive by tg_mesh.opaque

int EngineEvalPoints(Model* model, int mode, const float* points, uint64_t count, void* out, std::string& error);
int EngineEvalLattice(Model* model, const tg_grid& grid, float* out, float* out_ms, std::string& error);
int EngineExportMesh(Model* model, const tg_grid& grid, const tg_mesh_options& options, tg_mesh* out, std::string& error);
int EngineExportPoints(Model* model, const float mn[3], const float mx[3], const float step[3], int refine, uint32_t flags, float scale, tg_mesh* out, std::string& error);
int EngineExportVoxels(Model* model, float grid_size, int32_t out_size[3], float* out_radius, int32_t** out_xyz, uint64_t* out_count, std::string& error);
void EngineFreeMesh(tg_mesh* mesh);
int EngineDownloadMesh(tg_mesh* mesh, uint32_t index_base, std::string& error);
int EngineTimerBegin(Context* context, std::string& error);
int EngineTimerEnd(Context* context, float* out_ms, std::string& error);
int EngineMeasureFp32Peak(Context* context, double* out_tflops, std::string& error);
int EngineFlushL2(Context* context, std::string& error);
int EngineSynchronize(Context* context, std::string& error);
int EngineUploadModel(Model* model, std::string& error);
int EngineBrickProfile(Model* model, const tg_grid& grid, uint32_t* out_layers, uint32_t layer_count, std::string& error);

} // namespace tg
