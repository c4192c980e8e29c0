// Host-visible interface of the CUDA engine (implemented in tg_engine.cu).  The C ABI in tg_api.cpp is a
// thin layer over these classes; nothing here is exported from the shared library directly.
#pragma once

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
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
	void* cull_events[2] = { nullptr, nullptr }; // fork / join of the two kernels of a culling level
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
	// Fine-grained progress, written by the kernels themselves into page-locked host memory: [0] bricks evaluated,
	// [1] vertices through the refinement / attribute pass, both in 1/1024ths of a slab on top of 1024 x slabs done.
	volatile uint32_t* progress_words = nullptr;
	uint32_t progress_base = 0;   // 1024 x index of the slab being enqueued
	uint32_t progress_slabs = 1;  // slabs of the export in flight
	// largest word values a poll has seen: persistent warps finish their items out of order, so the words themselves can step back
	mutable std::atomic<uint32_t> progress_seen[2] = { { 0u }, { 0u } };

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
	std::shared_ptr<FlatModel> flat_owner; // one host copy of the tables, shared by the replicas of a multi-GPU model
	std::shared_ptr<const Tree> source;    // the CSG tree the tables were built from: reference statistics on demand
	float source_target_size = 0.25f;
	FlatModel& flat;
	Model* primary = nullptr; // replica: uploads from the primary's page-locked staging copies
	void* d_nodes = nullptr;
	void* d_interp = nullptr;
	void* d_tree = nullptr;
	void* d_materials = nullptr;
	void* d_regions = nullptr;
	void* d_node_rank = nullptr;
	void* d_node_material = nullptr;
	void* staging[7] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr }; // page-locked host copies of the seven tables
	uint64_t device_bytes = 0;
	double upload_seconds = 0.0;
	int leaf_count = 0;
	// slab cuts of the last multi-GPU export (the estimate walks every terminus cell: milliseconds, so it is made once per grid)
	struct { tg_grid grid; int ranks = 0; std::vector<uint32_t> cuts; std::vector<double> layer_cost; int feedback_rounds = 0; } plan;
	// the same estimate for the slab pipeline of a single-GPU export with host results (cuts by cost, not by thickness)
	struct { tg_grid grid; std::vector<double> layer_cost; bool valid = false; } pipeline_plan;
	// vertex / quad counts of this device's last slab of a multi-GPU export: capacities of the next one when the cuts moved
	uint64_t last_slab_vertices = 0, last_slab_quads = 0;

	Model() : flat_owner(std::make_shared<FlatModel>()), flat(*flat_owner) {}
	explicit Model(const std::shared_ptr<FlatModel>& shared) : flat_owner(shared), flat(*flat_owner) {}
	static Model* Create(Context* context, const Tree& tree, float target_size, int threads, std::string& error, bool live_octree = false);
	// The same tables on another device of a DeviceGroup (no second octree build).
	static Model* CreateReplica(Context* context, Model* primary, std::string& error);
	~Model();
};

// The devices of a multi-GPU context (tg_context_create_multi): one host thread per device and one NCCL communicator
// per device (ncclCommInitAll, one process).  libnccl is loaded at run time (dlopen), so single-GPU hosts need none.
class DeviceGroup
{
public:
	std::vector<Context*> contexts; // borrowed; contexts[0] is the primary
	std::vector<void*> comms;       // ncclComm_t, by rank
	std::string nccl_version;

	static DeviceGroup* Create(const std::vector<Context*>& contexts, std::string& error);
	~DeviceGroup();
	int size() const { return int(contexts.size()); }
	// Runs task(rank, error) on every worker thread and waits; returns the first non-zero status (with its message).
	int Run(const std::function<int(int, std::string&)>& task, std::string& error);
	// Rendezvous of the worker threads inside a task.
	void Barrier();

private:
	std::vector<std::thread> threads;
	std::mutex lock;
	std::condition_variable wake, done;
	const std::function<int(int, std::string&)>* task = nullptr;
	std::atomic<uint64_t> generation{ 0 };
	std::atomic<int> pending{ 0 };
	std::atomic<bool> quit{ false };
	std::vector<int> status;
	std::vector<std::string> errors;
	std::atomic<int> barrier_count{ 0 };
	std::atomic<int> barrier_phase{ 0 };
	void Worker(int rank);
};

struct MeshResultDevice; // opaque device-side result kept alive by tg_mesh.opaque

int EngineEvalPoints(Model* model, int mode, const float* points, uint64_t count, void* out, std::string& error);
// K0's cooperative evaluation of long programs against the plain interpreter: out = { probes, block mismatches, warp mismatches }.
int EngineCheckLongPrograms(Model* model, float reach, uint64_t out[3], std::string& error);
int EngineWeld(Context* ctx, const float* vertices, uint64_t count, float* out_vertices4, uint32_t* out_indices, uint64_t* out_unique, std::string& error);
int EngineRayMarch(Model* model, const float* rays, uint64_t count, int max_iterations, float epsilon, int magnet, float* out5, std::string& error);
int EngineEvalLattice(Model* model, const tg_grid& grid, uint32_t flags, float* out, float* out_ms, std::string& error);
int EngineExportMesh(Model* model, const tg_grid& grid, const tg_mesh_options& options, tg_mesh* out, std::string& error);
// The same export on all devices of a group: z-slabs cut from a host-side work estimate, per-slab vertex counts combined
// by an NCCL all-gather on the devices, every device copying straight into its slice of ONE host mesh.
int EngineExportMeshMulti(DeviceGroup* group, const std::vector<Model*>& models, const tg_grid& grid, const tg_mesh_options& options, tg_mesh* out, std::string& error);
// Per-rank detail of a multi-GPU export: returns the number of ranks (or -1), slab [begin, end) and stage timings of `rank`.
int EngineMeshRankInfo(const tg_mesh* mesh, int rank, uint64_t* slab_begin, uint64_t* slab_end, tg_mesh_timings* timings);
// Host-side slab planning (no device work): estimated cost of every cell layer, and the cuts of `ranks` z-slabs.
std::vector<double> EstimateLayerCost(const FlatModel& flat, const tg_grid& grid);
std::vector<uint32_t> PlanSlabs(const std::vector<double>& cost, uint32_t sz, int ranks);
int EngineTimerBeginMulti(DeviceGroup* group, std::string& error);
int EngineTimerEndMulti(DeviceGroup* group, float* out_ms, std::string& error);
int EngineExportPoints(Model* model, const float mn[3], const float mx[3], const float step[3], int refine, uint32_t flags, float scale, tg_mesh* out, std::string& error);
int EngineExportVoxels(Model* model, float grid_size, int32_t out_size[3], float* out_radius, int32_t** out_xyz, uint64_t* out_count, std::string& error);
void EngineFreeMesh(tg_mesh* mesh);
int EngineDownloadMesh(tg_mesh* mesh, uint32_t index_base, std::string& error);
int EngineTimerBegin(Context* context, std::string& error);
int EngineTimerEnd(Context* context, float* out_ms, std::string& error);
int EngineMeasureFp32Peak(Context* context, double* out_tflops, std::string& error);
int EngineFlushL2(Context* context, std::string& error);
int EngineSynchronize(Context* context, std::string& error);
// GetExportProgress's ratios (export.cpp:483-492) from the device-written words of these contexts (one, or a group's).
void EngineProgress(const std::vector<const Context*>& contexts, float out_ratios[4], int* out_stage);
int EngineUploadModel(Model* model, std::string& error);
int EngineBrickProfile(Model* model, const tg_grid& grid, uint32_t* out_layers, uint32_t layer_count, std::string& error);

} // namespace tg
