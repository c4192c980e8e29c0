// Multi-GPU export inside the library (included at the end of tg_engine.cu, inside namespace tg).
//
// The reference's export is ONE call that ends with ONE mesh in one address space handed to the writer
// (tangerine/export.cpp:320-381, writer at :283).  A multi-device context (tg_context_create_multi) keeps that shape:
// tg_export_mesh cuts the grid into z-slabs, one per GPU, each driven by its own host thread and
// stream; the only exchange is an ncclAllGather of the per-slab vertex counts on the devices, from which every GPU
// rebases its triangle indices itself; then each GPU copies its arrays straight into its slice of one page-locked
// host mesh -- that is the "stitch".  One process, ncclCommInitAll, no torch.distributed anywhere.

#include <dlfcn.h>
#include <nccl.h>

namespace
{

// libnccl is resolved at run time: the library must load (and serve single-GPU callers) on hosts without NCCL.
struct NcclApi
{
	void* handle = nullptr;
	ncclResult_t (*GetVersion)(int*) = nullptr;
	ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* LoadNccl(std::string& error)
{
	static NcclApi api;
	static std::mutex once;
	std::lock_guard<std::mutex> guard(once);
	if (api.handle) return &api;
	const char* names[] = { "libnccl.so.2", "libnccl.so" };
	void* h = nullptr;
	for (const char* name : names)
	{
		if ((h = dlopen(name, RTLD_NOW | RTLD_GLOBAL)) != nullptr) break;
	}
	if (!h)
	{
		error = std::string("multi-GPU contexts need NCCL: ") + (dlerror() ? dlerror() : "libnccl.so.2 not found");
		return nullptr;
	}
	bool ok = true;
	auto sym = [&](const char* name) { void* p = dlsym(h, name); if (!p) ok = false; return p; };
	api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
	api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
	api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
	api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
	api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
	api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
	api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
	api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
	if (!ok)
	{
		error = "libnccl.so.2 lacks a required entry point";
		dlclose(h);
		return nullptr;
	}
	api.handle = h;
	return &api;
}

int NcclAllGatherCount(void* comm, const void* send, void* recv, cudaStream_t stream, std::string& error)
{
	NcclApi* nccl = LoadNccl(error);
	if (!nccl) return TG_ERR_UNSUPPORTED;
	const ncclResult_t r = nccl->AllGather(send, recv, 1, ncclUint64, static_cast<ncclComm_t>(comm), stream);
	if (r != ncclSuccess)
	{
		error = std::string("ncclAllGather: ") + nccl->GetErrorString(r);
		return TG_ERR_CUDA;
	}
	return TG_OK;
}

} // namespace

// ------------------------------------------------------------------------------------------------
// DeviceGroup: worker threads + communicators
// ------------------------------------------------------------------------------------------------

DeviceGroup* DeviceGroup::Create(const std::vector<Context*>& contexts, std::string& error)
{
	NcclApi* nccl = LoadNccl(error);
	if (!nccl) return nullptr;
	std::unique_ptr<DeviceGroup> g(new DeviceGroup());
	g->contexts = contexts;
	const int n = int(contexts.size());
	std::vector<int> devices(size_t(n), 0);
	for (int r = 0; r < n; ++r) devices[size_t(r)] = contexts[size_t(r)]->device;
	std::vector<ncclComm_t> comms(size_t(n), nullptr);
	const ncclResult_t rc = nccl->CommInitAll(comms.data(), n, devices.data());
	if (rc != ncclSuccess)
	{
		error = std::string("ncclCommInitAll: ") + nccl->GetErrorString(rc);
		return nullptr;
	}
	for (ncclComm_t c : comms) g->comms.push_back(c);
	int version = 0;
	nccl->GetVersion(&version);
	g->nccl_version = std::to_string(version / 10000) + "." + std::to_string(version / 100 % 100) + "." + std::to_string(version % 100);
	if (std::getenv("TG_TRACE_NCCL")) std::fprintf(stderr, "tangerine_b200: NCCL %s, ncclCommInitAll over %d devices (comm_nranks %d)\n", g->nccl_version.c_str(), n, n);
	g->status.assign(size_t(n), TG_OK);
	g->errors.assign(size_t(n), std::string());
	for (int r = 0; r < n; ++r) g->threads.emplace_back(&DeviceGroup::Worker, g.get(), r);
	return g.release();
}

DeviceGroup::~DeviceGroup()
{
	{
		std::lock_guard<std::mutex> guard(lock);
		quit.store(true);
		generation.fetch_add(1);
	}
	wake.notify_all();
	for (std::thread& t : threads)
	{
		if (t.joinable()) t.join();
	}
	std::string ignored;
	if (NcclApi* nccl = LoadNccl(ignored))
	{
		for (void* c : comms)
		{
			if (c) nccl->CommDestroy(static_cast<ncclComm_t>(c));
		}
	}
}

// A worker first spins on the generation counter for a while (an export is a few hundred microseconds, and a caller that
// exports repeatedly should not pay a futex wake-up per device each time), then sleeps on the condition variable.
void DeviceGroup::Worker(int rank)
{
	cudaSetDevice(contexts[size_t(rank)]->device);
	uint64_t seen = 0;
	for (;;)
	{
		for (int spin = 0; spin < 40000 && generation.load(std::memory_order_acquire) == seen; ++spin)
		{
#if defined(__x86_64__)
			__builtin_ia32_pause();
#endif
		}
		if (generation.load(std::memory_order_acquire) == seen)
		{
			std::unique_lock<std::mutex> guard(lock);
			wake.wait(guard, [&] { return generation.load(std::memory_order_acquire) != seen; });
		}
		seen = generation.load(std::memory_order_acquire);
		if (quit.load()) return;
		const std::function<int(int, std::string&)>* job = task;
		int rc = TG_ERR_INVALID;
		std::string message;
		try
		{
			rc = (*job)(rank, message);
		}
		catch (const std::bad_alloc&)
		{
			rc = TG_ERR_MEMORY;
			message = "out of host memory";
		}
		catch (const std::exception& e)
		{
			rc = TG_ERR_INVALID;
			message = std::string("internal error: ") + e.what();
		}
		status[size_t(rank)] = rc;
		errors[size_t(rank)] = message;
		if (pending.fetch_sub(1, std::memory_order_acq_rel) == 1)
		{
			std::lock_guard<std::mutex> guard(lock);
			done.notify_all();
		}
	}
}

int DeviceGroup::Run(const std::function<int(int, std::string&)>& fn, std::string& error)
{
	{
		std::lock_guard<std::mutex> guard(lock);
		task = &fn;
		pending.store(size(), std::memory_order_release);
		generation.fetch_add(1, std::memory_order_acq_rel);
	}
	wake.notify_all();
	for (int spin = 0; spin < 2000000 && pending.load(std::memory_order_acquire) != 0; ++spin)
	{
#if defined(__x86_64__)
		__builtin_ia32_pause();
#endif
	}
	if (pending.load(std::memory_order_acquire) != 0)
	{
		std::unique_lock<std::mutex> guard(lock);
		done.wait(guard, [&] { return pending.load(std::memory_order_acquire) == 0; });
	}
	task = nullptr;
	for (int r = 0; r < size(); ++r)
	{
		if (status[size_t(r)] != TG_OK)
		{
			error = "device " + std::to_string(contexts[size_t(r)]->device) + ": " + errors[size_t(r)];
			return status[size_t(r)];
		}
	}
	return TG_OK;
}

// Sense-reversing spin barrier: the waits are microseconds long (the ranks run the same enqueue sequence).
void DeviceGroup::Barrier()
{
	const int phase = barrier_phase.load(std::memory_order_acquire);
	if (barrier_count.fetch_add(1, std::memory_order_acq_rel) + 1 == size())
	{
		barrier_count.store(0, std::memory_order_relaxed);
		barrier_phase.store(phase + 1, std::memory_order_release);
		return;
	}
	int spins = 0;
	while (barrier_phase.load(std::memory_order_acquire) == phase)
	{
		if (++spins > 2000) std::this_thread::yield();
	}
}

// ------------------------------------------------------------------------------------------------
// Replicated models
// ------------------------------------------------------------------------------------------------

Model* Model::CreateReplica(Context* context, Model* primary, std::string& error)
{
	Model* m = new Model(primary->flat_owner);
	m->context = context;
	context->live_results.fetch_add(1);
	m->primary = primary;
	m->leaf_count = primary->leaf_count;
	if (UploadModel(m, error) != TG_OK)
	{
		delete m;
		return nullptr;
	}
	return m;
}

// ------------------------------------------------------------------------------------------------
// Slab planning.  The cuts come from a HOST-side estimate of the work per cell layer, made from the octree's terminus
// cells alone.  Model::Create has marked, for every terminus cell, which of its 4 x 4 x 4 sub-cells can hold surface
// (LeafProfileKernel: the cell's program at the sub-cell centre against the sub-cell's half diagonal -- K0's own test at
// a grid-independent granularity).  A marked sub-cell of side c cells is taken to hold one sheet of surface, c^2 / 64
// active bricks, each costing the FLOPs of the cell's program plus a constant for descent, classification, numbering
// and the per-vertex passes; the weight is spread evenly over the cell layers the sub-cell spans.  No device work at
// export time, no warm-up exports and no feedback: the first export of a model already runs on these cuts.
// ------------------------------------------------------------------------------------------------

std::vector<double> EstimateLayerCost(const FlatModel& flat, const tg_grid& grid)
{
	std::vector<double> cost(size_t(grid.sz), 0.0);
	if (grid.sz == 0) return cost;
	const double gmin[3] = { grid.x, grid.y, grid.z };
	const double step[3] = { grid.dx, grid.dy, grid.dz };
	const double size[3] = { double(grid.sx), double(grid.sy), double(grid.sz) };
	double constant = 1600.0; // (with the straddle weight below: swept at 4 and 8 GPUs on seaside_town 1024^3, profiles/r2_plan_sweep.txt)
	if (const char* env = std::getenv("TG_PLAN_CONSTANT")) constant = std::atof(env);
	const bool have_masks = flat.leaf_mask.size() == flat.leaf_nodes.size();
	for (size_t leaf = 0; leaf < flat.leaf_nodes.size(); ++leaf)
	{
		const FlatNode& node = flat.nodes[flat.leaf_nodes[leaf]];
		const uint64_t mask = have_masks ? flat.leaf_mask[leaf] : ~0ull;
		if (mask == 0ull) continue;
		const double quarter = double(flat.leaf_span[leaf]) * 0.25;
		// A brick that straddles octree cells runs one batch per cell (fewer samples per interpreter dispatch, more box
		// resolution), and small cells mean many primitives: longer programs for K0 and the attribute pass, which the FLOPs of
		// the evaluated program do not show.  (1 + 8 / cells per leaf side) is the mean number of cells a brick meets per
		// axis; the weight is fitted.
		const double leaf_cells = std::max(1.0, double(flat.leaf_span[leaf]) / step[0]);
		double straddle = 1.0 + 3.0 * double(kBrick) / leaf_cells;
		if (const char* env = std::getenv("TG_PLAN_STRADDLE")) straddle = 1.0 + std::atof(env) * double(kBrick) / leaf_cells;
		const double per_brick = (double(node.flops) + constant) * straddle;
		for (int sz = 0; sz < 4; ++sz)
		{
			const uint32_t layer_mask = uint32_t(mask >> (16 * sz)) & 0xFFFFu;
			if (!layer_mask) continue;
			// z range of this sub-cell layer, in cells, clipped to the grid
			const double z0 = (double(node.pivot[2]) + (sz - 2) * quarter - gmin[2]) / step[2], z1 = z0 + quarter / step[2];
			const double k_lo = std::max(0.0, z0), k_hi = std::min(size[2], z1);
			if (!(k_hi > k_lo)) continue;
			double bricks = 0.0;
			for (int s = 0; s < 16; ++s)
			{
				if (!((layer_mask >> s) & 1u)) continue;
				// the sub-cell's footprint in x / y, clipped to the grid
				double side[2];
				bool inside = true;
				for (int a = 0; a < 2; ++a)
				{
					const int index = a == 0 ? (s & 3) : (s >> 2);
					const double lo = (double(node.pivot[a]) + (index - 2) * quarter - gmin[a]) / step[a];
					const double c_lo = std::max(0.0, lo), c_hi = std::min(size[a], lo + quarter / step[a]);
					side[a] = c_hi - c_lo;
					if (!(side[a] > 0.0)) inside = false;
				}
				if (inside) bricks += side[0] * side[1] / double(kBrick * kBrick);
			}
			if (!(bricks > 0.0)) continue;
			const uint32_t k0 = uint32_t(k_lo), k1 = std::min<uint32_t>(uint32_t(grid.sz) - 1, uint32_t(k_hi));
			const double per_layer = bricks * per_brick * std::min(1.0, (k_hi - k_lo) * step[2] / quarter) / double(k1 - k0 + 1);
			for (uint32_t k = k0; k <= k1; ++k) cost[k] += per_layer;
		}
	}
	return cost;
}

// n contiguous slabs of (nearly) equal estimated cost; every slab gets at least one brick row.
std::vector<uint32_t> PlanSlabs(const std::vector<double>& cost, uint32_t sz, int n)
{
	std::vector<uint32_t> cuts(size_t(n) + 1, 0u);
	cuts[size_t(n)] = sz;
	double total = 0.0;
	for (double c : cost) total += c;
	const uint32_t min_layers = kBrick;
	if (!(total > 0.0))
	{
		for (int r = 1; r < n; ++r) cuts[size_t(r)] = uint32_t(uint64_t(sz) * uint64_t(r) / uint64_t(n));
		return cuts;
	}
	double running = 0.0;
	uint32_t k = 0;
	for (int r = 1; r < n; ++r)
	{
		const double target = total * double(r) / double(n);
		while (k < sz && running + cost[k] * 0.5 < target)
		{
			running += cost[k];
			++k;
		}
		uint32_t cut = std::max(k, cuts[size_t(r) - 1] + min_layers);
		cut = std::min(cut, sz - min_layers * uint32_t(n - r));
		while (k < cut)
		{
			running += cost[k];
			++k;
		}
		cuts[size_t(r)] = cut;
	}
	return cuts;
}

// ------------------------------------------------------------------------------------------------
// The export
// ------------------------------------------------------------------------------------------------

namespace
{

struct MultiExport
{
	DeviceGroup* group = nullptr;
	const std::vector<Model*>* models = nullptr;
	tg_grid grid_in;
	DeviceGrid grid;
	tg_mesh_options options;
	std::vector<uint32_t> cuts;
	bool want_host = true;
	// per rank
	std::vector<std::unique_ptr<MeshJob>> jobs;
	std::vector<MeshCounts> counts;
	std::vector<uint32_t> cap_v, cap_q;
	std::vector<tg_mesh_timings> timings;
	std::vector<void*> gathered; // device scratch of the all-gather, one per rank
	std::atomic<int> overflowed{ 0 };
	std::atomic<int> failed{ 0 };
	// one host mesh
	HostArrays host;
	bool host_ok = true;
	std::string host_error;
	uint64_t total_v = 0, total_t = 0;
};

} // namespace

int EngineTimerBeginMulti(DeviceGroup* group, std::string& error)
{
	for (Context* c : group->contexts)
	{
		const int rc = EngineTimerBegin(c, error);
		if (rc != TG_OK) return rc;
	}
	return TG_OK;
}

// Elapsed device time of the slowest rank (every rank's events sit on its own stream).  All end events are recorded
// before the first is waited for: a record-and-wait per device in turn would date the later devices' ends later.
int EngineTimerEndMulti(DeviceGroup* group, float* out_ms, std::string& error)
{
	for (Context* c : group->contexts)
	{
		TG_CUDA(cudaSetDevice(c->device));
		TG_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(c->timer_events[1]), StreamOf(c)));
	}
	float worst = 0.0f;
	for (Context* c : group->contexts)
	{
		TG_CUDA(cudaSetDevice(c->device));
		cudaEvent_t e1 = static_cast<cudaEvent_t>(c->timer_events[1]);
		TG_CUDA(cudaEventSynchronize(e1));
		float ms = 0.0f;
		TG_CUDA(cudaEventElapsedTime(&ms, static_cast<cudaEvent_t>(c->timer_events[0]), e1));
		worst = std::max(worst, ms);
	}
	TG_CUDA(cudaSetDevice(group->contexts[0]->device));
	*out_ms = worst;
	return TG_OK;
}

int EngineExportMeshMulti(DeviceGroup* group, const std::vector<Model*>& models, const tg_grid& grid_in, const tg_mesh_options& options, tg_mesh* out, std::string& error)
{
	std::memset(out, 0, sizeof(*out));
	const int n = group->size();
	Context* primary = group->contexts[0];
	DeviceGrid grid;
	if (!MakeDeviceGrid(grid_in, grid, error)) return TG_ERR_INVALID;
	// Small grids, slab requests and STL face normals (which need the whole mesh on one device) stay on the first GPU.
	if (n == 1 || grid.sz < uint32_t(2 * kBrick * n) || options.slab_begin != 0 || options.slab_end != 0 || (options.flags & TG_MESH_FACE_NORMALS))
	{
		return EngineExportMesh(models[0], grid_in, options, out, error);
	}
	if (primary->Cancelled()) return TG_ERR_CANCELLED;
	primary->stage.store(1);
	primary->progress_done[0] = 0;
	primary->progress_total[0] = uint64_t(n);

	MultiExport ex;
	ex.group = group;
	ex.models = &models;
	ex.grid_in = grid_in;
	ex.grid = grid;
	ex.options = options;
	ex.want_host = !(options.flags & TG_MESH_DEVICE_ONLY);
	{
		auto& plan = models[0]->plan;
		if (plan.ranks != n || std::memcmp(&plan.grid, &grid_in, sizeof(tg_grid)) != 0 || std::getenv("TG_PLAN_CONSTANT"))
		{
			plan.layer_cost = EstimateLayerCost(models[0]->flat, grid_in);
			plan.cuts = PlanSlabs(plan.layer_cost, grid.sz, n);
			plan.grid = grid_in;
			plan.ranks = n;
			plan.feedback_rounds = 0;
		}
		ex.cuts = plan.cuts;
	}
	if (const char* env = std::getenv("TG_SLAB_CUTS")) // diagnostics: "k1,k2,..." overrides the planned cuts
	{
		std::vector<uint32_t> forced(1, 0u);
		for (const char* p = env; *p;)
		{
			forced.push_back(uint32_t(std::strtoul(p, const_cast<char**>(&p), 10)));
			if (*p == ',') ++p;
		}
		forced.push_back(grid.sz);
		if (int(forced.size()) == n + 1) ex.cuts = forced;
	}
	ex.jobs.resize(size_t(n));
	ex.counts.resize(size_t(n));
	ex.cap_v.assign(size_t(n), 0u);
	ex.cap_q.assign(size_t(n), 0u);
	ex.timings.resize(size_t(n));
	ex.gathered.assign(size_t(n), nullptr);
	ex.host.ctx = primary;
	ex.host.want_normals = (options.flags & TG_MESH_NORMALS) != 0;
	ex.host.want_colors = (options.flags & TG_MESH_COLORS) != 0 && models[0]->flat.has_paint;
	const auto h0 = std::chrono::steady_clock::now();

	auto rank_task = [&](int rank, std::string& err) -> int
	{
		Context* ctx = group->contexts[size_t(rank)];
		Model* model = models[size_t(rank)];
		cudaStream_t stream = StreamOf(ctx);
		cudaStream_t copy_stream = static_cast<cudaStream_t>(ctx->copy_stream);
		const bool trace = std::getenv("TG_TRACE_MULTI") != nullptr;
		double t_enqueued = 0.0, t_counts = 0.0, t_finish = 0.0;
		auto since = [&] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - h0).count(); };
		const double t_start = since();
#define TG_RANK_CUDA(call)                                                                  \
		do                                                                                  \
		{                                                                                   \
			cudaError_t e_ = (call);                                                        \
			if (e_ != cudaSuccess && rc == TG_OK)                                           \
			{                                                                               \
				err = std::string(#call) + ": " + cudaGetErrorString(e_);                  \
				rc = e_ == cudaErrorMemoryAllocation ? TG_ERR_MEMORY : TG_ERR_CUDA;         \
			}                                                                               \
		} while (0)
		int rc = TG_OK;
		TG_RANK_CUDA(cudaSetDevice(ctx->device));
		ctx->progress_base = 0;
		ctx->progress_slabs = 1;
		if (ctx->progress_words) ctx->progress_words[0] = ctx->progress_words[1] = 0u;
		ctx->progress_seen[0] = ctx->progress_seen[1] = 0u;
		if (!ctx->index_base) TG_RANK_CUDA(cudaMalloc(&ctx->index_base, 8));
		if (!ex.gathered[size_t(rank)])
		{
			void* g = ctx->AcquireDevice(size_t(n) * 8 + 256, err);
			if (!g && rc == TG_OK) rc = TG_ERR_MEMORY;
			ex.gathered[size_t(rank)] = g;
		}
		tg_mesh_options slab = ex.options;
		slab.flags |= TG_MESH_DEVICE_ONLY;
		slab.slab_begin = ex.cuts[size_t(rank)];
		slab.slab_end = ex.cuts[size_t(rank) + 1];
		MultiHook hook;
		hook.comm = group->comms[size_t(rank)];
		hook.rank = rank;
		hook.ranks = n;
		hook.gathered = static_cast<unsigned long long*>(ex.gathered[size_t(rank)]);
		hook.all_gather = &NcclAllGatherCount;
		// Every rank makes the same sequence of collective calls whatever happens to it: a rank that failed before the
		// all-gather still has to enter it, or its peers would wait forever.  So errors are remembered, not returned early.
		if (ex.cap_v[size_t(rank)] == 0 && (ex.options.flags & TG_MESH_REBALANCE) && model->last_slab_vertices > 0)
		{
			// the cuts move from export to export: size the arrays by this device's last slab (plus half) instead of by
			// the cautious default for an unknown slab, so that the result buffers of the last export fit again
			ex.cap_v[size_t(rank)] = uint32_t(std::min<uint64_t>(model->last_slab_vertices + model->last_slab_vertices / 2 + 4096, 0xFFFFFFF0ull));
			ex.cap_q[size_t(rank)] = uint32_t(std::min<uint64_t>(model->last_slab_quads + model->last_slab_quads / 2 + 4096, 0x2AAAAAA0ull));
		}
		for (int attempt = 0; attempt < 2; ++attempt)
		{
			ex.jobs[size_t(rank)].reset(new MeshJob());
			MeshJob& job = *ex.jobs[size_t(rank)];
			bool entered = false;
			if (rc == TG_OK)
			{
				rc = EnqueueMesh(job, model, ex.grid_in, slab, ex.cap_v[size_t(rank)], ex.cap_q[size_t(rank)], static_cast<unsigned long long*>(ctx->index_base), err, 0, nullptr, nullptr, &hook);
				entered = job.bricks_done != nullptr; // the collective was reached
			}
			if (!entered && hook.gathered)
			{
				std::string ignored;
				TG_RANK_CUDA(cudaMemsetAsync(ctx->index_base, 0, 8, stream));
				NcclAllGatherCount(hook.comm, ctx->index_base, hook.gathered, stream, ignored);
			}
			MeshCounts& counts = ex.counts[size_t(rank)];
			counts = MeshCounts();
			t_enqueued = since();
			if (rc == TG_OK) rc = WaitCounts(job, counts, err);
			t_counts = since();
			if (rc != TG_OK) ex.failed.store(1);
			if (rc == TG_OK && counts.overflow) ex.overflowed.store(1);
			group->Barrier();
			if (ex.failed.load() || !ex.overflowed.load() || attempt == 1) break;
			// some slab outgrew its arrays: every rank repeats with exact sizes (the counts are exact even then)
			cudaStreamSynchronize(stream);
			cudaStreamSynchronize(static_cast<cudaStream_t>(ctx->stream2));
			FreeResultDevice(job.result, stream);
			delete job.result;
			job.result = nullptr;
			ex.cap_v[size_t(rank)] = uint32_t(std::max<uint64_t>(counts.vertices, 1));
			ex.cap_q[size_t(rank)] = uint32_t(std::max<uint64_t>(counts.quads, 1));
			group->Barrier();
			if (rank == 0) ex.overflowed.store(0);
			group->Barrier();
		}
		MeshJob& job = *ex.jobs[size_t(rank)];
		if (ex.failed.load() || ex.overflowed.load())
		{
			cudaStreamSynchronize(stream);
			cudaStreamSynchronize(static_cast<cudaStream_t>(ctx->stream2));
			if (rc == TG_OK && ex.overflowed.load())
			{
				err = "mesh capacities overflowed twice";
				rc = TG_ERR_CUDA;
			}
			return rc;
		}
		// ---- the stitch: offsets of this slab in the one host mesh -----------------------------------
		uint64_t v_before = 0, t_before = 0, v_all = 0, t_all = 0;
		for (int r = 0; r < n; ++r)
		{
			if (r < rank)
			{
				v_before += ex.counts[size_t(r)].vertices;
				t_before += ex.counts[size_t(r)].quads * 2;
			}
			v_all += ex.counts[size_t(r)].vertices;
			t_all += ex.counts[size_t(r)].quads * 2;
		}
		if (rank == 0)
		{
			ex.total_v = v_all;
			ex.total_t = t_all;
			if (v_all > 0xFFFFFFF0ull || t_all > 0xFFFFFFF0ull)
			{
				ex.host_ok = false;
				ex.host_error = "mesh exceeds 2^32 vertices or triangles";
			}
			else if (ex.want_host)
			{
				// exact sizes, blocks of the primary context's pinned cache (a repeated export allocates nothing)
				HostArrays& h = ex.host;
				h.positions = static_cast<float*>(primary->AcquirePinned(size_t(std::max<uint64_t>(v_all, 1)) * 12, ex.host_error));
				if (h.want_normals) h.normals = static_cast<float*>(primary->AcquirePinned(size_t(std::max<uint64_t>(v_all, 1)) * 12, ex.host_error));
				if (h.want_colors) h.colors = static_cast<uint8_t*>(primary->AcquirePinned(size_t(std::max<uint64_t>(v_all, 1)) * 3, ex.host_error));
				h.triangles = static_cast<uint32_t*>(primary->AcquirePinned(size_t(std::max<uint64_t>(t_all, 1)) * 12, ex.host_error));
				if (!h.positions || !h.triangles || (h.want_normals && !h.normals) || (h.want_colors && !h.colors)) ex.host_ok = false;
			}
		}
		group->Barrier();
		const MeshCounts& counts = ex.counts[size_t(rank)];
		const uint64_t v = counts.vertices, t = counts.quads * 2;
		model->last_slab_vertices = counts.vertices;
		model->last_slab_quads = counts.quads;
		MeshResultDevice* r = job.result;
		if (ex.want_host && ex.host_ok)
		{
			if (t > 0)
			{
				TG_RANK_CUDA(cudaStreamWaitEvent(copy_stream, job.faces_ready, 0));
				TG_RANK_CUDA(cudaMemcpyAsync(ex.host.triangles + t_before * 3, r->d_triangles, size_t(t) * 12, cudaMemcpyDeviceToHost, copy_stream));
			}
		}
		tg_mesh part;
		std::memset(&part, 0, sizeof(part));
		if (rc == TG_OK) rc = FinishJob(job, counts, &part, err);
		t_finish = since();
		ex.timings[size_t(rank)] = part.timings;
		std::free(part.layer_vertices);
		std::free(part.layer_vertex_cost);
		if (ex.want_host && ex.host_ok && v > 0)
		{
			TG_RANK_CUDA(cudaMemcpyAsync(ex.host.positions + v_before * 3, r->d_positions, size_t(v) * 12, cudaMemcpyDeviceToHost, stream));
			if (ex.host.normals && r->d_normals) TG_RANK_CUDA(cudaMemcpyAsync(ex.host.normals + v_before * 3, r->d_normals, size_t(v) * 12, cudaMemcpyDeviceToHost, stream));
			if (ex.host.colors && r->d_colors) TG_RANK_CUDA(cudaMemcpyAsync(ex.host.colors + v_before * 3, r->d_colors, size_t(v) * 3, cudaMemcpyDeviceToHost, stream));
		}
		TG_RANK_CUDA(cudaStreamSynchronize(copy_stream));
		TG_RANK_CUDA(cudaStreamSynchronize(stream));
		TG_RANK_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(ctx->stream2)));
		if (ex.want_host)
		{
			FreeResultDevice(r, stream);
			delete r;
			job.result = nullptr;
		}
		primary->progress_done[0].fetch_add(1);
		if (trace) std::fprintf(stderr, "rank %d host us: start %.0f  enqueued %.0f  counts %.0f  finished %.0f  done %.0f  (device total %.0f)\n", rank, t_start, t_enqueued, t_counts, t_finish, since(), double(part.timings.total_device_ms) * 1e3);
#undef TG_RANK_CUDA
		return rc;
	};

	int rc = group->Run(rank_task, error);
	if (std::getenv("TG_TRACE_MULTI")) std::fprintf(stderr, "multi export host us: all ranks done %.0f\n", std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - h0).count());
	for (int r = 0; r < n; ++r)
	{
		Context* ctx = group->contexts[size_t(r)];
		ctx->ReleaseDevice(ex.gathered[size_t(r)]);
	}
	if (rc == TG_OK && !ex.host_ok)
	{
		error = ex.host_error;
		rc = ex.host_error.find("2^32") != std::string::npos ? TG_ERR_UNSUPPORTED : TG_ERR_MEMORY;
	}
	primary->stage.store(0);
	if (rc != TG_OK)
	{
		for (auto& j : ex.jobs)
		{
			if (j && j->result)
			{
				cudaSetDevice(j->result->context->device);
				cudaStreamSynchronize(StreamOf(j->result->context));
				FreeResultDevice(j->result, nullptr);
				delete j->result;
				j->result = nullptr;
			}
		}
		ex.host.Release();
		cudaSetDevice(primary->device);
		return rc;
	}
	cudaSetDevice(primary->device);

	// ---- one result -----------------------------------------------------------------------------------
	MeshResultDevice* result = new MeshResultDevice(primary);
	if (ex.want_host)
	{
		if (ex.host.positions) result->pinned.push_back(ex.host.positions);
		if (ex.host.normals) result->pinned.push_back(ex.host.normals);
		if (ex.host.colors) result->pinned.push_back(ex.host.colors);
		if (ex.host.triangles) result->pinned.push_back(ex.host.triangles);
		out->positions = ex.total_v ? ex.host.positions : nullptr;
		out->normals = ex.total_v ? ex.host.normals : nullptr;
		out->colors = ex.total_v ? ex.host.colors : nullptr;
		out->triangles = ex.total_t ? ex.host.triangles : nullptr;
	}
	else
	{
		for (auto& j : ex.jobs) // the per-device arrays stay in HBM until tg_mesh_free
		{
			result->parts.push_back(j->result);
			j->result = nullptr;
		}
	}
	out->opaque = result;
	out->vertex_count = ex.total_v;
	out->triangle_count = ex.total_t;
	tg_mesh_timings& tm = out->timings;
	for (int r = 0; r < n; ++r)
	{
		const tg_mesh_timings& p = ex.timings[size_t(r)];
		tm.cull_ms = std::max(tm.cull_ms, p.cull_ms);
		tm.evaluate_ms = std::max(tm.evaluate_ms, p.evaluate_ms);
		tm.compact_ms = std::max(tm.compact_ms, p.compact_ms);
		tm.faces_ms = std::max(tm.faces_ms, p.faces_ms);
		tm.attributes_ms = std::max(tm.attributes_ms, p.attributes_ms);
		tm.total_device_ms = std::max(tm.total_device_ms, p.total_device_ms);
		tm.bricks_total += p.bricks_total;
		tm.bricks_evaluated += p.bricks_evaluated;
		tm.samples_evaluated += p.samples_evaluated;
		tm.algorithmic_flops += p.algorithmic_flops;
		tm.kernel_launches += p.kernel_launches;
	}
	tm.download_ms = ex.want_host ? float(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count()) : 0.0f;
	if ((options.flags & TG_MESH_REBALANCE) && models[0]->plan.layer_cost.size() == grid.sz)
	{
		// Opt-in feedback: the estimate of every slab's layers is scaled towards the share of the work its device just
		// measured (culling + evaluation + numbering + attributes; the wait for the all-gather is not work), and the cuts
		// of the next export of this model and grid are planned on the corrected estimate.
		auto& plan = models[0]->plan;
		std::vector<double> measured(size_t(n), 0.0), predicted(size_t(n), 0.0);
		double measured_sum = 0.0, predicted_sum = 0.0;
		for (int r = 0; r < n; ++r)
		{
			const tg_mesh_timings& p = ex.timings[size_t(r)];
			measured[size_t(r)] = double(p.cull_ms) + p.evaluate_ms + p.compact_ms + p.attributes_ms;
			for (uint32_t k = ex.cuts[size_t(r)]; k < ex.cuts[size_t(r) + 1]; ++k) predicted[size_t(r)] += plan.layer_cost[k];
			measured_sum += measured[size_t(r)];
			predicted_sum += predicted[size_t(r)];
		}
		if (measured_sum > 0.0 && predicted_sum > 0.0)
		{
			for (int r = 0; r < n; ++r)
			{
				if (!(predicted[size_t(r)] > 0.0) || !(measured[size_t(r)] > 0.0)) continue;
				// (square root: half the correction per export, so that one noisy measurement cannot throw the cuts)
				const double scale = std::sqrt((measured[size_t(r)] / measured_sum) / (predicted[size_t(r)] / predicted_sum));
				for (uint32_t k = ex.cuts[size_t(r)]; k < ex.cuts[size_t(r) + 1]; ++k) plan.layer_cost[k] *= scale;
			}
			plan.cuts = PlanSlabs(plan.layer_cost, grid.sz, n);
			plan.feedback_rounds++;
		}
	}
	// per-rank detail for tg_mesh_rank_timings (bench.py reports the balance)
	result->rank_timings = ex.timings;
	result->rank_cuts = ex.cuts;
	if (primary->Cancelled())
	{
		EngineFreeMesh(out);
		return TG_ERR_CANCELLED;
	}
	return TG_OK;
}
