// Host mirror of the reference's export interface (include/tangerine_b200.hpp), written against the C ABI only.
#include "../../include/tangerine_b200.hpp"

#include <atomic>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <thread>

namespace tangerine_b200
{

namespace
{
std::atomic<int> g_device{ 0 };
std::mutex g_state_lock;
tg_context* g_active_context = nullptr; // context of the export in flight, for progress / cancel
std::atomic<int> g_stage{ 0 };
std::atomic<bool> g_cancel{ false }; // CancelExport seen by the export in flight, whatever phase it is in
std::atomic<int> g_status{ TG_OK };
std::string g_error;

// One context per device for the life of the process: a context owns the scratch arena, the result caches and the
// page-locked staging memory, and making them costs more than an export (first export on a fresh context 220 ms, on a
// kept one under 30 ms for seaside_town 1024^3).  `busy` serialises the callers (a context is used by one thread at a time).
struct KeptContext
{
	tg_context* context = nullptr;
	std::mutex busy;
};

KeptContext* ContextOf(int device)
{
	static std::mutex lock;
	static std::map<int, std::unique_ptr<KeptContext>> kept;
	std::lock_guard<std::mutex> guard(lock);
	std::unique_ptr<KeptContext>& slot = kept[device];
	if (!slot) slot.reset(new KeptContext());
	if (!slot->context) slot->context = tg_context_create(device);
	return slot->context ? slot.get() : nullptr;
}

void Finish(int status)
{
	std::lock_guard<std::mutex> lock(g_state_lock);
	g_status.store(status);
	g_error = status == TG_OK ? std::string() : std::string(tg_last_error());
}

int RunExport(const tg_tree* tree, const std::string& path, const float mn[3], const float mx[3], const float step[3], int refine,
	ExportFormat format, bool point_cloud, float scale)
{
	if (format != ExportFormat::PLY && format != ExportFormat::STL) return TG_ERR_INVALID;
	KeptContext* kept = ContextOf(g_device.load());
	if (!kept) return TG_ERR_NO_DEVICE;
	std::lock_guard<std::mutex> busy(kept->busy);
	tg_context* context = kept->context;
	{
		std::lock_guard<std::mutex> lock(g_state_lock);
		tg_rearm(context); // a cancel of the previous export must not outlive it
		g_active_context = context;
		if (g_cancel.load()) tg_cancel(context, 1); // CancelExport ran before this export had its context
	}
	int rc = TG_ERR_INVALID;
	tg_model* model = tg_model_create(context, tree, 0.25f, 0); // SDFOctree::Create(Evaluator, 0.25), export.cpp:322 / 388
	if (model && g_cancel.load())
	{
		rc = TG_ERR_CANCELLED; // cancelled while the octree was being built
		tg_model_destroy(model);
		model = nullptr;
	}
	if (model)
	{
		tg_mesh mesh;
		std::memset(&mesh, 0, sizeof(mesh));
		if (point_cloud)
		{
			// PointCloudExportThread only writes PLY (export.cpp:473-477)
			rc = format == ExportFormat::PLY ? tg_export_points(model, mn, mx, step, refine, TG_MESH_NORMALS | TG_MESH_COLORS | TG_MESH_KEEP_CANCEL, scale, &mesh) : TG_ERR_INVALID;
		}
		else
		{
			tg_grid grid;
			rc = tg_export_grid(mn, mx, step, &grid);
			if (rc == TG_OK)
			{
				tg_mesh_options options;
				std::memset(&options, 0, sizeof(options));
				options.flags = (format == ExportFormat::STL ? TG_MESH_FACE_NORMALS : (TG_MESH_NORMALS | TG_MESH_COLORS)) | TG_MESH_KEEP_CANCEL;
				options.refine_iterations = refine;
				options.scale = scale;
				rc = tg_export_mesh(model, &grid, &options, &mesh);
			}
		}
		if (rc == TG_OK && g_cancel.load())
		{
			rc = TG_ERR_CANCELLED; // no file is written for a cancelled export
			tg_mesh_free(&mesh);
		}
		else if (rc == TG_OK)
		{
			g_stage.store(3);
			rc = format == ExportFormat::STL ? tg_write_stl(path.c_str(), &mesh) : tg_write_ply(path.c_str(), &mesh);
			tg_mesh_free(&mesh);
		}
		tg_model_destroy(model);
	}
	{
		std::lock_guard<std::mutex> lock(g_state_lock);
		g_active_context = nullptr;
	}
	return rc;
}
} // namespace

void SetExportDevice(int CudaDevice)
{
	g_device.store(CudaDevice);
}

void MeshExport(const tg_tree* Evaluator, std::string Path, const float ModelMin[3], const float ModelMax[3], const float Step[3],
	int RefineIterations, ExportFormat Format, bool ExportPointCloud, float Scale)
{
	g_status.store(TG_OK);
	g_cancel.store(false); // MeshExport re-arms ExportActive (export.cpp:568)
	g_stage.store(1);
	tg_tree* copy = tg_tree_copy(Evaluator); // the detached thread must not depend on the caller's lifetime
	const float mn[3] = { ModelMin[0], ModelMin[1], ModelMin[2] };
	const float mx[3] = { ModelMax[0], ModelMax[1], ModelMax[2] };
	const float st[3] = { Step[0], Step[1], Step[2] };
	std::thread worker([=]()
	{
		int rc = TG_ERR_INVALID;
		try
		{
			if (copy) rc = RunExport(copy, Path, mn, mx, st, RefineIterations, Format, ExportPointCloud, Scale);
		}
		catch (...)
		{
			rc = TG_ERR_MEMORY; // nothing may unwind out of a detached thread
		}
		tg_tree_free(copy);
		Finish(rc);
		g_stage.store(0);
	});
	worker.detach();
}

void CancelExport(bool Halt)
{
	g_cancel.store(true);
	std::lock_guard<std::mutex> lock(g_state_lock);
	if (g_active_context) tg_cancel(g_active_context, Halt ? 1 : 0);
}

ExportProgress GetExportProgress()
{
	ExportProgress progress = { g_stage.load(), 0.0f, 0.0f, 0.0f, 0.0f };
	std::lock_guard<std::mutex> lock(g_state_lock);
	if (g_active_context)
	{
		float ratios[4] = { 0, 0, 0, 0 };
		int stage = 0;
		tg_progress(g_active_context, ratios, &stage);
		if (stage != 0) progress.Stage = stage;
		progress.Generation = ratios[0];
		progress.Refinement = ratios[1];
		progress.Secondary = ratios[2];
		progress.Write = ratios[3];
	}
	return progress;
}

int LastExportStatus()
{
	return g_status.load();
}

std::string LastExportError()
{
	std::lock_guard<std::mutex> lock(g_state_lock);
	return g_error;
}

int ExportCommon(const tg_tree* Evaluator, float GridSize, int RefineIterations, const char* Path, ExportFormat Format, float Scale)
{
	float mn[3], mx[3];
	int rc = tg_tree_bounds(Evaluator, mn, mx);
	if (rc != TG_OK) return rc;
	if (!Path || !(GridSize > 0.0f)) return TG_ERR_INVALID;
	const float step = float(1.0 / GridSize);
	const float steps[3] = { step, step, step };
	g_cancel.store(false);
	g_stage.store(1);
	rc = RunExport(Evaluator, Path, mn, mx, steps, RefineIterations, Format, false, Scale);
	Finish(rc);
	g_stage.store(0);
	return rc;
}

int PopulateDrawable(const tg_tree* Evaluator, float MeshingDensityPush, LiveDrawable& Painter)
{
	Painter = LiveDrawable();
	KeptContext* kept = ContextOf(g_device.load());
	if (!kept) return TG_ERR_NO_DEVICE;
	std::lock_guard<std::mutex> busy(kept->busy);
	tg_context* context = kept->context;
	tg_rearm(context);
	int rc = TG_ERR_INVALID;
	tg_model* model = tg_model_create_live(context, Evaluator, 0.25f, 0); // SDFOctree::Create(Evaluator, .25, false, 3, Margin = 0), sodapop.cpp:240
	if (model)
	{
		tg_grid grid;
		rc = tg_live_grid(model, 20.0f + MeshingDensityPush, &grid); // DefaultMeshingDensity + MeshingDensityPush, sodapop.cpp:43, 221
		tg_mesh mesh;
		std::memset(&mesh, 0, sizeof(mesh));
		if (rc == TG_OK)
		{
			tg_mesh_options options;
			std::memset(&options, 0, sizeof(options));
			options.flags = TG_MESH_NORMALS | TG_MESH_LIVE_FIELD;
			rc = tg_export_mesh(model, &grid, &options, &mesh);
		}
		if (rc == TG_OK)
		{
			try
			{
				const size_t vertices = size_t(mesh.vertex_count);
				Painter.Positions.resize(vertices * 4);
				Painter.Normals.resize(vertices * 4);
				Painter.Colors.assign(vertices * 4, 0.0f);
				for (size_t v = 0; v < vertices; ++v)
				{
					for (int a = 0; a < 3; ++a)
					{
						Painter.Positions[v * 4 + a] = mesh.positions[v * 3 + a];
						Painter.Normals[v * 4 + a] = mesh.normals[v * 3 + a];
					}
					Painter.Positions[v * 4 + 3] = 1.0f;
					Painter.Normals[v * 4 + 3] = 1.0f;
					Painter.Colors[v * 4 + 3] = 1.0f;
				}
				Painter.Indices.assign(mesh.triangles, mesh.triangles + size_t(mesh.triangle_count) * 3);
			}
			catch (...)
			{
				rc = TG_ERR_MEMORY;
			}
		}
		tg_mesh_free(&mesh);
		tg_model_destroy(model);
	}
	return rc;
}

int VoxExport(const tg_tree* Evaluator, const std::string& Path, float GridSize, int ColorIndex)
{
	return tg_export_magica_voxel(Evaluator, GridSize, ColorIndex, Path.c_str(), g_device.load());
}

} // namespace tangerine_b200
