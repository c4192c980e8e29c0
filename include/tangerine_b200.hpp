// C++ host interface of the B200 export path: the same names, argument meaning and threading behaviour
// as the reference's tangerine/export.h:19-39 and tangerine/magica.h:20, implemented entirely on top of
// the C ABI in tangerine_b200.h.  A maintainer swapping the CPU path for this one changes the callee,
// not the call sites.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "tangerine_b200.h"

namespace tangerine_b200
{

enum class ExportFormat // tangerine/export.h:19-25
{
	STL,
	PLY,
	VOX,
	Unknown
};

struct ExportProgress // tangerine/export.h:27-34
{
	int Stage;
	float Generation;
	float Refinement;
	float Secondary;
	float Write;
};

// CUDA device used by the calls below (default 0).
TG_API void SetExportDevice(int CudaDevice);

// tangerine/export.h:36 / export.cpp:566-579.  Starts the export on a detached thread and returns at once;
// poll GetExportProgress() (Stage returns to 0 when the file is written).  ExportPointCloud selects
// PointCloudExportThread's behaviour (PLY only).  Unlike the reference's mesh path in this snapshot,
// RefineIterations is honoured for meshes too (BASELINE.json north_star).
TG_API void MeshExport(const tg_tree* Evaluator, std::string Path, const float ModelMin[3], const float ModelMax[3], const float Step[3],
	int RefineIterations, ExportFormat Format, bool ExportPointCloud, float Scale = 1.0f);

// tangerine/export.h:38-39 / export.cpp:483-492, 582-592
TG_API void CancelExport(bool Halt);
TG_API ExportProgress GetExportProgress();
// Status of the most recent MeshExport once Stage is back to 0: TG_OK or an error code; message via LastExportError().
TG_API int LastExportStatus();
TG_API std::string LastExportError();

// export.cpp:595-607: synchronous export over the evaluator's own bounds with Step = 1 / GridSize.
TG_API int ExportCommon(const tg_tree* Evaluator, float GridSize, int RefineIterations, const char* Path, ExportFormat Format, float Scale = 1.0f);

// tangerine/magica.h:20 / magica.cpp:27-72: synchronous MagicaVoxel export.
TG_API int VoxExport(const tg_tree* Evaluator, const std::string& Path, float GridSize, int ColorIndex);

// The arrays the live mesher fills on a Drawable (tangerine/sdf_model.h:59-62), four floats per vertex as there:
// Positions (x, y, z, 1), Normals (SDFOctree::Gradient, 1) and Colors; Indices three per triangle.
struct LiveDrawable
{
	std::vector<float> Positions;
	std::vector<float> Normals;
	std::vector<float> Colors;
	std::vector<uint32_t> Indices;
};

// Sodapop::Populate (tangerine/sodapop.cpp:214-225) with the NaiveSurfaceNets algorithm (:562-897), synchronous: the
// octree of MeshingJob::Run (:240), the grid of NaiveSurfaceNetsScratch at density 20 + MeshingDensityPush (:43, 153-179,
// 221), the clamped inexact field (:583-587), vertex and face loops, gradient normals (:816).  Vertices come in (k, j, i)
// cell order (the reference's order depends on its thread schedule).  Colors are (0, 0, 0, 1) as the reference leaves
// them for materials without a Chthonic evaluator (:862-870); evaluating those materials is not part of this path.
TG_API int PopulateDrawable(const tg_tree* Evaluator, float MeshingDensityPush, LiveDrawable& Painter);

} // namespace tangerine_b200
