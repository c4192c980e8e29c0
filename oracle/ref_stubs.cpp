// TEST INFRASTRUCTURE ONLY (oracle/_ref build).
//
// The reference's Lua front-end and export code reference ~18 symbols that live in its UI / GL
// translation units (tangerine.cpp, sdf_model.cpp, painting_set.cpp, gl_boilerplate.cpp, lights.cpp).
// Those units need SDL2 + OpenGL and are not on the hot path, so the oracle build replaces them with
// the no-op definitions below.  The only behaviour kept is "remember the tree the
// script produced" so that ref_tool.cpp can hand it to the reference's own export entry points.

#include <functional>
#include <string>
#include <cstdio>

#include "sdf_model.h"
#include "painting_set.h"
#include "lights.h"
#include "gl_boilerplate.h"

SDFNodeShared g_CapturedTree = nullptr;
std::string g_ScriptError;

void FlagSceneRepaint() {}
void PostPendingRepaintRequest() {}
void ShowDebugMenu() {}
void HideDebugMenu() {}
void SetWindowTitle(std::string) {}
void SetClearColor(glm::vec3&) {}
void SetFixedCamera(glm::vec3&, glm::vec3&, glm::vec3&) {}
void ClearTreeEvaluator() { g_CapturedTree.reset(); }
void SetTreeEvaluator(SDFNodeShared& InTreeEvaluator) { g_CapturedTree = InTreeEvaluator; }

void LoadModelCommon(std::function<void()> LoadingCallback)
{
	LoadingCallback();
}

void PostScriptError(std::string ErrorMessage)
{
	g_ScriptError = ErrorMessage;
	std::fprintf(stderr, "script error: %s\n", ErrorMessage.c_str());
}

// --- PaintingSet: the front-end only needs a non-null handle to pass around.
std::shared_ptr<PaintingSet> PaintingSet::Create() { return nullptr; }
void PaintingSet::GlobalApply(std::function<void(SDFModelShared)>&) {}
SDFModelShared PaintingSet::GlobalSelect(std::function<bool(SDFModelShared)>&) { return nullptr; }

// --- SDFModel: capture the evaluator instead of building a drawable.
SDFModel::SDFModel(SDFNodeShared& InEvaluator, const std::string& InName, const float, const float, const VertexSequence)
{
	Evaluator = InEvaluator;
	Name = InName;
}
SDFModel::~SDFModel() {}
void SDFModel::RegisterNewModel(std::shared_ptr<PaintingSet>&, SDFModelShared&) {}
SDFModelShared SDFModel::Create(std::shared_ptr<PaintingSet>&, SDFNodeShared& InEvaluator, const std::string&,
	const float, const float, const VertexSequence)
{
	// The real implementation registers a drawable; the oracle only needs the tree.
	g_CapturedTree = InEvaluator;
	return nullptr;
}

// --- GL buffer wrapper: never touches GL here.
Buffer::Buffer(Buffer&& Old) : BufferID(0), DebugName(Old.DebugName), LastSize(0) {}
Buffer::Buffer(const char* InDebugName) : BufferID(0), DebugName(InDebugName), LastSize(0) {}
Buffer::~Buffer() {}
void Buffer::Release() {}

LightShared PointLight::Create(glm::vec3) { return nullptr; }
