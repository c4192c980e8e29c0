// TEST INFRASTRUCTURE ONLY (oracle/_ref build).
//
// Command-line driver around the reference's OWN hot-path code, compiled unmodified from
// $(REF)/tangerine and $(REF)/third_party (see oracle/Makefile).  It exists to
//   * turn the reference's Lua models into portable tree files (.tgm) for the tests,
//   * produce golden vectors (raw SDF samples, gradients, material colours, octree programs,
//     exported PLY / STL / VOX files) from the reference implementation itself,
//   * time the reference CPU export path (bench.py --impl reference, cpu_baseline kind "reference").
//
// The first include pulls in the reference's sdf_evaluator.cpp *textually, unmodified, from where it
// lies*: the node classes (BrushNode, SetNode, FlateNode, StencilMaskNode) are private to that
// translation unit, and the tree dumper / loader below needs to see them.  Nothing from the
// reference is copied into this repository.

#include REF_SDF_EVALUATOR_CPP

#include <cstdio>
#include <cstring>
#include <cstdint>
#include <chrono>
#include <map>
#include <set>
#include <thread>
#include <atomic>
#include <fstream>
#include <surface_nets.h>
#include <numeric>
#include <algorithm>

#include "lua_env.h"
#include "export.h"
#include "magica.h"
#include "mesh_generators.h"

extern SDFNodeShared g_CapturedTree;
extern std::atomic_bool ExportActive;
extern std::atomic_int ExportState;
extern std::atomic_int GenerationProgress;
extern std::atomic_int RefinementProgress;
extern std::atomic_int SecondaryProgress;
extern std::atomic_int WriteProgress;
void ExportCommon(SDFNodeShared Evaluator, float GridSize, int RefineIterations, const char* Path, ExportFormat Format, float Scale);
void MeshExportThread(SDFNodeShared Evaluator, vec3 ModelMin, vec3 ModelMax, vec3 Step, int RefineIterations, std::string Path, ExportFormat Format, float Scale);
void PointCloudExportThread(SDFNodeShared Evaluator, vec3 ModelMin, vec3 ModelMax, vec3 Step, int RefineIterations, std::string Path, ExportFormat Format, float Scale);

using Clock = std::chrono::steady_clock;
static double Seconds(Clock::time_point A, Clock::time_point B)
{
	return std::chrono::duration<double>(B - A).count();
}

// ------------------------------------------------------------------------------------------------
// .tgm tree files (format documented in tangerine_b200/host/tgm.h)
// ------------------------------------------------------------------------------------------------

static const uint32_t TGM_NONE = 0xFFFFFFFFu;
static const uint32_t TGM_STENCIL_POS = 100; // StencilMaskNode<false>
static const uint32_t TGM_STENCIL_NEG = 101; // StencilMaskNode<true>

struct TgmNode
{
	uint32_t Kind;
	uint32_t A;
	uint32_t B;
	uint32_t Material;
	float Params[4];
	float Quat[4]; // w x y z
	float Trans[3];
	float Scale;
	float BoundsMin[3];
	float BoundsMax[3];
};
static_assert(sizeof(TgmNode) == 88);

struct TgmWriter
{
	std::vector<TgmNode> Nodes;
	std::vector<vec3> MaterialColors;
	std::map<MaterialInterface*, uint32_t> MaterialIds;

	uint32_t MaterialId(MaterialShared& Material)
	{
		if (!Material)
		{
			return TGM_NONE;
		}
		auto Found = MaterialIds.find(Material.get());
		if (Found != MaterialIds.end())
		{
			return Found->second;
		}
		// The export path only ever looks at SampleColor(GuessColor()) (export.cpp:303-311).
		vec3 Color = SampleColor(Material->GuessColor());
		uint32_t Id = uint32_t(MaterialColors.size());
		MaterialColors.push_back(Color);
		MaterialIds[Material.get()] = Id;
		return Id;
	}

	template<typename BrushT>
	bool TryBrush(SDFNode* Node, uint32_t& OutIndex)
	{
		BrushT* Brush = dynamic_cast<BrushT*>(Node);
		if (!Brush)
		{
			return false;
		}
		TgmNode Rec = {};
		Rec.Kind = uint32_t(Brush->Opcode);
		Rec.A = TGM_NONE;
		Rec.B = TGM_NONE;
		Rec.Material = MaterialId(Brush->Material);
		for (size_t i = 0; i < Brush->NodeParams.size(); ++i)
		{
			Rec.Params[i] = Brush->NodeParams[i];
		}
		Rec.Quat[0] = Brush->LocalToWorld.Rotation.w;
		Rec.Quat[1] = Brush->LocalToWorld.Rotation.x;
		Rec.Quat[2] = Brush->LocalToWorld.Rotation.y;
		Rec.Quat[3] = Brush->LocalToWorld.Rotation.z;
		for (int i = 0; i < 3; ++i)
		{
			Rec.Trans[i] = Brush->LocalToWorld.Translation[i];
			Rec.BoundsMin[i] = Brush->BrushAABB.Min[i];
			Rec.BoundsMax[i] = Brush->BrushAABB.Max[i];
		}
		Rec.Scale = Brush->LocalToWorld.Scalation;
		OutIndex = uint32_t(Nodes.size());
		Nodes.push_back(Rec);
		return true;
	}

	template<SetFamily Family, bool Blend>
	bool TrySet(SDFNode* Node, uint32_t& OutIndex)
	{
		auto* Set = dynamic_cast<SetNode<Family, Blend>*>(Node);
		if (!Set)
		{
			return false;
		}
		TgmNode Rec = {};
		Rec.Kind = uint32_t(Set->Opcode);
		Rec.A = Visit(Set->LHS.get());
		Rec.B = Visit(Set->RHS.get());
		Rec.Material = TGM_NONE;
		Rec.Params[0] = Set->Threshold;
		Rec.Quat[0] = 1.0f;
		Rec.Scale = 1.0f;
		OutIndex = uint32_t(Nodes.size());
		Nodes.push_back(Rec);
		return true;
	}

	template<bool ApplyToNegative>
	bool TryStencil(SDFNode* Node, uint32_t& OutIndex)
	{
		auto* Stencil = dynamic_cast<StencilMaskNode<ApplyToNegative>*>(Node);
		if (!Stencil)
		{
			return false;
		}
		TgmNode Rec = {};
		Rec.Kind = ApplyToNegative ? TGM_STENCIL_NEG : TGM_STENCIL_POS;
		Rec.A = Visit(Stencil->Child.get());
		Rec.B = Visit(Stencil->StencilMask.get());
		Rec.Material = MaterialId(Stencil->Material);
		Rec.Quat[0] = 1.0f;
		Rec.Scale = 1.0f;
		OutIndex = uint32_t(Nodes.size());
		Nodes.push_back(Rec);
		return true;
	}

	uint32_t Visit(SDFNode* Node)
	{
		uint32_t Index = TGM_NONE;
		if (TryBrush<BrushNode<std::array<float, 1>>>(Node, Index)) return Index;
		if (TryBrush<BrushNode<std::array<float, 2>>>(Node, Index)) return Index;
		if (TryBrush<BrushNode<std::array<float, 3>>>(Node, Index)) return Index;
		if (TrySet<SetFamily::Union, false>(Node, Index)) return Index;
		if (TrySet<SetFamily::Union, true>(Node, Index)) return Index;
		if (TrySet<SetFamily::Inter, false>(Node, Index)) return Index;
		if (TrySet<SetFamily::Inter, true>(Node, Index)) return Index;
		if (TrySet<SetFamily::Diff, false>(Node, Index)) return Index;
		if (TrySet<SetFamily::Diff, true>(Node, Index)) return Index;
		if (FlateNode* Flate = dynamic_cast<FlateNode*>(Node))
		{
			TgmNode Rec = {};
			Rec.Kind = uint32_t(OpcodeT::Flate);
			Rec.A = Visit(Flate->Child.get());
			Rec.B = TGM_NONE;
			Rec.Material = TGM_NONE;
			Rec.Params[0] = Flate->Radius;
			Rec.Quat[0] = 1.0f;
			Rec.Scale = 1.0f;
			Index = uint32_t(Nodes.size());
			Nodes.push_back(Rec);
			return Index;
		}
		if (TryStencil<false>(Node, Index)) return Index;
		if (TryStencil<true>(Node, Index)) return Index;
		std::fprintf(stderr, "tgm: unknown node class\n");
		std::exit(2);
	}

	bool Write(SDFNodeShared& Tree, const char* Path)
	{
		uint32_t Root = Visit(Tree.get());
		FILE* File = std::fopen(Path, "wb");
		if (!File)
		{
			return false;
		}
		uint32_t Header[4] = { 0x314D4754u /* "TGM1" */, uint32_t(Nodes.size()), uint32_t(MaterialColors.size()), Root };
		std::fwrite(Header, 4, 4, File);
		for (vec3& Color : MaterialColors)
		{
			std::fwrite(&Color, 4, 3, File);
		}
		std::fwrite(Nodes.data(), sizeof(TgmNode), Nodes.size(), File);
		std::fclose(File);
		return true;
	}
};


// A material whose only property is the colour recorded in a .tgm file.
static std::vector<MaterialShared> g_TgmMaterials;

static SDFNodeShared TgmBuild(const std::vector<TgmNode>& Nodes, uint32_t Index)
{
	const TgmNode& Rec = Nodes[Index];
	MaterialShared Material = (Rec.Material == TGM_NONE) ? nullptr : g_TgmMaterials[Rec.Material];
	auto SetXform = [&](auto* Brush)
	{
		Brush->LocalToWorld.Rotation = quat(Rec.Quat[0], Rec.Quat[1], Rec.Quat[2], Rec.Quat[3]);
		Brush->LocalToWorld.Translation = vec3(Rec.Trans[0], Rec.Trans[1], Rec.Trans[2]);
		Brush->LocalToWorld.Scalation = Rec.Scale;
		Brush->BrushAABB.Min = vec3(Rec.BoundsMin[0], Rec.BoundsMin[1], Rec.BoundsMin[2]);
		Brush->BrushAABB.Max = vec3(Rec.BoundsMax[0], Rec.BoundsMax[1], Rec.BoundsMax[2]);
		Brush->Material = Material;
	};
	const float* P = Rec.Params;
	SDFNodeShared Node;
	switch (Rec.Kind)
	{
	case uint32_t(OpcodeT::Sphere): Node = SDF::Sphere(P[0]); break;
	case uint32_t(OpcodeT::Ellipsoid): Node = SDF::Ellipsoid(P[0], P[1], P[2]); break;
	case uint32_t(OpcodeT::Box): Node = SDF::Box(P[0], P[1], P[2]); break;
	case uint32_t(OpcodeT::Torus): Node = SDF::Torus(P[0], P[1]); break;
	case uint32_t(OpcodeT::Cylinder): Node = SDF::Cylinder(P[0], P[1]); break;
	case uint32_t(OpcodeT::Cone):
	{
		// Params are (Tangent, Height); build the node directly so they are not re-derived.
		std::array<float, 2> Params = { P[0], P[1] };
		BrushMixin Eval = std::bind(SDFMath::Cone, _1, P[0], P[1]);
		AABB Bounds = {};
		Node = SDFNodeShared(new BrushNode(OpcodeT::Cone, Params, Eval, Bounds));
		break;
	}
	case uint32_t(OpcodeT::Coninder):
	{
		std::array<float, 3> Params = { P[0], P[1], P[2] };
		BrushMixin Eval = std::bind(SDFMath::Coninder, _1, P[0], P[1], P[2]);
		AABB Bounds = {};
		Node = SDFNodeShared(new BrushNode(OpcodeT::Coninder, Params, Eval, Bounds));
		break;
	}
	case uint32_t(OpcodeT::Plane):
	{
		std::array<float, 3> Params = { P[0], P[1], P[2] };
		using PlanePtr = float(*)(vec3, vec3);
		BrushMixin Eval = std::bind((PlanePtr)SDFMath::Plane, _1, vec3(P[0], P[1], P[2]));
		AABB Bounds = {};
		Node = SDFNodeShared(new BrushNode(OpcodeT::Plane, Params, Eval, Bounds));
		break;
	}
	case uint32_t(OpcodeT::Union):
	case uint32_t(OpcodeT::Inter):
	case uint32_t(OpcodeT::Diff):
	case uint32_t(OpcodeT::BlendUnion):
	case uint32_t(OpcodeT::BlendInter):
	case uint32_t(OpcodeT::BlendDiff):
	{
		SDFNodeShared LHS = TgmBuild(Nodes, Rec.A);
		SDFNodeShared RHS = TgmBuild(Nodes, Rec.B);
		switch (Rec.Kind)
		{
		case uint32_t(OpcodeT::Union): return SDF::Union(LHS, RHS);
		case uint32_t(OpcodeT::Inter): return SDF::Inter(LHS, RHS);
		case uint32_t(OpcodeT::Diff): return SDF::Diff(LHS, RHS);
		case uint32_t(OpcodeT::BlendUnion): return SDF::BlendUnion(P[0], LHS, RHS);
		case uint32_t(OpcodeT::BlendInter): return SDF::BlendInter(P[0], LHS, RHS);
		default: return SDF::BlendDiff(P[0], LHS, RHS);
		}
	}
	case uint32_t(OpcodeT::Flate):
	{
		SDFNodeShared Child = TgmBuild(Nodes, Rec.A);
		return SDF::Flate(Child, P[0]);
	}
	case TGM_STENCIL_POS:
	case TGM_STENCIL_NEG:
	{
		SDFNodeShared Child = TgmBuild(Nodes, Rec.A);
		SDFNodeShared Mask = TgmBuild(Nodes, Rec.B);
		return SDF::Stencil(Child, Mask, Material, Rec.Kind == TGM_STENCIL_NEG);
	}
	default:
		std::fprintf(stderr, "tgm: bad node kind %u\n", Rec.Kind);
		std::exit(2);
	}
	// Brush fall-through: install transform, bounds and paint exactly as recorded.
	if (auto* B1 = dynamic_cast<BrushNode<std::array<float, 1>>*>(Node.get())) SetXform(B1);
	else if (auto* B2 = dynamic_cast<BrushNode<std::array<float, 2>>*>(Node.get())) SetXform(B2);
	else if (auto* B3 = dynamic_cast<BrushNode<std::array<float, 3>>*>(Node.get())) SetXform(B3);
	return Node;
}

static SDFNodeShared TgmLoad(const char* Path)
{
	FILE* File = std::fopen(Path, "rb");
	if (!File)
	{
		return nullptr;
	}
	uint32_t Header[4];
	if (std::fread(Header, 4, 4, File) != 4 || Header[0] != 0x314D4754u)
	{
		std::fclose(File);
		return nullptr;
	}
	g_TgmMaterials.clear();
	for (uint32_t i = 0; i < Header[2]; ++i)
	{
		vec3 Color;
		if (std::fread(&Color, 4, 3, File) != 3) return nullptr;
		g_TgmMaterials.push_back(MaterialShared(new MaterialPBRBR(ColorPoint(Color))));
	}
	std::vector<TgmNode> Nodes(Header[1]);
	if (std::fread(Nodes.data(), sizeof(TgmNode), Nodes.size(), File) != Nodes.size()) return nullptr;
	std::fclose(File);
	return TgmBuild(Nodes, Header[3]);
}


static bool EndsWith(const std::string& Text, const char* Suffix)
{
	size_t Len = std::strlen(Suffix);
	return Text.size() >= Len && Text.compare(Text.size() - Len, Len, Suffix) == 0;
}


static SDFNodeShared LoadModel(const std::string& Path)
{
	if (EndsWith(Path, ".tgm"))
	{
		return TgmLoad(Path.c_str());
	}
	LuaEnvironment* Env = new LuaEnvironment();
	Env->LoadFromPath(Path);
	// Env is leaked on purpose: the tree holds references into the Lua-owned material table.
	return g_CapturedTree;
}


// ------------------------------------------------------------------------------------------------
// Octree dump / checksum
// ------------------------------------------------------------------------------------------------

static uint64_t Fnv(uint64_t Hash, const void* Data, size_t Bytes)
{
	const uint8_t* Cursor = static_cast<const uint8_t*>(Data);
	for (size_t i = 0; i < Bytes; ++i)
	{
		Hash ^= Cursor[i];
		Hash *= 0x100000001B3ull;
	}
	return Hash;
}

struct OctreeStats
{
	uint64_t Nodes = 0;
	uint64_t Leaves = 0;
	uint64_t Words = 0;
	uint64_t LeafWords = 0;
	uint64_t MaxWords = 0;
	uint64_t MaxStack = 0;
	uint64_t Hash = 0xCBF29CE484222325ull;
	FILE* Dump = nullptr;
};

// Pre-order walk; per node hashes (pivot, terminus, child mask, program words).  With Dump set,
// writes the same fields as: f32 pivot[3], u32 terminus, u32 childmask, u32 nwords, u32 words[].
// A child without evaluator (a node of the live octree that was populated after its parent and lost everything,
// sdf_evaluator.cpp:1751-1757) is as good as absent -- Descend finds nothing in it either way (:1801-1835) -- and is skipped.
static void WalkOctree(SDFOctree* Node, OctreeStats& Stats)
{
	uint32_t ChildMask = 0;
	for (int i = 0; i < 8; ++i)
	{
		if (Node->Children[i] && Node->Children[i]->Evaluator) ChildMask |= (1u << i);
	}
	uint32_t Terminus = Node->Terminus ? 1 : 0;
	const std::vector<ProgramBuffer::Word>& Words = Node->Interpreter->Program.Words;
	uint32_t WordCount = uint32_t(Words.size());
	Stats.Nodes++;
	Stats.Words += WordCount;
	if (Terminus)
	{
		Stats.Leaves++;
		Stats.LeafWords += WordCount;
	}
	Stats.MaxWords = std::max<uint64_t>(Stats.MaxWords, WordCount);
	Stats.MaxStack = std::max<uint64_t>(Stats.MaxStack, Node->Interpreter->StackSize);
	// per node FNV-1a over (pivot, terminus, child mask, words); the octree hash is FNV-1a over those in pre-order
	// (two levels so that the product can hash its nodes on several threads)
	uint64_t NodeHash = 0xCBF29CE484222325ull;
	NodeHash = Fnv(NodeHash, &Node->Pivot, 12);
	NodeHash = Fnv(NodeHash, &Terminus, 4);
	NodeHash = Fnv(NodeHash, &ChildMask, 4);
	NodeHash = Fnv(NodeHash, Words.data(), WordCount * 4);
	Stats.Hash = Fnv(Stats.Hash, &NodeHash, 8);
	if (Stats.Dump)
	{
		std::fwrite(&Node->Pivot, 4, 3, Stats.Dump);
		std::fwrite(&Terminus, 4, 1, Stats.Dump);
		std::fwrite(&ChildMask, 4, 1, Stats.Dump);
		std::fwrite(&WordCount, 4, 1, Stats.Dump);
		std::fwrite(Words.data(), 4, WordCount, Stats.Dump);
	}
	for (int i = 0; i < 8; ++i)
	{
		if (Node->Children[i] && Node->Children[i]->Evaluator) WalkOctree(Node->Children[i], Stats);
	}
}


// ------------------------------------------------------------------------------------------------
// Commands
// ------------------------------------------------------------------------------------------------


// The octree the live mesher works on (sodapop.cpp:227-247, 562-600): SDFOctree::Create(Evaluator, .25, false, 3, 0.0),
// then every node the depth limit left Incomplete is populated (MeshingScratch's IncompleteSearch walk :108-120 and
// MeshingOctreeTask's `Incomplete->Populate(false, 3, -1)` :568-571).
static SDFOctreeShared CreateLiveOctree(SDFNodeShared Tree)
{
	SDFOctreeShared Octree = SDFOctree::Create(Tree, .25, false, 3, 0.0);
	if (!Octree) return Octree;
	std::vector<SDFOctree*> Incompletes;
	SDFOctree::CallbackType IncompleteSearch = [&](SDFOctree& Leaf)
	{
		if (Leaf.Incomplete) Incompletes.push_back(&Leaf);
	};
	Octree->Walk(IncompleteSearch);
	for (SDFOctree* Incomplete : Incompletes) Incomplete->Populate(false, 3, -1);
	Octree->LinkLeaves();
	return Octree;
}

// The live mesher's implicit function (sodapop.cpp:583-587).
static float LiveField(SDFOctree* Octree, vec3 Point)
{
	return glm::clamp(Octree->Eval(Point, false), -100.0f, 100.0f);
}

static int Usage()
{
	std::fprintf(stderr,
		"usage: tangerine_ref <command> ...\n"
		"  dump-tgm  <model.lua|.tgm> <out.tgm>\n"
		"  info      <model>                                  bounds / tree / octree statistics + hash (JSON)\n"
		"  info-live <model>                                  the same for the live mesher's octree (bounds = the octree's Bounds)\n"
		"  octree    <model> <out.bin>                        dump every octree node's pruned program\n"
		"  eval      <model> <mode> <points.f32> <out.bin>    mode: octree | tree | interp | gradient | color | live | live-gradient | raycast | magnet (6 floats per ray)\n"
		"  export    <model> <cells_per_unit> <refine> <out.ply|.stl>   reference ExportCommon (as shipped)\n"
		"  export-grid <model> <minx miny minz maxx maxy maxz> <step> <refine> <pointcloud 0|1> <out.ply|.stl>\n"
		"  vox       <model> <grid_size> <color_index> <out.vox>\n"
		"  bench     <model> <minx..maxz> <step> <threads> <slice_stride> reference thunks on std::threads (JSON)\n"
		"  slices    <model> <minx..maxz> <step> <threads> <k_begin> <k_end|0> <attributes 0|1> <out.json>   per-layer digests of the reference mesh\n"
		"  weld      <vertices.f32> <out.bin>                 MeshGenerator::Accumulate over float3 vertices (mesh_generators.cpp)\n"
		"  slices-live <model> <density> <threads> <attributes 0|1> <out.json>   the same for the live mesher (sodapop.cpp) at a meshing density\n");
	return 1;
}


static SDFOctreeShared CreateLiveOctree(SDFNodeShared Tree);

// Live: the live mesher's octree (CreateLiveOctree); bounds are then the octree's own Bounds, which its grid comes from.
static int CmdInfo(SDFNodeShared Tree, bool Live = false)
{
	AABB Bounds = Tree->Bounds();
	SDFInterpreter Root(Tree);
	auto T0 = Clock::now();
	SDFOctreeShared Octree = Live ? CreateLiveOctree(Tree) : SDFOctree::Create(Tree, 0.25);
	auto T1 = Clock::now();
	if (Live && Octree) Bounds = Octree->Bounds;
	OctreeStats Stats;
	if (Octree)
	{
		WalkOctree(Octree.get(), Stats);
	}
	std::printf("{\"bounds_min\": [%.9g, %.9g, %.9g], \"bounds_max\": [%.9g, %.9g, %.9g], \"leaf_count\": %d, "
		"\"stack_size\": %zu, \"root_words\": %zu, \"has_paint\": %s, \"octree_nodes\": %llu, \"octree_leaves\": %llu, "
		"\"octree_words\": %llu, \"octree_leaf_words\": %llu, \"octree_max_words\": %llu, \"octree_max_stack\": %llu, "
		"\"octree_hash\": \"%016llx\", \"octree_build_s\": %.6f}\n",
		Bounds.Min.x, Bounds.Min.y, Bounds.Min.z, Bounds.Max.x, Bounds.Max.y, Bounds.Max.z, Tree->LeafCount(),
		Tree->StackSize, Root.Program.Size(), Tree->HasPaint() ? "true" : "false",
		(unsigned long long)Stats.Nodes, (unsigned long long)Stats.Leaves, (unsigned long long)Stats.Words,
		(unsigned long long)Stats.LeafWords, (unsigned long long)Stats.MaxWords, (unsigned long long)Stats.MaxStack,
		(unsigned long long)Stats.Hash, Seconds(T0, T1));
	return 0;
}


static std::vector<float> ReadFloats(const char* Path)
{
	std::vector<float> Data;
	FILE* File = std::fopen(Path, "rb");
	if (!File) return Data;
	std::fseek(File, 0, SEEK_END);
	long Bytes = std::ftell(File);
	std::fseek(File, 0, SEEK_SET);
	Data.resize(Bytes / 4);
	if (std::fread(Data.data(), 4, Data.size(), File) != Data.size()) Data.clear();
	std::fclose(File);
	return Data;
}


static int CmdEval(SDFNodeShared Tree, const std::string& Mode, const char* PointsPath, const char* OutPath)
{
	std::vector<float> Points = ReadFloats(PointsPath);
	size_t Count = Points.size() / 3;
	FILE* Out = std::fopen(OutPath, "wb");
	if (!Out) return 2;
	if (Mode == "raycast" || Mode == "magnet")
	{
		// SDFNode::RayMarch on the model, with the Lua binding's defaults (lua_sdf.cpp:410-444): 6 floats per ray in,
		// { u32 hit, f32 travel, f32 position[3] } out
		for (size_t i = 0; i + 1 < Count; i += 2)
		{
			vec3 Origin(Points[i * 3 + 0], Points[i * 3 + 1], Points[i * 3 + 2]);
			vec3 Direction(Points[i * 3 + 3], Points[i * 3 + 4], Points[i * 3 + 5]);
			if (Mode == "magnet")
			{
				Direction = glm::normalize(Direction - Origin);
			}
			RayHit Hit = Tree->RayMarch(Origin, Direction, 100, 0.001f);
			uint32_t Flag = Hit.Hit ? 1 : 0;
			std::fwrite(&Flag, 4, 1, Out);
			std::fwrite(&Hit.Travel, 4, 1, Out);
			std::fwrite(&Hit.Position, 4, 3, Out);
		}
		std::fclose(Out);
		return 0;
	}
	if (Mode == "live" || Mode == "live-gradient")
	{
		SDFOctreeShared Live = CreateLiveOctree(Tree);
		if (!Live) return 2;
		for (size_t i = 0; i < Count; ++i)
		{
			vec3 Point(Points[i * 3 + 0], Points[i * 3 + 1], Points[i * 3 + 2]);
			if (Mode == "live")
			{
				float Dist = LiveField(Live.get(), Point);
				std::fwrite(&Dist, 4, 1, Out);
			}
			else
			{
				// the live mesher's normals (sodapop.cpp:816, USE_GRADIENT_NORMALS)
				vec3 Normal = Live->Gradient(Point);
				std::fwrite(&Normal, 4, 3, Out);
			}
		}
		std::fclose(Out);
		return 0;
	}
	SDFOctreeShared Octree = SDFOctree::Create(Tree, 0.25);
	SDFInterpreter Root(Tree);
	const bool ExportColor = Tree->HasPaint();
	for (size_t i = 0; i < Count; ++i)
	{
		vec3 Point(Points[i * 3 + 0], Points[i * 3 + 1], Points[i * 3 + 2]);
		if (Mode == "octree")
		{
			// What the mesher samples (export.cpp:339-342).
			float Dist = Octree->Eval(Point);
			std::fwrite(&Dist, 4, 1, Out);
		}
		else if (Mode == "tree")
		{
			// What VoxExport samples (magica.cpp:61).
			float Dist = Tree->Eval(Point);
			std::fwrite(&Dist, 4, 1, Out);
		}
		else if (Mode == "interp")
		{
			float Dist = Root.Eval(Point);
			std::fwrite(&Dist, 4, 1, Out);
		}
		else if (Mode == "gradient")
		{
			// What WritePLY stores as the normal (export.cpp:300).
			vec3 Normal = Octree->Gradient(Point);
			std::fwrite(&Normal, 4, 3, Out);
		}
		else if (Mode == "color")
		{
			// export.cpp:303-311, including the truncating float -> u8 conversion.
			vec3 Color = vec3(1.0f);
			if (ExportColor)
			{
				MaterialShared Material = Octree->GetMaterial(Point);
				if (Material)
				{
					Color = SampleColor(Material->GuessColor());
				}
			}
			uint8_t Bytes[3];
			Bytes[0] = 0xFF * Color.r;
			Bytes[1] = 0xFF * Color.g;
			Bytes[2] = 0xFF * Color.b;
			std::fwrite(Bytes, 1, 3, Out);
		}
		else
		{
			return Usage();
		}
	}
	std::fclose(Out);
	return 0;
}


static ExportFormat FormatFromPath(const std::string& Path)
{
	if (EndsWith(Path, ".stl")) return ExportFormat::STL;
	return ExportFormat::PLY;
}


static void ResetExportAtomics()
{
	ExportActive.store(true);
	ExportState.store(0);
	GenerationProgress.store(0);
	RefinementProgress.store(0);
	SecondaryProgress.store(0);
	WriteProgress.store(0);
	ExportState.store(1);
}


struct GridArgs
{
	vec3 Min;
	vec3 Max;
	vec3 Step;
};

static GridArgs ParseGrid(char** Argv)
{
	GridArgs Grid;
	Grid.Min = vec3(std::atof(Argv[0]), std::atof(Argv[1]), std::atof(Argv[2]));
	Grid.Max = vec3(std::atof(Argv[3]), std::atof(Argv[4]), std::atof(Argv[5]));
	Grid.Step = vec3(float(std::atof(Argv[6])));
	return Grid;
}


// Reference surface-nets thunks driven by plain std::threads, the way sodapop.cpp:685-713 drives them.
// slice_stride > 1 times only every Nth z-slice of loop 1 (stratified sample for very large grids).
static int CmdBench(SDFNodeShared Tree, GridArgs Args, int ThreadCount, int SliceStride)
{
	auto T0 = Clock::now();
	SDFOctreeShared Octree = SDFOctree::Create(Tree, 0.25);
	auto T1 = Clock::now();

	vec3 ModelMin = Args.Min - Args.Step * vec3(2.0);
	ivec3 Extent = ivec3(ceil((Args.Max - ModelMin) / Args.Step));

	isosurface::AsyncParallelSurfaceNets Task;
	Task.Grid.x = ModelMin.x;
	Task.Grid.y = ModelMin.y;
	Task.Grid.z = ModelMin.z;
	Task.Grid.dx = Args.Step.x;
	Task.Grid.dy = Args.Step.y;
	Task.Grid.dz = Args.Step.z;
	Task.Grid.sx = Extent.x;
	Task.Grid.sy = Extent.y;
	Task.Grid.sz = Extent.z;
	Task.ImplicitFunction = [&](float X, float Y, float Z) -> float
	{
		return Octree->Eval(vec3(X, Y, Z));
	};
	Task.Setup();

	// Loop 1 over z slices (all config grids have z as a non-strictly-shorter axis or are cubic; the
	// reference's axis swap quirk is avoided by iterating cells explicitly).
	std::vector<size_t> Slices;
	for (size_t k = 0; k < Task.Grid.sz; k += SliceStride)
	{
		Slices.push_back(k);
	}
	// Work items are (slice, row) pairs so that a bounded sample of a few slices still keeps every host
	// thread busy (the reference's own PSTL loop hands out whole slices, surface_nets.cpp:1174-1195).
	std::atomic<size_t> NextRow(0);
	const size_t RowCount = Slices.size() * size_t(Task.Grid.sy);
	auto T2 = Clock::now();
	{
		std::vector<std::thread> Threads;
		for (int t = 0; t < ThreadCount; ++t)
		{
			Threads.emplace_back([&]()
			{
				while (true)
				{
					size_t Index = NextRow.fetch_add(1);
					if (Index >= RowCount) break;
					size_t k = Slices[Index / size_t(Task.Grid.sy)];
					size_t j = Index % size_t(Task.Grid.sy);
					for (size_t i = 0; i < Task.Grid.sx; ++i)
					{
						Task.FirstLoopInnerThunk(Task, { i, j, k });
					}
				}
			});
		}
		for (auto& Thread : Threads) Thread.join();
	}
	auto T3 = Clock::now();

	// Loop 2 only makes sense when every slice was visited.
	double Loop2 = 0.0;
	if (SliceStride == 1)
	{
		std::vector<std::pair<size_t, uint64_t>> Domain(Task.SecondLoopDomain.begin(), Task.SecondLoopDomain.end());
		std::atomic<size_t> NextCell(0);
		auto T4 = Clock::now();
		std::vector<std::thread> Threads;
		for (int t = 0; t < ThreadCount; ++t)
		{
			Threads.emplace_back([&]()
			{
				while (true)
				{
					size_t Begin = NextCell.fetch_add(256);
					if (Begin >= Domain.size()) break;
					size_t End = std::min(Begin + 256, Domain.size());
					for (size_t c = Begin; c < End; ++c)
					{
						Task.SecondLoopThunk(Task, Domain[c]);
					}
				}
			});
		}
		for (auto& Thread : Threads) Thread.join();
		Loop2 = Seconds(T4, Clock::now());
	}

	double Loop1 = Seconds(T2, T3);
	double CellsTimed = double(Slices.size()) * double(Task.Grid.sx) * double(Task.Grid.sy);
	double CellsTotal = double(Task.Grid.sx) * double(Task.Grid.sy) * double(Task.Grid.sz);
	std::printf("{\"grid\": [%zu, %zu, %zu], \"threads\": %d, \"slice_stride\": %d, \"slices_timed\": %zu, "
		"\"octree_build_s\": %.6f, \"loop1_s\": %.6f, \"loop2_s\": %.6f, \"cells_timed\": %.0f, \"cells_total\": %.0f, "
		"\"vertices\": %zu, \"faces\": %zu, \"mvoxels_per_s_loop1\": %.6f}\n",
		Task.Grid.sx, Task.Grid.sy, Task.Grid.sz, ThreadCount, SliceStride, Slices.size(),
		Seconds(T0, T1), Loop1, Loop2, CellsTimed, CellsTotal,
		Task.OutputMesh.vertex_count(), Task.OutputMesh.face_count(), CellsTimed / Loop1 * 1e-6);
	return 0;
}


// ------------------------------------------------------------------------------------------------
// `slices`: the reference's loop 1 + loop 2 + attribute loop over a whole grid (or cell layers [k0, k1)) on
// std::threads, reported as per-z-slice digests in the serial (k, j, i) order -- the fixture the GPU parity tests
// compare the BENCHED grid sizes against (tests/golden/make_slices.py).  Vertex ids are renumbered to that order so
// that triangle indices can be compared as they are.
// ------------------------------------------------------------------------------------------------

struct Sha256
{
	uint32_t H[8] = { 0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u };
	uint8_t Block[64];
	size_t Fill = 0;
	uint64_t Total = 0;

	static uint32_t Rotr(uint32_t X, int N) { return (X >> N) | (X << (32 - N)); }

	void Compress()
	{
		static const uint32_t K[64] = {
			0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
			0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
			0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
			0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2 };
		uint32_t W[64];
		for (int t = 0; t < 16; ++t) W[t] = (uint32_t(Block[t * 4]) << 24) | (uint32_t(Block[t * 4 + 1]) << 16) | (uint32_t(Block[t * 4 + 2]) << 8) | uint32_t(Block[t * 4 + 3]);
		for (int t = 16; t < 64; ++t)
		{
			uint32_t S0 = Rotr(W[t - 15], 7) ^ Rotr(W[t - 15], 18) ^ (W[t - 15] >> 3);
			uint32_t S1 = Rotr(W[t - 2], 17) ^ Rotr(W[t - 2], 19) ^ (W[t - 2] >> 10);
			W[t] = W[t - 16] + S0 + W[t - 7] + S1;
		}
		uint32_t A = H[0], B = H[1], C = H[2], D = H[3], E = H[4], F = H[5], G = H[6], Hh = H[7];
		for (int t = 0; t < 64; ++t)
		{
			uint32_t T1 = Hh + (Rotr(E, 6) ^ Rotr(E, 11) ^ Rotr(E, 25)) + ((E & F) ^ (~E & G)) + K[t] + W[t];
			uint32_t T2 = (Rotr(A, 2) ^ Rotr(A, 13) ^ Rotr(A, 22)) + ((A & B) ^ (A & C) ^ (B & C));
			Hh = G; G = F; F = E; E = D + T1; D = C; C = B; B = A; A = T1 + T2;
		}
		H[0] += A; H[1] += B; H[2] += C; H[3] += D; H[4] += E; H[5] += F; H[6] += G; H[7] += Hh;
	}

	void Update(const void* Data, size_t Bytes)
	{
		const uint8_t* P = static_cast<const uint8_t*>(Data);
		Total += Bytes;
		while (Bytes)
		{
			size_t N = std::min(Bytes, size_t(64) - Fill);
			std::memcpy(Block + Fill, P, N);
			Fill += N; P += N; Bytes -= N;
			if (Fill == 64) { Compress(); Fill = 0; }
		}
	}

	std::string Hex(int Chars = 64)
	{
		uint64_t Bits = Total * 8;
		uint8_t Pad = 0x80;
		Update(&Pad, 1);
		uint8_t Zero = 0;
		while (Fill != 56) Update(&Zero, 1);
		uint8_t Len[8];
		for (int i = 0; i < 8; ++i) Len[i] = uint8_t(Bits >> (56 - 8 * i));
		Update(Len, 8);
		char Out[65];
		for (int i = 0; i < 8; ++i) std::snprintf(Out + i * 8, 9, "%08x", H[i]);
		return std::string(Out).substr(0, size_t(Chars));
	}
};

// Bit pattern of a float with the sign of zero and NaN payloads canonicalised (the tests compare "equal up to the sign
// of zero and NaN payloads", tests/golden_util.py same_floats).
static uint32_t CanonicalBits(float Value)
{
	if (Value != Value) return 0x7FC00000u;
	if (Value == 0.0f) return 0u;
	uint32_t Bits;
	std::memcpy(&Bits, &Value, 4);
	return Bits;
}

template <typename Fn>
static void ParallelFor(size_t Count, int ThreadCount, size_t Chunk, Fn Body)
{
	std::atomic<size_t> Next(0);
	std::vector<std::thread> Threads;
	for (int t = 0; t < ThreadCount; ++t)
	{
		Threads.emplace_back([&]()
		{
			while (true)
			{
				size_t Begin = Next.fetch_add(Chunk);
				if (Begin >= Count) break;
				size_t End = std::min(Begin + Chunk, Count);
				for (size_t i = Begin; i < End; ++i) Body(i);
			}
		});
	}
	for (auto& Thread : Threads) Thread.join();
}

// LiveDensity > 0: the live mesher instead of the export (sodapop.cpp:153-179 grid from the octree bounds and the
// meshing density, :583-587 clamped inexact field, :624-652 loop 1 over the point cache of the octree leaves).
static int CmdSlices(SDFNodeShared Tree, GridArgs Args, int ThreadCount, long K0, long K1, int Attributes, const char* OutPath, float LiveDensity = 0.0f)
{
	const bool Live = LiveDensity > 0.0f;
	auto T0 = Clock::now();
	SDFOctreeShared Octree = Live ? CreateLiveOctree(Tree) : SDFOctree::Create(Tree, 0.25);
	if (!Octree) return 2;
	auto T1 = Clock::now();

	vec3 ModelMin = Args.Min - Args.Step * vec3(2.0);
	ivec3 Extent = ivec3(ceil((Args.Max - ModelMin) / Args.Step));

	isosurface::AsyncParallelSurfaceNets Task;
	Task.Grid.x = ModelMin.x;
	Task.Grid.y = ModelMin.y;
	Task.Grid.z = ModelMin.z;
	Task.Grid.dx = Args.Step.x;
	Task.Grid.dy = Args.Step.y;
	Task.Grid.dz = Args.Step.z;
	Task.Grid.sx = Extent.x;
	Task.Grid.sy = Extent.y;
	Task.Grid.sz = Extent.z;
	Task.ImplicitFunction = [&](float X, float Y, float Z) -> float
	{
		return Octree->Eval(vec3(X, Y, Z));
	};
	if (Live)
	{
		isosurface::regular_grid_t& Grid = Task.Grid;
		const float Density = glm::floor(LiveDensity);
		const glm::vec3 SamplesPerUnit = glm::max(Octree->Bounds.Extent() * glm::vec3(Density), glm::vec3(8.0));
		Grid.x = Octree->Bounds.Min.x;
		Grid.y = Octree->Bounds.Min.y;
		Grid.z = Octree->Bounds.Min.z;
		Grid.sx = size_t(glm::ceil(SamplesPerUnit.x));
		Grid.sy = size_t(glm::ceil(SamplesPerUnit.y));
		Grid.sz = size_t(glm::ceil(SamplesPerUnit.z));
		Grid.dx = Octree->Bounds.Extent().x / static_cast<float>(Grid.sx);
		Grid.dy = Octree->Bounds.Extent().y / static_cast<float>(Grid.sy);
		Grid.dz = Octree->Bounds.Extent().z / static_cast<float>(Grid.sz);
		Grid.x -= Grid.dx * 2;
		Grid.y -= Grid.dy * 2;
		Grid.z -= Grid.dz * 2;
		Grid.sx += 3;
		Grid.sy += 3;
		Grid.sz += 3;
		SDFOctree* Raw = Octree.get();
		Task.ImplicitFunction = [Raw](float X, float Y, float Z) -> float
		{
			return LiveField(Raw, vec3(X, Y, Z));
		};
	}
	Task.Setup();
	const size_t SX = Task.Grid.sx, SY = Task.Grid.sy, SZ = Task.Grid.sz;
	// live: the cells of the point cache (:624-652), every octree leaf's box rounded outwards, as a sorted set
	std::vector<size_t> LiveCells;
	if (Live)
	{
		std::set<size_t> Cache;
		const size_t IndexRange = SX * SY * SZ;
		SDFOctree::CallbackType Gather = [&](SDFOctree& LeafNode)
		{
			if (!LeafNode.Evaluator) return;
			glm::vec3 Origin(Task.Grid.x, Task.Grid.y, Task.Grid.z);
			glm::vec3 Step(Task.Grid.dx, Task.Grid.dy, Task.Grid.dz);
			glm::vec3 AlignedMin = glm::floor(glm::max(glm::vec3(0.0), LeafNode.Bounds.Min - Origin) / Step);
			glm::vec3 AlignedMax = glm::ceil((LeafNode.Bounds.Max - Origin) / Step);
			for (float z = AlignedMin.z; z <= AlignedMax.z; ++z)
			{
				for (float y = AlignedMin.y; y <= AlignedMax.y; ++y)
				{
					for (float x = AlignedMin.x; x <= AlignedMax.x; ++x)
					{
						const size_t Index = size_t(x) + size_t(y) * SX + size_t(z) * SX * SY;
						if (Index < IndexRange) Cache.insert(Index); // `Bin < PointCache.size()` (:645)
					}
				}
			}
		};
		Octree->Walk(Gather);
		LiveCells.assign(Cache.begin(), Cache.end());
	}
	if (K1 <= 0 || size_t(K1) > SZ) K1 = long(SZ);
	if (K0 < 0) K0 = 0;
	// loop 2 of layer K0 needs the vertices of layer K0 - 1
	const size_t First = K0 > 0 ? size_t(K0 - 1) : 0;
	const size_t Layers = size_t(K1) - First;

	// Loop 1: every cell of the layers, work items are (layer, row) pairs.
	// TG_LIVE_FULL_GRID=1 walks every cell instead (checks that nothing outside the point cache would yield a vertex)
	if (Live && !std::getenv("TG_LIVE_FULL_GRID"))
	{
		ParallelFor(LiveCells.size(), ThreadCount, 64, [&](size_t c)
		{
			const size_t GridIndex = LiveCells[c];
			Task.FirstLoopInnerThunk(Task, { GridIndex % SX, (GridIndex / SX) % SY, GridIndex / (SX * SY) });
		});
	}
	else ParallelFor(Layers * SY, ThreadCount, 1, [&](size_t Index)
	{
		size_t k = First + Index / SY;
		size_t j = Index % SY;
		for (size_t i = 0; i < SX; ++i)
		{
			Task.FirstLoopInnerThunk(Task, { i, j, k });
		}
	});
	auto T2 = Clock::now();

	// Serial order: rank of every vertex among the active cells sorted by flat cell index (= k, j, i lexicographic).
	std::vector<std::pair<size_t, uint64_t>> Cells(Task.SecondLoopDomain.begin(), Task.SecondLoopDomain.end());
	std::sort(Cells.begin(), Cells.end());
	const size_t VertexCount = Cells.size();
	std::vector<uint32_t> Rank(VertexCount);
	for (size_t r = 0; r < VertexCount; ++r) Rank[size_t(Cells[r].second)] = uint32_t(r);

	// Loop 2 over the cells of layers [K0, K1).
	std::vector<std::pair<size_t, uint64_t>> Domain;
	for (auto& Cell : Cells)
	{
		if (Cell.first / (SX * SY) >= size_t(K0)) Domain.push_back(Cell);
	}
	ParallelFor(Domain.size(), ThreadCount, 256, [&](size_t c)
	{
		Task.SecondLoopThunk(Task, Domain[c]);
	});
	auto T3 = Clock::now();

	// Quads (two consecutive faces appended under one lock) in owner-cell order; the emission order inside a cell
	// (edge 0, 1, 2) survives the stable sort.
	const auto& Faces = Task.OutputMesh.faces_;
	const size_t QuadCount = Faces.size() / 2;
	std::vector<uint32_t> QuadOrder(QuadCount);
	std::iota(QuadOrder.begin(), QuadOrder.end(), 0u);
	std::stable_sort(QuadOrder.begin(), QuadOrder.end(), [&](uint32_t A, uint32_t B)
	{
		return Rank[size_t(Faces[size_t(A) * 2].v0)] < Rank[size_t(Faces[size_t(B) * 2].v0)];
	});

	// Attribute loop of WritePLY (export.cpp:297-312), threaded; vertices in serial order.
	const auto& Vertices = Task.OutputMesh.vertices_;
	const bool ExportColor = !Live && Octree->Evaluator->HasPaint(); // the live mesher colours by material evaluation, not GuessColor
	std::vector<vec3> Normals;
	std::vector<uint8_t> Colors;
	if (Attributes)
	{
		Normals.resize(VertexCount);
		if (ExportColor) Colors.resize(VertexCount * 3);
		ParallelFor(VertexCount, ThreadCount, 256, [&](size_t r)
		{
			const auto& P = Vertices[size_t(Cells[r].second)];
			vec3 Point(P.x, P.y, P.z);
			Normals[r] = Octree->Gradient(Point);
			if (ExportColor)
			{
				vec3 Color = vec3(1.0f);
				MaterialShared Material = Octree->GetMaterial(Point);
				if (Material)
				{
					Color = SampleColor(Material->GuessColor());
				}
				Colors[r * 3 + 0] = 0xFF * Color.r;
				Colors[r * 3 + 1] = 0xFF * Color.g;
				Colors[r * 3 + 2] = 0xFF * Color.b;
			}
		});
	}
	auto T4 = Clock::now();

	FILE* Out = std::fopen(OutPath, "w");
	if (!Out) return 2;
	if (Live)
	{
		std::fprintf(Out, "{\"live_grid\": {\"origin_bits\": [%u, %u, %u], \"step_bits\": [%u, %u, %u], \"cache_cells\": %zu},\n",
			CanonicalBits(Task.Grid.x), CanonicalBits(Task.Grid.y), CanonicalBits(Task.Grid.z), CanonicalBits(Task.Grid.dx), CanonicalBits(Task.Grid.dy), CanonicalBits(Task.Grid.dz),
			LiveCells.size());
	}
	std::fprintf(Out, "%s\"grid\": [%zu, %zu, %zu], \"k_begin\": %ld, \"k_end\": %ld, \"threads\": %d, \"has_color\": %s, \"attributes\": %s,\n"
		" \"octree_build_s\": %.3f, \"loop1_s\": %.3f, \"loop2_s\": %.3f, \"attributes_s\": %.3f,\n",
		Live ? " " : "{", SX, SY, SZ, K0, K1, ThreadCount, ExportColor ? "true" : "false", Attributes ? "true" : "false",
		Seconds(T0, T1), Seconds(T1, T2), Seconds(T2, T3), Seconds(T3, T4));
	// per layer: [k, vertices, triangles, sha(positions), sha(normals), sha(colours), sha(triangle indices)]
	std::fprintf(Out, " \"digest\": \"sha256 (first 16 hex digits) over the layer's records in (k, j, i) order: positions / normals as canonical float32 bits x3 "
		"(-0 -> +0, NaN -> 7fc00000), colours u8 x3, triangles u32 x3 in the serial vertex numbering (rank among active cells), two per quad, by owner cell then edge\",\n");
	std::fprintf(Out, " \"layers\": [\n");
	size_t v = 0, q = 0, OwnedVertices = 0, HelperVertices = 0;
	bool FirstRow = true;
	Sha256 All;
	for (size_t k = First; k < size_t(K1); ++k)
	{
		Sha256 HP, HN, HC, HT;
		size_t nv = 0, nt = 0;
		for (; v < VertexCount && Cells[v].first / (SX * SY) == k; ++v, ++nv)
		{
			const auto& P = Vertices[size_t(Cells[v].second)];
			uint32_t Bits[3] = { CanonicalBits(P.x), CanonicalBits(P.y), CanonicalBits(P.z) };
			HP.Update(Bits, 12);
			if (Attributes)
			{
				uint32_t NBits[3] = { CanonicalBits(Normals[v].x), CanonicalBits(Normals[v].y), CanonicalBits(Normals[v].z) };
				HN.Update(NBits, 12);
				if (ExportColor) HC.Update(&Colors[v * 3], 3);
			}
		}
		for (; q < QuadCount; ++q)
		{
			const auto& F0 = Faces[size_t(QuadOrder[q]) * 2];
			const auto& F1 = Faces[size_t(QuadOrder[q]) * 2 + 1];
			size_t Owner = Cells[Rank[size_t(F0.v0)]].first;
			if (Owner / (SX * SY) != k) break;
			// ranks are relative to the first layer walked; subtract nothing: with K0 = 0 they are global ids
			uint32_t Tri[6] = { Rank[size_t(F0.v0)], Rank[size_t(F0.v1)], Rank[size_t(F0.v2)], Rank[size_t(F1.v0)], Rank[size_t(F1.v1)], Rank[size_t(F1.v2)] };
			HT.Update(Tri, 24);
			nt += 2;
		}
		if (k < size_t(K0))
		{
			HelperVertices = nv; // the helper layer below the range: its vertices hold the first ranks
			continue;
		}
		OwnedVertices += nv;
		if (nv == 0 && nt == 0) continue;
		std::string SP = HP.Hex(16), SN = HN.Hex(16), SC = HC.Hex(16), ST = HT.Hex(16);
		std::fprintf(Out, "%s  [%zu, %zu, %zu, \"%s\", \"%s\", \"%s\", \"%s\"]", FirstRow ? "" : ",\n", k, nv, nt, SP.c_str(), SN.c_str(), SC.c_str(), ST.c_str());
		FirstRow = false;
	}
	std::fprintf(Out, "\n ],\n \"first_rank\": %zu, \"vertices\": %zu, \"triangles\": %zu}\n", HelperVertices, OwnedVertices, QuadCount * 2);
	std::fclose(Out);
	std::printf("{\"vertices\": %zu, \"triangles\": %zu, \"octree_build_s\": %.3f, \"loop1_s\": %.3f, \"loop2_s\": %.3f, \"attributes_s\": %.3f}\n",
		OwnedVertices, QuadCount * 2, Seconds(T0, T1), Seconds(T1, T2), Seconds(T2, T3), Seconds(T3, T4));
	return 0;
}


int main(int Argc, char** Argv)
{
	if (Argc < 3)
	{
		return Usage();
	}
	std::string Command = Argv[1];
	if (Command == "weld" && Argc == 4)
	{
		// MeshGenerator::Accumulate(vertex) over a stream of float3 (mesh_generators.cpp:42-50): out = u32 distinct count,
		// the distinct vertices as vec4, one u32 index per input vertex
		std::ifstream In(Argv[2], std::ios::binary);
		std::vector<char> Raw((std::istreambuf_iterator<char>(In)), std::istreambuf_iterator<char>());
		const size_t Count = Raw.size() / 12;
		const float* Values = reinterpret_cast<const float*>(Raw.data());
		MeshGenerator Generator;
		for (size_t i = 0; i < Count; ++i)
		{
			Generator.Accumulate(glm::vec3(Values[i * 3 + 0], Values[i * 3 + 1], Values[i * 3 + 2]));
		}
		FILE* Out = std::fopen(Argv[3], "wb");
		if (!Out) return 2;
		uint32_t Distinct = uint32_t(Generator.Vertices.size());
		std::fwrite(&Distinct, 4, 1, Out);
		std::fwrite(Generator.Vertices.data(), 16, Generator.Vertices.size(), Out);
		std::fwrite(Generator.Indices.data(), 4, Generator.Indices.size(), Out);
		std::fclose(Out);
		return 0;
	}
	SDFNodeShared Tree = LoadModel(Argv[2]);
	if (!Tree)
	{
		std::fprintf(stderr, "failed to load model %s\n", Argv[2]);
		return 2;
	}

	if (Command == "dump-tgm" && Argc == 4)
	{
		TgmWriter Writer;
		return Writer.Write(Tree, Argv[3]) ? 0 : 2;
	}
	else if (Command == "info")
	{
		return CmdInfo(Tree);
	}
	else if (Command == "info-live")
	{
		return CmdInfo(Tree, true);
	}
	else if (Command == "octree" && Argc == 4)
	{
		SDFOctreeShared Octree = SDFOctree::Create(Tree, 0.25);
		OctreeStats Stats;
		Stats.Dump = std::fopen(Argv[3], "wb");
		if (!Stats.Dump || !Octree) return 2;
		WalkOctree(Octree.get(), Stats);
		std::fclose(Stats.Dump);
		return 0;
	}
	else if (Command == "eval" && Argc == 6)
	{
		return CmdEval(Tree, Argv[3], Argv[4], Argv[5]);
	}
	else if (Command == "export" && Argc == 6)
	{
		auto T0 = Clock::now();
		ExportCommon(Tree, float(std::atof(Argv[3])), std::atoi(Argv[4]), Argv[5], FormatFromPath(Argv[5]), 1.0f);
		std::printf("{\"export_s\": %.6f}\n", Seconds(T0, Clock::now()));
		return 0;
	}
	else if (Command == "export-grid" && Argc == 13)
	{
		GridArgs Grid = ParseGrid(Argv + 3);
		int Refine = std::atoi(Argv[10]);
		bool PointCloud = std::atoi(Argv[11]) != 0;
		std::string Path = Argv[12];
		ResetExportAtomics();
		auto T0 = Clock::now();
		if (PointCloud)
		{
			PointCloudExportThread(Tree, Grid.Min, Grid.Max, Grid.Step, Refine, Path, FormatFromPath(Path), 1.0f);
		}
		else
		{
			MeshExportThread(Tree, Grid.Min, Grid.Max, Grid.Step, Refine, Path, FormatFromPath(Path), 1.0f);
		}
		std::printf("{\"export_s\": %.6f}\n", Seconds(T0, Clock::now()));
		return 0;
	}
	else if (Command == "vox" && Argc == 6)
	{
		std::string Path = Argv[5];
		auto T0 = Clock::now();
		VoxExport(Tree, Path, float(std::atof(Argv[3])), std::atoi(Argv[4]));
		std::printf("{\"vox_s\": %.6f}\n", Seconds(T0, Clock::now()));
		return 0;
	}
	else if (Command == "slices" && Argc == 15)
	{
		GridArgs Grid = ParseGrid(Argv + 3);
		return CmdSlices(Tree, Grid, std::atoi(Argv[10]), std::atol(Argv[11]), std::atol(Argv[12]), std::atoi(Argv[13]), Argv[14]);
	}
	else if (Command == "slices-live" && Argc == 7)
	{
		GridArgs Unused;
		Unused.Min = Unused.Max = vec3(0.0f);
		Unused.Step = vec3(1.0f);
		return CmdSlices(Tree, Unused, std::atoi(Argv[4]), 0, 0, std::atoi(Argv[5]), Argv[6], float(std::atof(Argv[3])));
	}
	else if (Command == "bench" && Argc == 12)
	{
		GridArgs Grid = ParseGrid(Argv + 3);
		return CmdBench(Tree, Grid, std::atoi(Argv[10]), std::atoi(Argv[11]));
	}
	return Usage();
}
