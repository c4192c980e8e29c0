// Compiled by tests/test_cpp_mirror.py against include/tangerine_b200.hpp: a C++ caller that uses the drop-in mirror of
// the reference's export interface exactly the way tangerine.cpp:1378-1481 uses tangerine/export.h -- MeshExport on a
// detached thread, GetExportProgress polled like a UI frame loop, CancelExport from the calling thread.
//
//   mirror_test link                        no device needed: proves the mirror links and the error path is loud
//   mirror_test export MODEL.tgm OUT.ply    async export, progress polled (monotonic, several distinct values)
//   mirror_test cancel MODEL.tgm OUT.ply    CancelExport(true) mid-flight: TG_ERR_CANCELLED, no file, next export works
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "tangerine_b200.hpp"

using namespace tangerine_b200;

static bool FileExists(const char* path)
{
	FILE* f = std::fopen(path, "rb");
	if (f) std::fclose(f);
	return f != nullptr;
}

static int Fail(const char* what)
{
	std::printf("FAIL: %s (%s)\n", what, LastExportError().c_str());
	return 1;
}

int main(int argc, char** argv)
{
	const std::string mode = argc > 1 ? argv[1] : "link";
	if (mode == "link")
	{
		// No model, no device: the calls must fail with a status, not abort (the reference's Assert aborts, errors.cpp:23-32).
		const float zero[3] = { 0, 0, 0 }, one[3] = { 1, 1, 1 }, step[3] = { 0.1f, 0.1f, 0.1f };
		if (ExportCommon(nullptr, 8.0f, 0, "/tmp/never.ply", ExportFormat::PLY) == TG_OK) return Fail("ExportCommon(null) succeeded");
		ExportProgress p = GetExportProgress();
		if (p.Stage != 0) return Fail("idle stage is not 0");
		CancelExport(true);
		(void)zero; (void)one; (void)step;
		std::printf("OK link %s\n", tg_version());
		return 0;
	}
	if (argc < 4) return Fail("usage");
	tg_tree* tree = tg_tree_load(argv[2]);
	if (!tree) return Fail("cannot load model");
	float mn[3], mx[3];
	tg_tree_bounds(tree, mn, mx);
	const float cells = argc > 4 ? float(std::atof(argv[4])) : 510.0f;
	const float s = (mx[0] - mn[0]) / cells;
	const float step[3] = { s, s, s };
	const char* out = argv[3];
	std::remove(out);

	if (mode == "export")
	{
		// a second thread hammers the C ABI's tg-level progress call through the mirror concurrently
		std::atomic<bool> stop{ false };
		std::atomic<int> side_polls{ 0 };
		std::thread side([&] { while (!stop.load()) { GetExportProgress(); side_polls++; } });
		MeshExport(tree, out, mn, mx, step, 0, ExportFormat::PLY, false, 1.0f);
		std::vector<float> seen;
		std::set<int> stages;
		float last = 0.0f;
		bool monotonic = true;
		int polls = 0;
		for (;;)
		{
			ExportProgress p = GetExportProgress();
			polls++;
			stages.insert(p.Stage);
			if (p.Stage == 0) break;
			if (p.Stage == 1)
			{
				if (p.Generation + 1e-6f < last) monotonic = false;
				if (seen.empty() || p.Generation != seen.back()) seen.push_back(p.Generation);
				last = p.Generation;
			}
		}
		stop.store(true);
		side.join();
		std::printf("polls %d (side thread %d), distinct generation values %zu, stages seen %zu, status %d\n", polls, side_polls.load(), seen.size(), stages.size(), LastExportStatus());
		if (LastExportStatus() != TG_OK) return Fail("export failed");
		if (!FileExists(out)) return Fail("no output file");
		if (!monotonic) return Fail("Generation progress went backwards");
		if (seen.size() <= 3) return Fail("progress did not move in steps (needs > 3 distinct values)");
		std::printf("OK export\n");
	}
	else if (mode == "cancel")
	{
		MeshExport(tree, out, mn, mx, step, 0, ExportFormat::PLY, false, 1.0f);
		std::this_thread::sleep_for(std::chrono::milliseconds(argc > 5 ? std::atoi(argv[5]) : 20));
		CancelExport(true);
		while (GetExportProgress().Stage != 0) std::this_thread::sleep_for(std::chrono::milliseconds(1));
		std::printf("status after cancel: %d (%s)\n", LastExportStatus(), LastExportError().c_str());
		if (LastExportStatus() != TG_ERR_CANCELLED) return Fail("a cancelled export must report TG_ERR_CANCELLED");
		if (FileExists(out)) return Fail("a cancelled export must not write its file");
		// the path is reusable straight away
		if (ExportCommon(tree, 4.0f, 0, out, ExportFormat::PLY) != TG_OK) return Fail("export after a cancel failed");
		if (!FileExists(out)) return Fail("no output file after the second export");
		std::printf("OK cancel\n");
	}
	else if (mode == "live")
	{
		// Sodapop::Populate's result on a Drawable (sodapop.cpp:214-225, sdf_model.h:59-62) at the default density
		LiveDrawable painter;
		const int rc = PopulateDrawable(tree, 0.0f, painter);
		if (rc != TG_OK) return Fail("PopulateDrawable failed");
		if (painter.Positions.size() != painter.Normals.size() || painter.Positions.size() != painter.Colors.size()) return Fail("attribute arrays differ in length");
		uint64_t hash = 0xCBF29CE484222325ull; // FNV-1a over the xyz bits of every position, in order
		for (size_t v = 0; v < painter.Positions.size() / 4; ++v)
		{
			if (painter.Positions[v * 4 + 3] != 1.0f || painter.Normals[v * 4 + 3] != 1.0f || painter.Colors[v * 4 + 3] != 1.0f) return Fail("w components");
			const unsigned char* bytes = reinterpret_cast<const unsigned char*>(&painter.Positions[v * 4]);
			for (int b = 0; b < 12; ++b) hash = (hash ^ bytes[b]) * 0x100000001B3ull;
		}
		std::printf("live vertices %zu triangles %zu fnv %016llx\n", painter.Positions.size() / 4, painter.Indices.size() / 3, (unsigned long long)hash);
		std::printf("OK live\n");
	}
	tg_tree_free(tree);
	return 0;
}
