// File writers for the export path: byte layouts of the reference's WritePLY / WriteSTL
// (tangerine/export.cpp:60-108, 198-280) and of the MagicaVoxel files its VoxExport produces through
// third_party/voxwriter (tangerine/magica.cpp:27-72, VoxWriter.cpp:482-670).
#pragma once

#include <cstdint>
#include <string>

namespace tg
{

bool WritePly(const char* path, const float* positions, const float* normals, const uint8_t* colors, uint64_t vertex_count,
	const uint32_t* triangles, uint64_t triangle_count, std::string& error);

// face_normals (one per triangle) are what the export path writes; without them the normal of each
// triangle is the normalised sum of its vertex normals, as in the reference's drawable export (export.cpp:83-86).
bool WriteStl(const char* path, const float* positions, const float* vertex_normals, const float* face_normals, uint64_t vertex_count,
	const uint32_t* triangles, uint64_t triangle_count, std::string& error);

// voxels: x, y, z triples; every voxel gets palette index abs(color_index) % 255 + 1 (magica.cpp:64).
bool WriteVox(const char* path, const int32_t size[3], const int32_t* xyz, uint64_t count, int color_index, std::string& error);

} // namespace tg
