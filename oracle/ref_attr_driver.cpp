// =====================================================================================
//  TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.
//  A 40-line driver (ours) around the UNMODIFIED reference sources, compiled by oracle/Makefile
//  into oracle/_ref/ref_attr_driver: it calls the one entry of the hot path that the reference's
//  own command line never exercises,
//      GeomOctree::buildSVO(levels, bbox, false, NULL, /*putMaterialIdInLeaves=*/true)
//  (src/symvox/geom_octree.cpp:171-280; the material id of the triangle being voxelized is stored
//  in the leaf node's child slot, :210-211, :252 -- so a voxel ends up with the material of the LAST
//  triangle, in file order, that touches it), and dumps the leaf level in the reference's node
//  order:  u64 n, then n x { u8 mask, u32 children[8] }.  Scene loading and the float -> double
//  bbox widening follow src/svbuilder/main.cpp:90, :150-155.
// =====================================================================================
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <string>

#include <symvox/geom_octree.hpp>
#include <symvox/scene.hpp>

int main(int argc, char** argv) {
	if (argc < 4) { fprintf(stderr, "usage: ref_attr_driver scene.obj levels out.bin\n"); return 2; }
	const unsigned levels = (unsigned)atoi(argv[2]);
	Scene scene;
	scene.loadObj(argv[1], true, false, false, false, true);
	// The command line's load path clears the indexed triangles once the flat triangle vector is built
	// (Scene::buildTriVector(true), scene.cpp:394-416), after which getTriangleMaterialId() answers 0 for every
	// triangle (scene.hpp:83-85).  A caller that wants material ids keeps them: read the cache a second time and hand the
	// indexed triangles (with their material field) back through the public accessor.
	{
		Scene withIds;
		withIds.loadBinObj(std::string(argv[1]) + ".bincache");
		*scene.getIndexedGeom() = *withIds.getIndexedGeom();
	}
	GeomOctree octree(&scene);
	auto minF = scene.getAABB()[0], maxF = scene.getAABB()[1];
	sl::aabox3d bbox(sl::point3d(minF[0], minF[1], minF[2]), sl::point3d(maxF[0], maxF[1], maxF[2]));
	octree.buildSVO(levels, bbox, false, NULL, true);
	const auto& leaf = octree.getNodeData()[levels - 1];
	FILE* f = fopen(argv[3], "wb");
	if (!f) return 3;
	const uint64_t n = leaf.size();
	fwrite(&n, 8, 1, f);
	for (const auto& nd : leaf) {
		const uint8_t m = nd.childrenBitmask;
		uint32_t ch[8];
		for (int i = 0; i < 8; ++i) ch[i] = (uint32_t)nd.children[i];
		fwrite(&m, 1, 1, f);
		fwrite(ch, 4, 8, f);
	}
	fclose(f);
	printf("\nleaf nodes %zu\n", (size_t)n);
	return 0;
}
