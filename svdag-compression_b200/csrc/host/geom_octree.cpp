#include "geom_octree.hpp"
#include "sharded_build.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <vector>

namespace svbhost {

GeomOctree::GeomOctree(Scene* scene, int device) : GeomOctree(scene, std::vector<int>(1, device)) {}

GeomOctree::GeomOctree(Scene* scene, const std::vector<int>& devices) : _scene(scene), _devices(devices) {
	memset(&_stats, 0, sizeof(_stats));
	if (_devices.empty()) _devices.push_back(0);
	for (int d : _devices) {
		svb_ctx* c = svb_create(d);
		if (!c) {
			fprintf(stderr, "ERROR: no usable CUDA device %d (this build of svbuilder has no CPU path)\n", d);
			exit(1);
		}
		_ctxs.push_back(c);
	}
	_ctx = _ctxs[0];
}
GeomOctree::~GeomOctree() { for (svb_ctx* c : _ctxs) svb_destroy(c); }

void GeomOctree::check(int rc, const char* what) {
	if (rc == SVB_OK) return;
	fprintf(stderr, "ERROR in %s: %s (code %d)\n", what, svb_last_error(_ctx), rc);
	exit(1);   // the reference's fatal paths exit(1) too (e.g. encoded_ssvdag.cpp:251-254)
}

void GeomOctree::buildSVO(unsigned levels, const double bmin[3], const double bmax[3]) {
	printf("* Building SVO... (deferred: voxelization and reduction run together on the GPU in toDAG)\n");
	_pendingLevels = levels;
	memcpy(_pendMin, bmin, 24);
	memcpy(_pendMax, bmax, 24);
	_state = S_SVO;
}

void GeomOctree::toDAG(bool internalCall) {
	if (_state != S_SVO) { printf("ERROR! This is not a SVO!\n"); return; }   // geom_octree.cpp:466-469
	if (!internalCall) { printf("* Transforming SVO -> DAG ... \n"); fflush(stdout); }
	if (!_trisUploaded) { check(svb_set_triangles(_ctx, _scene->getTrianglePtr(), _scene->getNRawTriangles()), "svb_set_triangles"); _trisUploaded = true; }
	check(svb_build(_ctx, _pendingLevels, 0, _pendMin, _pendMax, &_stats), "svb_build");
	_levels = _pendingLevels;
	for (unsigned lev = _levels - 1; lev > 0; --lev) {
		uint64_t a = 0, b = 0;
		svb_level_count_svo(_ctx, lev, &a);
		svb_level_count(_ctx, lev, &b);
		if (!internalCall) printf("Reduced level %u from %lu to %lu nodes\n", lev, (unsigned long)a, (unsigned long)b);   // :509
	}
	_state = S_DAG;
	if (!internalCall) printf("OK! [%.2f ms on the GPU]\n", _stats.msTotal);
}

void GeomOctree::buildDAG(unsigned levels, unsigned stepLevel, const double bmin[3], const double bmax[3], bool verbose) {
	printf("* Building DAG [stepLevel: %i]\n", stepLevel); fflush(stdout);
	if (_ctxs.size() > 1 && stepLevel > 0) {
		printf("\t- %zu GPUs: sub-octrees dealt over the devices, levels merged over NCCL\n", _ctxs.size()); fflush(stdout);
		std::string err;
		if (!build_dag_sharded(_devices, _ctxs, _scene->getTrianglePtr(), _scene->getNRawTriangles(), levels, stepLevel, bmin, bmax, &_stats, &err, &_msUpload, &_msExchange)) {
			fprintf(stderr, "ERROR in the multi-GPU build: %s\n", err.c_str());
			exit(1);
		}
		_trisUploaded = false;
	} else {
		if (!_trisUploaded) { check(svb_set_triangles(_ctx, _scene->getTrianglePtr(), _scene->getNRawTriangles()), "svb_set_triangles"); _trisUploaded = true; }
		check(svb_build(_ctx, levels, stepLevel, bmin, bmax, &_stats), "svb_build");
	}
	_levels = levels;
	_state = S_DAG;
	if (verbose) printf("\t- %lu subtrees in %lu device batches, %lu (triangle,node) pairs\n", (unsigned long)_stats.nTiles, (unsigned long)_stats.nBatches, (unsigned long)_stats.nPairsTotal);
	printf("\t- Finished! Total time [%.2f ms on the GPU: voxelize %.2f, reduce %.2f, rank %.2f]\n", _stats.msTotal, _stats.msVoxelize, _stats.msDedup, _stats.msFinalize);
}

void GeomOctree::toSDAG(bool internalCall, bool skipSymmetry) {
	if (skipSymmetry) { fprintf(stderr, "toSDAG(skipSymmetry=true) is not used by svbuilder and not provided\n"); exit(1); }
	if (_state != S_DAG) { printf("ERROR! This is not a DAG or SDAG!\n"); return; }   // :560-563
	if (!internalCall) { printf("* Transforming DAG -> SDAG (normal)... "); fflush(stdout); }
	check(svb_to_sdag(_ctx, &_stats), "svb_to_sdag");
	_state = S_SDAG;
	if (!internalCall) printf("OK! [%.2f ms]\n", _stats.msSdag);
}

unsigned GeomOctree::mergeAcrossAllLevels() {
	check(svb_cross_merge(_ctx, &_stats), "svb_cross_merge");
	return (unsigned)_stats.nCrossLevelMerged;
}

bool GeomOctree::loadSVDAG(const std::string& fileName) {
	printf("* Loading SVDAG '%s'... ", fileName.c_str()); fflush(stdout);
	std::ifstream in(fileName, std::ios::binary);
	if (!in.is_open()) { printf("FAILED!!!\n"); return false; }
	std::vector<uint8_t> img((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
	printf("OK!\nDecoding nodes...\n");
	check(svb_load_svdag(_ctx, img.data(), img.size(), &_stats), "svb_load_svdag");
	_levels = svb_levels(_ctx);
	_state = S_DAG;
	return true;
}

void GeomOctree::resizeSceneBbox(const float mn[3], const float mx[3]) {
	_bboxOverride = true;
	float s = 0;
	for (int k = 0; k < 3; ++k) { _bboxF[k] = mn[k]; _bboxF[3 + k] = mx[k]; float side = ((mx[k] - mn[k]) * 0.5f) * 2.0f; if (k == 0 || side > s) s = side; }
	_rootSideOverride = s;
}

OctreeData GeomOctree::getNodeData() {
	OctreeData o;
	o.levels.resize(_levels);
	for (unsigned l = 0; l < _levels; ++l) {
		LevelSoA& h = o.levels[l];
		uint64_t n = 0;
		check(svb_level_count(_ctx, l, &n), "svb_level_count");
		h.n = n;
		h.mask.resize(n); h.child.resize(n * 8); h.mirror.resize(n * 3); h.inv.resize(n); h.childLevel.resize(n * 8);
		check(svb_download_level(_ctx, l, h.mask.data(), h.child.data(), h.mirror.data(), h.inv.data(), h.childLevel.data()), "svb_download_level");
	}
	memcpy(o.bboxF, _bboxOverride ? _bboxF : _stats.bboxF, 24);
	o.rootSide = _bboxOverride ? _rootSideOverride : _stats.rootSide;
	o.nNodes = _stats.nNodes;
	o.nVoxels = _stats.nTotalVoxels;
	o.state = (int)_state;
	return o;
}

bool GeomOctree::encodeToFile(int kind, const std::string& fileName, size_t* fileBytes, size_t* dataBytes) {
	const uint8_t* img = nullptr;
	uint64_t size = 0;
	int rc = svb_encode_view(_ctx, kind, &img, &size);
	if (rc != SVB_OK) { printf("%s\n", svb_last_error(_ctx)); return false; }   // e.g. "FAILED! Octree is not in DAG state" (encoded_svdag.cpp:109-112)
	FILE* f = fopen(fileName.c_str(), "wb");
	if (!f) { printf("FAILED!!!\n"); return false; }
	bool ok = true;
	if (_bboxOverride) {   // main.cpp:179-193: the header carries the rescaled scene box and root side
		uint8_t head[28];
		memcpy(head, _bboxF, 24);
		const float rs = (float)_rootSideOverride;
		memcpy(head + 24, &rs, 4);
		ok = fwrite(head, 1, 28, f) == 28 && fwrite(img + 28, 1, size - 28, f) == size - 28;
	} else ok = fwrite(img, 1, size, f) == size;
	fclose(f);
	if (!ok) { printf("FAILED!!!\n"); return false; }
	if (fileBytes) *fileBytes = size;
	if (dataBytes) *dataBytes = size - (kind == SVB_FILE_SSVDAG ? 36 + 12 : 44);
	return true;
}

}  // namespace svbhost
