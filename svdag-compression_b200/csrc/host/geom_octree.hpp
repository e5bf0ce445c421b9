// geom_octree.hpp -- C++ host class with the call surface of the reference's GeomOctree
// (src/symvox/geom_octree.hpp:107-157), implemented on the C ABI of libsvb.so (include/svb.h).
// A maintainer of the reference swaps `#include <symvox/geom_octree.hpp>` for this header and keeps
// svbuilder/main.cpp's call sequence; all per-voxel / per-node work runs on the GPU.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "octree_data.hpp"
#include "scene.hpp"
#include "svb.h"

namespace svbhost {

class GeomOctree {
public:
	enum State { S_EMPTY = SVB_S_EMPTY, S_SVO = SVB_S_SVO, S_DAG = SVB_S_DAG, S_SDAG = SVB_S_SDAG };
	typedef svb_stats Stats;

	explicit GeomOctree(Scene* scene, int device = 0);
	// several GPUs of one node: buildDAG (step > 0) spreads the sub-octrees over them (sharded_build.hpp); everything
	// after the build (toSDAG, cross-level merge, encoders) runs on devices[0]
	GeomOctree(Scene* scene, const std::vector<int>& devices);
	~GeomOctree();
	GeomOctree(const GeomOctree&) = delete;
	GeomOctree& operator=(const GeomOctree&) = delete;

	// Main methods (geom_octree.hpp:120-134).  bbox is the double-widened scene bbox (main.cpp:150-155).
	void buildDAG(unsigned levels, unsigned stepLevel, const double bmin[3], const double bmax[3], bool verbose = false);
	void buildSVO(unsigned levels, const double bmin[3], const double bmax[3]);   // records the request; the SVO only ever feeds toDAG()
	void toDAG(bool internalCall = false);
	void toSDAG(bool internalCall = false, bool skipSymmetry = false);
	unsigned mergeAcrossAllLevels();
	// EncodedSVDAG::load + decode + GeomOctree(data, ...) (main.cpp:96-101): start from a saved .svdag
	bool loadSVDAG(const std::string& fileName);
	void initChildLevels() {}   // pointers of a fresh DAG always target the next level; svb_download_level reports lev+1

	State getState() const { return _state; }
	Stats getStats() const { return _stats; }
	unsigned getLevels() const { return _levels; }
	size_t getNVoxels() const { return (size_t)_stats.nTotalVoxels; }
	size_t getNNodes() const { return (size_t)_stats.nNodes; }
	float getRootSide() const { return (float)_stats.rootSide; }
	void resizeSceneBbox(const float mn[3], const float mx[3]);   // octree.hpp:108-112 (main.cpp:192)

	// getNodeData(): copies the levels D2H
	OctreeData getNodeData();
	// encode(const GeomOctree&) + save() of the three encoders: the image is written on the GPU (svb_encode_view) and
	// handed to fwrite from pinned memory.  *dataBytes receives what the reference's getDataSize() reports for that
	// encoder (the payload without headers and length prefixes: encoded_svdag.hpp:47, encoded_ssvdag.hpp:43-45).
	bool encodeToFile(int kind, const std::string& fileName, size_t* fileBytes = nullptr, size_t* dataBytes = nullptr);
	double lastUploadMs() const { return _msUpload; }
	double lastExchangeMs() const { return _msExchange; }
	size_t nDevices() const { return _devices.size(); }

private:
	void check(int rc, const char* what);
	Scene* _scene;
	svb_ctx* _ctx;                       // == _ctxs[0]
	std::vector<int> _devices;
	std::vector<svb_ctx*> _ctxs;
	double _msUpload = 0, _msExchange = 0;
	State _state = S_EMPTY;
	Stats _stats;
	unsigned _levels = 0;
	bool _trisUploaded = false;
	unsigned _pendingLevels = 0;
	double _pendMin[3], _pendMax[3];
	bool _bboxOverride = false;
	float _bboxF[6];
	double _rootSideOverride = 0;
};

}  // namespace svbhost
