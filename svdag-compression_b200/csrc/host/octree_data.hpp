// octree_data.hpp -- host-side SoA snapshot of a GeomOctree (what getNodeData() exposes to the
// encoders in the reference, src/symvox/geom_octree.hpp:117 + src/symvox/octree.hpp:101-106).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace svbhost {

struct LevelSoA {
	uint64_t n = 0;
	std::vector<uint8_t> mask;          // childrenBitmask
	std::vector<uint32_t> child;        // n*8, 0xFFFFFFFE = nullNode
	std::vector<uint8_t> mirror;        // n*3, childrenMirroredBitmask[x,y,z]
	std::vector<uint8_t> inv;           // invariantBitmask
	std::vector<uint32_t> childLevel;   // n*8, Octree::Node::childLevels
};

struct OctreeData {
	std::vector<LevelSoA> levels;
	float bboxF[6] = {0, 0, 0, 0, 0, 0};   // Octree::_bbox
	double rootSide = 0;                   // Octree::_rootSide
	uint64_t nNodes = 0;                   // Octree::_nNodes as the encoder reads it (getNNodes())
	uint64_t nVoxels = 0;
	int state = 0;                         // GeomOctree::State
};

// File images exactly as the reference's encode()+save() pairs write them.
// kind 0: .svdag   (EncodedSVDAG,   encoded_svdag.cpp:76-199)    needs state DAG
// kind 1: .ussvdag (EncodedUSSVDAG, encoded_ussvdag.cpp:60-170)  needs state SDAG
// kind 2: .ssvdag / .esvdag (EncodedSSVDAG, encoded_ssvdag.cpp:84-117,194-466) needs DAG or SDAG
bool encode_file(const OctreeData& o, int kind, std::vector<uint8_t>& out, std::string* err = nullptr);

// EncodedSSVDAG::encode's node order (encoded_ssvdag.cpp:247-276) from per-level reference counts: refs / order hold levels
// 0 .. nLevels-1 concatenated, level l at [start[l], start[l+1]).  order[start[l] + r] = index of the node at rank r.
void ssvdag_order_from_refs(const uint32_t* refs, const uint32_t* start, int nLevels, uint32_t* order);

// EncodedSVDAG::load + decode (encoded_svdag.cpp:43-74, :200-270): a .svdag image back into DAG levels (state DAG).
bool decode_svdag(const uint8_t* file, uint64_t size, OctreeData& o, std::string* err = nullptr);

}  // namespace svbhost
