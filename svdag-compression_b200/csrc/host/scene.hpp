// scene.hpp -- host-side mesh loader with the reference's Scene semantics
// (src/symvox/scene.{hpp,cpp}): ASCII OBJ ('v' / 'f' lines, triangles + the quad rule of
// scene.cpp:183), the `<obj>.bincache` binary cache (read preferred, written after an ASCII parse,
// scene.cpp:55-71, 257-259, 324-374), a float bounding box grown vertex by vertex (:131), and the
// flat float triangle soup handed to the voxelizer (buildTriVector, :394-416).
#pragma once
#include <cstddef>
#include <string>
#include <vector>

namespace svbhost {

class Scene {
public:
	// mirrors Scene::loadObj(fileName, tryLoadBinCache=true, ...); returns false if the file cannot be read
	bool loadObj(const std::string& fileName, bool tryLoadBinCache = true);
	const float* getTrianglePtr() const { return _triangles.data(); }     // 9 floats per triangle
	std::size_t getNRawTriangles() const { return _triangles.size() / 9; }
	void getBounds(float mn[3], float mx[3]) const { for (int k = 0; k < 3; ++k) { mn[k] = _bbox[k]; mx[k] = _bbox[3 + k]; } }
	void setAABB(const float mn[3], const float mx[3]) { for (int k = 0; k < 3; ++k) { _bbox[k] = mn[k]; _bbox[3 + k] = mx[k]; } }

private:
	bool loadBinObj(const std::string& fileName);
	void saveBinObj(const std::string& fileName) const;
	void buildTriVector();
	std::vector<float> _vertices;                  // xyz
	std::vector<std::size_t> _indexed;             // 3 vertex indices per triangle
	std::vector<float> _triangles;
	float _bbox[6];
};

}  // namespace svbhost
