#include "scene.hpp"

#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>

namespace svbhost {

namespace {
// on-disk layout of the reference's cache (scene.hpp:112-119 header; 300-byte TMaterial, 56-byte TIndexedTri)
struct BinObjHeader { uint64_t nVertices, nNormals, nMaterials, nIndexedTris; float bboxMin[3], bboxMax[3]; };
const size_t kMaterialBytes = 300, kIndexedTriBytes = 56;
}

bool Scene::loadObj(const std::string& fileName, bool tryLoadBinCache) {
	printf("* Loading '%s'...\n", fileName.c_str());
	for (int k = 0; k < 3; ++k) { _bbox[k] = FLT_MAX; _bbox[3 + k] = -FLT_MAX; }
	_vertices.clear(); _indexed.clear(); _triangles.clear();
	if (tryLoadBinCache) {
		std::string cache = fileName + ".bincache";
		std::ifstream probe(cache);
		if (probe.good()) {
			printf("\t- Binary cache found! Loading '%s'... ", cache.c_str());
			if (!loadBinObj(cache)) { printf("FAILED\n"); return false; }
			printf("OK!\n\t- Building tri vector... ");
			buildTriVector();
			printf("OK!\n\t- Loaded %zu triangles\n", getNRawTriangles());
			printf("\t- Bbox: [%.3f %.3f %.3f]  [%.3f %.3f %.3f]\n", _bbox[0], _bbox[1], _bbox[2], _bbox[3], _bbox[4], _bbox[5]);
			return true;
		}
	}
	printf("\t- Reading ASCII OBJ (reading lines)...");
	FILE* f = fopen(fileName.c_str(), "r");
	if (!f) { printf("ERROR: Scene:loadOBJ: Can't open file %s\n", fileName.c_str()); return false; }
	static char line[10000];
	size_t nLine = 0;
	while (fgets(line, sizeof line, f)) {
		++nLine;
		if (line[0] == 'v' && line[1] != 'n' && line[1] != 't') {
			float p[3];
			if (sscanf(line + 1, "%f %f %f", &p[0], &p[1], &p[2]) != 3) { printf("Read in line %zu a bad vertex format\n", nLine); continue; }
			for (int k = 0; k < 3; ++k) { if (p[k] < _bbox[k]) _bbox[k] = p[k]; if (p[k] > _bbox[3 + k]) _bbox[3 + k] = p[k]; _vertices.push_back(p[k]); }
		} else if (line[0] == 'f') {
			// tokens separated by blanks; the reference drops the last character of the last token ('\n')
			std::vector<std::string> tok;
			char* save = nullptr;
			for (char* t = strtok_r(line + 1, " ", &save); t; t = strtok_r(nullptr, " ", &save)) tok.push_back(t);
			if (tok.empty()) continue;
			tok.back().resize(tok.back().size() - 1);
			if (tok.size() < 3) continue;
			const int nv = (int)(_vertices.size() / 3);
			auto index_of = [&](const std::string& s) -> size_t {
				int iv = 0, in = 0;
				sscanf(s.c_str(), "%d//%d", &iv, &in);
				if (iv < 0) iv = nv + iv + 1;
				return (size_t)(iv - 1);
			};
			for (int i = 0; i < 3; ++i) _indexed.push_back(index_of(tok[i]));
			if (tok.size() > 4) for (int i = 0; i < 3; ++i) _indexed.push_back(index_of(tok[(i + 2) % 4]));   // scene.cpp:183-186
		}
	}
	fclose(f);
	printf(" OK!\n\t- Loaded %zu triangles in %.2fM obj lines\n", _indexed.size() / 3, nLine / 1e6);
	printf("\t- Bbox: [%.3f %.3f %.3f]  [%.3f %.3f %.3f]\n", _bbox[0], _bbox[1], _bbox[2], _bbox[3], _bbox[4], _bbox[5]);
	printf("\t- Saving '%s' binary cache... ", (fileName + ".bincache").c_str());
	saveBinObj(fileName + ".bincache");
	printf("OK!\n\t- Building tri vector... ");
	buildTriVector();
	printf("OK!\n");
	return true;
}

bool Scene::loadBinObj(const std::string& fileName) {
	std::ifstream in(fileName, std::ios::binary);
	BinObjHeader h;
	if (!in.read((char*)&h, sizeof h)) return false;
	memcpy(_bbox, h.bboxMin, 12);
	memcpy(_bbox + 3, h.bboxMax, 12);
	_vertices.resize(h.nVertices * 3);
	in.read((char*)_vertices.data(), (std::streamsize)(h.nVertices * 12));
	in.seekg((std::streamoff)(h.nNormals * 12 + h.nMaterials * kMaterialBytes), std::ios::cur);
	std::vector<uint64_t> rec(h.nIndexedTris * 7);
	in.read((char*)rec.data(), (std::streamsize)(h.nIndexedTris * kIndexedTriBytes));
	if (!in) return false;
	_indexed.resize(h.nIndexedTris * 3);
	for (size_t t = 0; t < h.nIndexedTris; ++t) for (int i = 0; i < 3; ++i) _indexed[t * 3 + i] = (size_t)rec[t * 7 + i];
	return true;
}

void Scene::saveBinObj(const std::string& fileName) const {
	BinObjHeader h;
	h.nVertices = _vertices.size() / 3; h.nNormals = 0; h.nMaterials = 1; h.nIndexedTris = _indexed.size() / 3;
	memcpy(h.bboxMin, _bbox, 12);
	memcpy(h.bboxMax, _bbox + 3, 12);
	std::ofstream out(fileName, std::ios::binary);
	out.write((const char*)&h, sizeof h);
	out.write((const char*)_vertices.data(), (std::streamsize)(_vertices.size() * 4));
	char mat[kMaterialBytes];
	memset(mat, 0, sizeof mat);
	snprintf(mat, 256, "Voxelator Default Mat");
	const float cols[9] = {0.9f, 0.9f, 0.9f, 0.2f, 0.2f, 0.2f, 0.1f, 0.1f, 0.1f};   // diffuse, spec, ambient (scene.hpp:32-39)
	memcpy(mat + 256, cols, sizeof cols);
	out.write(mat, sizeof mat);
	std::vector<uint64_t> rec(h.nIndexedTris * 7, 0);
	for (size_t t = 0; t < h.nIndexedTris; ++t) for (int i = 0; i < 3; ++i) rec[t * 7 + i] = _indexed[t * 3 + i];
	out.write((const char*)rec.data(), (std::streamsize)(rec.size() * 8));
}

void Scene::buildTriVector() {   // scene.cpp:394-416: triangles with an out-of-range index are skipped
	const size_t nv = _vertices.size() / 3;
	_triangles.clear();
	_triangles.reserve(_indexed.size() * 3);
	for (size_t t = 0; t + 2 < _indexed.size(); t += 3) {
		if (_indexed[t] >= nv || _indexed[t + 1] >= nv || _indexed[t + 2] >= nv) continue;
		for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) _triangles.push_back(_vertices[_indexed[t + i] * 3 + k]);
	}
	_vertices.clear(); _vertices.shrink_to_fit();
	_indexed.clear(); _indexed.shrink_to_fit();
}

}  // namespace svbhost
