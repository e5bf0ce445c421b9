// svbuilder -- drop-in for the reference's command line tool (src/svbuilder/main.cpp): same
// arguments, same output files next to the input (<base>_<L>.svdag/.ussvdag/.ssvdag/.esvdag, or
// <base>_<L>.svdag + <base>_<L>-multi.svdag with -c), same result block and ./stats.txt append.
// Everything between "load the mesh" and "encode" runs on the GPU through libsvb.so.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "geom_octree.hpp"

using namespace svbhost;

static void printUsage() {
	printf("\nUsage:\n"
	       "      svbuilder input_model.obj numLevels numBuildSteps [--cross-level-merging or -c] [<output.[svdag | ussvdag | ssvdag | esvdag]>]\n"
	       "Where:\n"
	       "      input_model.obj: a 3D model in ASCII OBJ format (an <obj>.bincache next to it is preferred)\n"
	       "      numLevels: levels of the octree to build (i.e. 10->1K^3, 13->8K^3...)\n"
	       "      numBuildSteps: 0 = one octree; > 0 = levels of the 'base octree' whose full children are built as\n"
	       "                     independent sub-octrees (device batches) and merged\n"
	       "      SVB_DEVICE=<ordinal> selects the GPU\n\n");
}

static std::string dir_of(const std::string& p) { size_t k = p.find_last_of('/'); return k == std::string::npos ? "." : p.substr(0, k); }
static std::string base_of(const std::string& p) {
	size_t k = p.find_last_of('/');
	std::string b = (k == std::string::npos) ? p : p.substr(k + 1);
	size_t d = b.find_last_of('.');
	return d == std::string::npos ? b : b.substr(0, d);
}

int main(int argc, char** argv) {
	printf("\n===============================================================================\n"
	       "==========   svbuilder (B200 / CUDA build of the SymVox builder path)   =========\n"
	       "===============================================================================\n");
	if (argc < 4) { printUsage(); exit(1); }
	std::string inputFile(argv[1]);
	int nLevels = atoi(argv[2]);
	int levelStep = atoi(argv[3]);
	printf(" MODEL: '%s'   [ %d levels, step %i ]   (%.0fK^3)\n", inputFile.c_str(), nLevels, levelStep, pow(2, nLevels) / 1024.f);
	printf("===============================================================================\n\n");
	auto t0 = std::chrono::steady_clock::now();

	bool multiLevel = false;
	for (int i = 4; i < argc; ++i) {
		std::string a = argv[i];
		if (a == "--cross-level-merging" || a == "-c") multiLevel = true;
		if (a == "--lossy" || a == "-l" || a == "--hidden-geometry" || a == "-h") {
			printf("Option '%s' (lossy / hidden-geometry DAGs) is outside this build's scope.\n", a.c_str());
			exit(1);
		}
	}
	printf("Lossy: %d, Cross-level: %d, Hidden geom: %d\n", 0, multiLevel, 0);

	Scene scene;
	const bool isObj = strstr(inputFile.c_str(), ".obj") || strstr(inputFile.c_str(), ".OBJ");
	const bool isInputDAG = !isObj && strstr(inputFile.c_str(), ".svdag");   // main.cpp:96-101
	if (!isObj && !isInputDAG) {
		printf("Can't read input file '%s'. Only supported ASCII Obj files.\n", inputFile.c_str());
		exit(1);
	}
	if (isObj && !scene.loadObj(inputFile, true)) exit(1);
	const char* devEnv = getenv("SVB_DEVICE");
	GeomOctree octree(&scene, devEnv ? atoi(devEnv) : 0);
	if (isInputDAG && !octree.loadSVDAG(inputFile)) exit(1);

	float mnF[3] = {0, 0, 0}, mxF[3] = {1, 1, 1};
	if (!isInputDAG) {
	scene.getBounds(mnF, mxF);
	double mnD[3] = {mnF[0], mnF[1], mnF[2]}, mxD[3] = {mxF[0], mxF[1], mxF[2]};   // main.cpp:150-155
	if (levelStep == 0) {
		octree.buildSVO(nLevels, mnD, mxD);
		octree.toDAG();
	} else {
		octree.buildDAG(nLevels, levelStep, mnD, mxD, true);
	}
	// main.cpp:179-193: header bbox rescale for very large / very small scenes
	float diag = std::sqrt((mxF[0] - mnF[0]) * (mxF[0] - mnF[0]) + (mxF[1] - mnF[1]) * (mxF[1] - mnF[1]) + (mxF[2] - mnF[2]) * (mxF[2] - mnF[2]));
	const float maxBboxSize = 100000.f;
	if (diag > maxBboxSize || diag < 0.1) {
		float nb[3] = {(mxF[0] - mnF[0]) / diag * maxBboxSize, (mxF[1] - mnF[1]) / diag * maxBboxSize, (mxF[2] - mnF[2]) / diag * maxBboxSize};
		float zero[3] = {0, 0, 0};
		printf("Normalizing bbox; too small/large: %f\n", diag);
		scene.setAABB(zero, nb);
		octree.resizeSceneBbox(zero, nb);
	}
	}   // !isInputDAG (main.cpp:147-201)
	octree.initChildLevels();

	std::string basePath = dir_of(inputFile) + "/" + base_of(inputFile) + "_" + std::to_string(nLevels);
	size_t szSvdag2 = 0, szSvdag = 0, szEsvdag = 0, szUssvdag = 0, szSsvdag = 0;
	printf("* Saving SVDAG '%s'... ", (basePath + ".svdag").c_str());
	octree.encodeToFile(0, basePath + ".svdag", &szSvdag2);   // main.cpp:220-222
	printf("OK!\n");
	szSvdag = szSvdag2;
	if (multiLevel) {
		octree.mergeAcrossAllLevels();
		printf("* Saving SVDAG '%s'... ", (basePath + "-multi.svdag").c_str());
		octree.encodeToFile(0, basePath + "-multi.svdag", &szSvdag);
		printf("OK!\n");
	} else {
		octree.encodeToFile(2, basePath + ".esvdag", &szEsvdag);   // SSVDAG encoder on the un-mirrored DAG (main.cpp:243-244)
		octree.toSDAG(false, false);
		octree.encodeToFile(1, basePath + ".ussvdag", &szUssvdag);
		octree.encodeToFile(2, basePath + ".ssvdag", &szSsvdag);
		printf("* Saved '%s'.{svdag,ussvdag,ssvdag,esvdag}\n", basePath.c_str());
	}
	for (int i = 4; i < argc; ++i) {   // trailing explicit output names (main.cpp:273-290)
		std::string o = argv[i];
		GeomOctree::State st = octree.getState();
		if (strstr(o.c_str(), ".ussvdag") || strstr(o.c_str(), ".USSVDAG")) { if (st == GeomOctree::S_SDAG) octree.encodeToFile(1, o); }
		else if (strstr(o.c_str(), ".ssvdag") || strstr(o.c_str(), ".SSVDAG")) { if (st == GeomOctree::S_SDAG) octree.encodeToFile(2, o); }
		else if (strstr(o.c_str(), ".svdag") || strstr(o.c_str(), ".SVDAG")) { if (st == GeomOctree::S_DAG) octree.encodeToFile(0, o); }
	}
	double totalS = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	GeomOctree::Stats s = octree.getStats();
	const size_t szHdr = 44, szHdrSS = 36 + 12;   // getDataSize() of the reference counts payload only; sizes below are file sizes minus headers
	(void)szHdr; (void)szHdrSS;
	printf("\n========= RESULTS '%s' [%d levels] (%.0fK^3) =========\n", inputFile.c_str(), nLevels, pow(2, nLevels) / 1024.f);
	printf("Voxels:     (%zu)\n", (size_t)s.nTotalVoxels);
	printf("SVO Nodes:  (%zu)\n", (size_t)s.nNodesSVO);
	printf("DAG Nodes:  (%zu)\n", (size_t)s.nNodesDAG);
	printf("SDAG Nodes: (%zu)\n", (size_t)s.nNodesSDAG);
	printf("Encoded SVDAG file   : %zu bytes\n", szSvdag);
	printf("Encoded ESVDAG file  : %zu bytes\n", szEsvdag);
	printf("Encoded USSVDAG file : %zu bytes\n", szUssvdag);
	printf("Encoded SSVDAG file  : %zu bytes\n", szSsvdag);
	printf("GPU build time       : %.2f ms (voxelize %.2f, reduce %.2f, rank %.2f)\n", s.msTotal, s.msVoxelize, s.msDedup, s.msFinalize);
	printf("SVDAG->SSVDAG time   : %.2f ms\n", s.msSdag);
	printf("Total time           : %.2f s\n", totalS);
	printf("===================================================================\n\n");
	FILE* f = fopen("stats.txt", "a");   // main.cpp:320-366
	if (f) {
		fprintf(f, "%s, %d, lossy: %d, cross-level: %d\n", base_of(inputFile).c_str(), nLevels, 0, multiLevel);
		fprintf(f, "#Voxels, %zu\n, SVDAG, ESVDAG, SSVDAG, SVO\n", (size_t)s.nTotalVoxels);
		fprintf(f, "#nodes, %zu, '', %zu, %zu\n", (size_t)s.nNodesDAG, (size_t)s.nNodesSDAG, (size_t)s.nNodesSVO);
		fprintf(f, "file bytes, %zu, %zu, %zu, %zu\n", szSvdag, szEsvdag, szSsvdag, (size_t)s.nNodesSVO);
		fprintf(f, "Construction times:\n,SVDAG, Total\ntime (ms), %zu, %zu\n", (size_t)s.msTotal, (size_t)(totalS * 1e3));
		if (multiLevel) fprintf(f, "Cross-level, nodes eliminated, svdag bytes, csvdag bytes\n, %zu, %zu, %zu\n", (size_t)s.nCrossLevelMerged, szSvdag2, szSvdag);
		fprintf(f, "\n\n");
		fclose(f);
	}
	return 0;
}
