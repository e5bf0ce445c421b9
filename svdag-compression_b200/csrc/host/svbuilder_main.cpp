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
#include <vector>

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
	       "      --devices 0,1,...  (or SVB_DEVICES=0,1,...) GPUs to build on: with more than one and numBuildSteps > 0 the\n"
	       "                     sub-octrees are spread over them and merged over NCCL; SVB_DEVICE=<ordinal> selects a single GPU\n\n");
}

// sl::human_readable_{quantity,size,duration}: SpaceLand is not vendored with the reference; these follow oracle/sl_shim,
// the stand-in the reference binary of this repository is built against, so both tools print the same text.
static std::string hr_quantity(double q) {
	char b[64];
	if (q < 1e3) snprintf(b, sizeof b, "%.0f", q);
	else if (q < 1e6) snprintf(b, sizeof b, "%.2f K", q / 1e3);
	else if (q < 1e9) snprintf(b, sizeof b, "%.2f M", q / 1e6);
	else snprintf(b, sizeof b, "%.2f G", q / 1e9);
	return b;
}
static std::string hr_size(double q) {
	char b[64];
	if (q < 1024.0) snprintf(b, sizeof b, "%.0f B", q);
	else if (q < 1048576.0) snprintf(b, sizeof b, "%.2f KB", q / 1024.0);
	else if (q < 1073741824.0) snprintf(b, sizeof b, "%.2f MB", q / 1048576.0);
	else snprintf(b, sizeof b, "%.2f GB", q / 1073741824.0);
	return b;
}
static std::string hr_duration(double s) {
	char b[64];
	if (s < 1e-3) snprintf(b, sizeof b, "%.0f us", s * 1e6);
	else if (s < 1.0) snprintf(b, sizeof b, "%.2f ms", s * 1e3);
	else if (s < 60.0) snprintf(b, sizeof b, "%.2f s", s);
	else snprintf(b, sizeof b, "%dm %.1fs", int(s / 60.0), s - 60.0 * int(s / 60.0));
	return b;
}
static std::vector<int> parse_devices(const char* s) {
	std::vector<int> v;
	while (s && *s) {
		char* e = nullptr;
		long d = strtol(s, &e, 10);
		if (e == s) break;
		v.push_back((int)d);
		s = (*e == ',') ? e + 1 : e;
	}
	return v;
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
	std::vector<int> devices = parse_devices(getenv("SVB_DEVICES"));
	for (int i = 4; i < argc; ++i) {
		std::string a = argv[i];
		if (a == "--cross-level-merging" || a == "-c") multiLevel = true;
		if (a == "--devices" && i + 1 < argc) devices = parse_devices(argv[++i]);
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
	if (devices.empty()) devices.push_back(devEnv ? atoi(devEnv) : 0);
	GeomOctree octree(&scene, devices);
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

	std::string baseName = base_of(inputFile);
	std::string basePath = dir_of(inputFile) + "/" + baseName + "_" + std::to_string(nLevels);
	// getDataSize() of the reference's encoder objects (main.cpp:220-271): payload bytes, headers and length prefixes excluded
	size_t dSvdag2 = 0, dSvdag = 0, dEsvdag = 0, dUssvdag = 0, dSsvdag = 0;
	octree.encodeToFile(SVB_FILE_SVDAG, basePath + ".svdag", nullptr, &dSvdag2);   // "Save base SVDAG ... in any case" (main.cpp:219-222)
	dSvdag = dSvdag2;
	if (multiLevel) {
		octree.mergeAcrossAllLevels();
		octree.encodeToFile(SVB_FILE_SVDAG, basePath + "-multi.svdag", nullptr, &dSvdag);
	} else {
		octree.encodeToFile(SVB_FILE_SSVDAG, basePath + ".esvdag", nullptr, &dEsvdag);   // SSVDAG encoder on the un-mirrored DAG (main.cpp:243-244)
		octree.toSDAG(false, false);
		octree.encodeToFile(SVB_FILE_USSVDAG, basePath + ".ussvdag", nullptr, &dUssvdag);
		octree.encodeToFile(SVB_FILE_SSVDAG, basePath + ".ssvdag", nullptr, &dSsvdag);
	}
	for (int i = 4; i < argc; ++i) {   // trailing explicit output names (main.cpp:273-290)
		std::string o = argv[i];
		if (o == "--devices") { ++i; continue; }
		GeomOctree::State st = octree.getState();
		if (strstr(o.c_str(), ".ussvdag") || strstr(o.c_str(), ".USSVDAG")) { if (st == GeomOctree::S_SDAG) octree.encodeToFile(SVB_FILE_USSVDAG, o); }
		else if (strstr(o.c_str(), ".ssvdag") || strstr(o.c_str(), ".SSVDAG")) { if (st == GeomOctree::S_SDAG) octree.encodeToFile(SVB_FILE_SSVDAG, o); }
		else if (strstr(o.c_str(), ".svdag") || strstr(o.c_str(), ".SVDAG")) { if (st == GeomOctree::S_DAG) octree.encodeToFile(SVB_FILE_SVDAG, o); }
	}
	double totalS = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	GeomOctree::Stats s = octree.getStats();
	const float nVox = float(octree.getNVoxels());
	// the reference's result block, line for line (main.cpp:300-317); the times are this build's (CUDA events / wall clock)
	printf("\n========= RESULTS '%s' [%d levels] (%.0fK^3) =========\n", inputFile.c_str(), nLevels, pow(2, nLevels) / 1024.f);
	printf("Voxels:     %s\t(%zu)\n", hr_quantity((double)s.nTotalVoxels).c_str(), (size_t)s.nTotalVoxels);
	printf("SVO Nodes:  %s\t(%zu)\n", hr_quantity((double)s.nNodesSVO).c_str(), (size_t)s.nNodesSVO);
	printf("DAG Nodes:  %s\t(%zu)\n", hr_quantity((double)s.nNodesDAG).c_str(), (size_t)s.nNodesDAG);
	printf("SDAG Nodes: %s\t(%zu)\n", hr_quantity((double)s.nNodesSDAG).c_str(), (size_t)s.nNodesSDAG);
	printf("Pointerless SVO    : %s\t(%.3f bits/vox)\n", hr_size((double)s.nNodesSVO).c_str(), (8 * s.nNodesSVO) / nVox);
	printf("Encoded SVDAG      : %s\t(%.3f bits/vox)\n", hr_size((double)dSvdag).c_str(), (8 * dSvdag) / nVox);
	printf("Encoded ESVDAG     : %s\t(%.3f bits/vox)\n", hr_size((double)dEsvdag).c_str(), (8 * dEsvdag) / nVox);
	printf("Encoded USSVDAG    : %s\t(%.3f bits/vox)\n", hr_size((double)dUssvdag).c_str(), (8 * dUssvdag) / nVox);
	printf("Encoded SSVDAG     : %s\t(%.3f bits/vox)\n", hr_size((double)dSsvdag).c_str(), (8 * dSsvdag) / nVox);
	printf("SSVDAG / DAG       : %.1f %%\n", 100.f * dSsvdag / (float)dSvdag);
	printf("SSVDAG / SVO       : %.1f %%\n", 100.f * dSsvdag / (float)s.nNodesSVO);
	printf("SVO->SVDAG time    : %s\n", hr_duration(s.msTotal * 1e-3).c_str());
	printf("SVDAG->LSVDAG time : %s\n", hr_duration(0).c_str());
	printf("SVDAG->SSVDAG time : %s\n", hr_duration(s.msSdag * 1e-3).c_str());
	printf("Total time         : %s\n", hr_duration(totalS).c_str());
	printf("===================================================================\n\n");
	printf("GPU build            : %.2f ms on %zu GPU(s) (voxelize %.2f, reduce %.2f, rank %.2f", s.msTotal, octree.nDevices(), s.msVoxelize, s.msDedup, s.msFinalize);
	if (octree.nDevices() > 1) printf("; triangle upload + all-gather %.2f, level exchange + finish %.2f", octree.lastUploadMs(), octree.lastExchangeMs());
	printf(")\n\n");
	// ... and its block for the console and ./stats.txt (main.cpp:320-366)
	for (int pass = 0; pass < 2; ++pass) {
		FILE* f = pass == 0 ? stdout : fopen("stats.txt", "a");
		if (!f) break;
		fprintf(f, "%s, %d, lossy: %d, cross-level: %d\n", baseName.c_str(), nLevels, 0, (int)multiLevel);
		fprintf(f, "#Voxels, %zu\n", (size_t)s.nTotalVoxels);
		fprintf(f, ", SVDAG, ESVDAG, SSVDAG, SVO\n");
		fprintf(f, "#nodes, %zu, '', %zu, %zu\n", (size_t)s.nNodesDAG, (size_t)s.nNodesSDAG, (size_t)s.nNodesSVO);
		fprintf(f, "memory (bytes), %zu, %zu, %zu, %zu\n", dSvdag, dEsvdag, dSsvdag, (size_t)s.nNodesSVO);
		if (multiLevel) {
			fprintf(f, "Construction times:\n");
			fprintf(f, ",SVDAG, Total, Cross-level\n");
			fprintf(f, "time (ms), %zu, %zu, %zu\n", (size_t)s.msTotal, (size_t)(totalS * 1e3), (size_t)s.msCrossMerge);
			fprintf(f, "Cross-level, nodes eliminated, svdag mem, csvdag mem\n");
			fprintf(f, ", %zu, %zu, %zu\n", (size_t)s.nCrossLevelMerged, dSvdag2, dSvdag);
		} else {
			fprintf(f, "Construction times:\n");
			fprintf(f, ",SVDAG, Total\n");
			fprintf(f, "time (ms), %zu, %zu\n", (size_t)s.msTotal, (size_t)(totalS * 1e3));
		}
		fprintf(f, "\n\n");
		if (pass == 1) fclose(f);
	}
	return 0;
}
