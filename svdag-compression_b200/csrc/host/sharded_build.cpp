#include "sharded_build.hpp"

#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>

namespace svbhost {

namespace {

// reusable barrier for the device threads (C++14: no std::barrier)
class Barrier {
public:
	explicit Barrier(int n) : _n(n) {}
	void wait() {
		std::unique_lock<std::mutex> lk(_m);
		const unsigned gen = _gen;
		if (++_count == _n) { _count = 0; ++_gen; _cv.notify_all(); }
		else _cv.wait(lk, [&] { return gen != _gen; });
	}
private:
	std::mutex _m;
	std::condition_variable _cv;
	int _n, _count = 0;
	unsigned _gen = 0;
};

struct Shared {
	int world = 0;
	std::vector<ncclComm_t> comm;
	std::vector<std::vector<uint64_t>> counts;     // [rank][level index]: records this rank exports per level
	std::vector<std::vector<uint32_t>> recBytes;
	std::vector<std::vector<uint64_t>> counters;   // [rank][5]
	std::vector<std::string> error;                // per rank
	std::vector<int> collision;                    // per rank: svb_shard_finish reported a tag collision of the level merge
	std::vector<double> msUpload, msExchange;
	bool failed() const { for (auto& e : error) if (!e.empty()) return true; return false; }
};

struct DevMem {   // plain cudaMalloc'ed scratch owned by one device thread
	std::vector<void*> ptrs;
	void* get(size_t bytes) {
		void* p = nullptr;
		if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) { cudaGetLastError(); return nullptr; }
		ptrs.push_back(p);
		return p;
	}
	~DevMem() { for (void* p : ptrs) cudaFree(p); }
};

#define RANK_FAIL(msg) do { S.error[rank] = (msg); } while (0)
#define CU(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess && S.error[rank].empty()) RANK_FAIL(std::string(#expr) + ": " + cudaGetErrorString(_e)); } while (0)
#define NC(expr) do { ncclResult_t _r = (expr); if (_r != ncclSuccess && S.error[rank].empty()) RANK_FAIL(std::string(#expr) + ": " + ncclGetErrorString(_r)); } while (0)
#define SV(expr) do { int _rc = (expr); if (_rc != SVB_OK && S.error[rank].empty()) RANK_FAIL(std::string(#expr) + ": " + svb_last_error(c)); } while (0)

// Every collective is entered by every rank even after a local failure (a rank that bailed out would hang the others);
// failures are recorded and reported once all threads have joined.
void rank_main(int rank, Shared& S, Barrier& bar, int device, svb_ctx* c, const float* tris, uint64_t ntris, unsigned levels, unsigned step,
               const double* bmin, const double* bmax, svb_stats* out, float** dTrisOut) {
	typedef std::chrono::steady_clock clk;
	const int world = S.world;
	CU(cudaSetDevice(device));
	cudaStream_t s = (cudaStream_t)svb_stream(c);
	ncclComm_t comm = S.comm[rank];
	DevMem mem;
	// ---- triangles: 1/world over this device's PCIe link, the rest over NVLink
	auto t0 = clk::now();
	const uint64_t nfloat = ntris * 9;
	const uint64_t chunk = ((nfloat + world - 1) / world + 3) / 4 * 4;
	float* dTris = nullptr;
	CU(cudaMalloc((void**)&dTris, (chunk * world + 16) * sizeof(float)));   // owned by the caller (the octree borrows it)
	*dTrisOut = dTris;
	bar.wait();                 // a device that could not allocate must not leave the others waiting inside the all-gather
	if (S.failed()) return;
	const uint64_t lo = std::min<uint64_t>((uint64_t)rank * chunk, nfloat), hi = std::min<uint64_t>((uint64_t)(rank + 1) * chunk, nfloat);
	if (dTris && hi > lo) CU(cudaMemcpyAsync(dTris + (uint64_t)rank * chunk, tris + lo, (hi - lo) * sizeof(float), cudaMemcpyHostToDevice, s));
	if (dTris) NC(ncclAllGather(dTris + (uint64_t)rank * chunk, dTris, chunk, ncclFloat, comm, s));
	CU(cudaStreamSynchronize(s));
	S.msUpload[rank] = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
	if (dTris) SV(svb_set_triangles_device(c, dTris, ntris));
	// ---- local phase
	if (S.error[rank].empty()) SV(svb_shard_build(c, levels, step, bmin, bmax, (uint32_t)rank, (uint32_t)world));
	uint32_t first = 0, last = 0;
	uint64_t nTiles = 0;
	S.counters[rank].assign(5, 0);
	const bool built = S.error[rank].empty();
	if (built) SV(svb_shard_info(c, &first, &last, &nTiles, S.counters[rank].data()));
	// level indices are the same everywhere even if this rank failed: derive them from the arguments
	first = step + 1; last = levels - 1;
	const int nLev = (int)(last - first + 1);
	S.counts[rank].assign(nLev, 0);
	S.recBytes[rank].assign(nLev, 16);
	for (int k = 0; k < nLev && built; ++k) SV(svb_shard_level_count(c, last - k, &S.counts[rank][k], &S.recBytes[rank][k]));
	bar.wait();   // the one host rendezvous: every rank's counts and counters are now visible to all
	if (S.failed()) return;   // decided identically by every thread (all wrote before the barrier)
	auto t1 = clk::now();
	// ---- bottom-up exchange, stream-ordered: export -> all-gather -> import, no host synchronisation in between
	std::vector<uint64_t> cnt(world);
	for (int k = 0; k < nLev; ++k) {
		const uint32_t g = last - k;
		uint64_t mx = 0;
		for (int r = 0; r < world; ++r) { cnt[r] = S.counts[r][k]; mx = std::max(mx, cnt[r]); }
		const uint32_t rec = S.recBytes[rank][k];
		const uint64_t stride = std::max<uint64_t>(16, (mx * rec + 15) / 16 * 16);
		char* all = (char*)mem.get(stride * world);
		if (!all) { RANK_FAIL("cudaMalloc(exchange buffer) failed"); all = nullptr; }
		char* mine = all ? all + (uint64_t)rank * stride : nullptr;
		if (mine) SV(svb_shard_export_level(c, g, mine));
		if (all) NC(ncclAllGather(mine, all, stride, ncclChar, comm, s));   // in place: this rank's records already sit in its row
		if (all) SV(svb_shard_import_level(c, g, all, cnt.data(), stride));
	}
	const uint64_t nt = std::max<uint64_t>(nTiles, 1);
	uint32_t* roots = (uint32_t*)mem.get(nt * 4 * world);
	if (!roots) RANK_FAIL("cudaMalloc(root buffer) failed");
	if (roots) {
		SV(svb_shard_export_roots(c, roots + (uint64_t)rank * nt));
		NC(ncclAllGather(roots + (uint64_t)rank * nt, roots, nt, ncclUint32, comm, s));
		SV(svb_shard_import_roots(c, roots));
	}
	uint64_t totals[5] = {0, 0, 0, 0, 0};
	for (int r = 0; r < world; ++r) for (int k = 0; k < 5; ++k) totals[k] += S.counters[r][k];
	svb_stats st;
	memset(&st, 0, sizeof(st));
	if (S.error[rank].empty()) {   // synchronises the stream: the scratch may go
		const int rc = svb_shard_finish(c, totals, &st);
		if (rc == SVB_ECOLLISION) S.collision[rank] = 1;
		if (rc != SVB_OK) RANK_FAIL(std::string("svb_shard_finish: ") + svb_last_error(c));
	} else CU(cudaStreamSynchronize(s));
	S.msExchange[rank] = std::chrono::duration<double, std::milli>(clk::now() - t1).count();
	if (rank == 0 && out) *out = st;
}

}  // namespace

static bool build_once(const std::vector<int>& devices, const std::vector<svb_ctx*>& ctx, const float* tris, uint64_t ntris,
                       unsigned levels, unsigned step, const double bmin[3], const double bmax[3], svb_stats* out, std::string* err,
                       double* msUpload, double* msExchange, bool* collided);

bool build_dag_sharded(const std::vector<int>& devices, const std::vector<svb_ctx*>& ctx, const float* tris, uint64_t ntris,
                       unsigned levels, unsigned step, const double bmin[3], const double bmax[3], svb_stats* out, std::string* err,
                       double* msUpload, double* msExchange) {
	const int world = (int)devices.size();
	if (world < 2 || ctx.size() != devices.size()) { if (err) *err = "build_dag_sharded needs >= 2 devices and one context per device"; return false; }
	if (step == 0 || step + 1 >= levels) { if (err) *err = "a sharded build needs 0 < step and step + 1 < levels (sub-octrees are the unit of distribution)"; return false; }
	// A 64-bit tag collision in the level merge is seen by every rank alike (identical records, identical seed): all of them
	// repeat the build under the next merge seed (include/svb.h: svb_set_merge_seed)
	for (uint64_t seed = 0;; ++seed) {
		for (svb_ctx* c : ctx) svb_set_merge_seed(c, seed);
		bool collided = false;
		const bool ok = build_once(devices, ctx, tris, ntris, levels, step, bmin, bmax, out, err, msUpload, msExchange, &collided);
		if (ok || !collided || seed >= 3) return ok;
	}
}

static bool build_once(const std::vector<int>& devices, const std::vector<svb_ctx*>& ctx, const float* tris, uint64_t ntris,
                       unsigned levels, unsigned step, const double bmin[3], const double bmax[3], svb_stats* out, std::string* err,
                       double* msUpload, double* msExchange, bool* collided) {
	const int world = (int)devices.size();
	Shared S;
	S.world = world;
	S.comm.resize(world);
	S.counts.resize(world); S.recBytes.resize(world); S.counters.resize(world); S.error.resize(world); S.collision.assign(world, 0);
	S.msUpload.assign(world, 0); S.msExchange.assign(world, 0);
	ncclResult_t nr = ncclCommInitAll(S.comm.data(), world, devices.data());
	if (nr != ncclSuccess) { if (err) *err = std::string("ncclCommInitAll: ") + ncclGetErrorString(nr); return false; }
	Barrier bar(world);
	std::vector<float*> dTris(world, nullptr);
	std::vector<std::thread> th;
	for (int r = 0; r < world; ++r)
		th.emplace_back(rank_main, r, std::ref(S), std::ref(bar), devices[r], ctx[r], tris, ntris, levels, step, bmin, bmax, out, &dTris[r]);
	for (auto& t : th) t.join();
	for (int r = 0; r < world; ++r) ncclCommDestroy(S.comm[r]);
	// the soup is only needed during the build (the octree borrowed it)
	for (int r = 0; r < world; ++r) if (dTris[r]) { cudaSetDevice(devices[r]); svb_set_triangles_device(ctx[r], nullptr, 0); cudaFree(dTris[r]); }
	if (msUpload) *msUpload = *std::max_element(S.msUpload.begin(), S.msUpload.end());
	if (msExchange) *msExchange = *std::max_element(S.msExchange.begin(), S.msExchange.end());
	if (collided) { *collided = true; for (int r = 0; r < world; ++r) if (!S.collision[r]) *collided = false; }
	for (int r = 0; r < world; ++r)
		if (!S.error[r].empty()) { if (err) *err = "device " + std::to_string(devices[r]) + ": " + S.error[r]; return false; }
	return true;
}

}  // namespace svbhost
