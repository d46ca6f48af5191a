// jp_bwt_api.cu -- the C-ABI of include/jp_bwt.h: context pool (device, stream, workspace), block
// sharding over devices, host<->device staging, error and stats plumbing. No kernels live here.
#include "bwt_internal.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <thread>
#include <vector>

namespace jp {

static thread_local char          t_detail[512] = "";
static thread_local jp_bwt_stats  t_stats = {};

void set_error_detail(const char* fmt, ...)
{
	va_list ap; va_start(ap, fmt);
	vsnprintf(t_detail, sizeof(t_detail), fmt, ap);
	va_end(ap);
}

int map_dev_err(int de)
{
	switch (de) {
	case DE_NONE: return JP_OK;
	case DE_BAD_INDEX: set_error_detail("stored primary index outside [1, nlen] or duplicated"); return JP_ERR_BAD_INDEX;
	case DE_CHAIN_LEN: set_error_detail("a decode unit's chain is not nlen/120 long"); return JP_ERR_CORRUPT;
	case DE_CHAIN_RANGE: set_error_detail("a sub-chain leaves its block"); return JP_ERR_CORRUPT;
	case DE_RANK_LOOP: set_error_detail("sub-chain ranking found a cycle"); return JP_ERR_CORRUPT;
	case DE_FWD_RANGE: set_error_detail("forward: rank lookup beyond the end of the block"); return JP_ERR_INTERNAL;
	case DE_FWD_ROUNDS: set_error_detail("forward: prefix doubling did not converge"); return JP_ERR_INTERNAL;
	case DE_STREAM_OVERFLOW: set_error_detail("inverse: stream overflow escaped the two-pass rerun"); return JP_ERR_INTERNAL;
	default: set_error_detail("device error flag %d", de); return JP_ERR_INTERNAL;
	}
}

// ---- device memory accounting ---------------------------------------------------------------------------
// Workspaces are per context and grow-only, so several large blocks in flight on one device (forward needs ~45N)
// can exhaust it. Every allocation goes through dev_alloc: it honours the optional JP_BWT_DEVICE_MEM_LIMIT (bytes
// the pool may hold per device) and turns a refused cudaMalloc into JP_ERR_OOM; the callers then make room
// (relieve_memory_pressure) and retry instead of failing the block.
static std::atomic<long long> g_dev_bytes[64];
static long long mem_limit()
{
	static long long lim = -2;
	if (lim == -2) { const char* e = getenv("JP_BWT_DEVICE_MEM_LIMIT"); lim = e ? atoll(e) : -1; }
	return lim;
}
static int dev_alloc(int device, void** p, size_t bytes)
{
	const long long lim = mem_limit();
	if (lim >= 0 && g_dev_bytes[device & 63].load() + (long long)bytes > lim) {
		set_error_detail("device %d: %zu more bytes would exceed JP_BWT_DEVICE_MEM_LIMIT", device, bytes);
		return JP_ERR_OOM;
	}
	const cudaError_t e = cudaMalloc(p, bytes);
	if (e != cudaSuccess) {
		cudaGetLastError();
		*p = nullptr;
		set_error_detail("cudaMalloc(%zu) on device %d -> %s", bytes, device, cudaGetErrorString(e));
		return e == cudaErrorMemoryAllocation ? JP_ERR_OOM : JP_ERR_CUDA;
	}
	g_dev_bytes[device & 63] += (long long)bytes;
	return JP_OK;
}
static void dev_free(int device, void* p, size_t bytes)
{
	if (!p) return;
	cudaFree(p);
	g_dev_bytes[device & 63] -= (long long)bytes;
}

int arena_reserve(Ctx& c, size_t total)
{
	if (c.arena.cap - c.arena.off >= total) return JP_OK;
	if (c.arena.off != 0) { set_error_detail("arena: %zu more bytes wanted with %zu live", total, c.arena.off); return JP_ERR_INTERNAL; }
	if (c.arena.base) { dev_free(c.device, c.arena.base, c.arena.cap); c.arena.base = nullptr; c.arena.cap = 0; }
	const size_t want = (total + (64u << 20) - 1) & ~(size_t)((64u << 20) - 1);
	JP_TRY(dev_alloc(c.device, (void**)&c.arena.base, want));
	c.arena.cap = want;
	return JP_OK;
}

int arena2_reserve(Ctx& c, size_t total, size_t keep)
{
	if (total > c.arena2.high) c.arena2.high = total;      // bytes of it this call uses (reported with the workspace)
	if (c.arena2.cap >= total) return JP_OK;
	const size_t want = (total + (16u << 20) - 1) & ~(size_t)((16u << 20) - 1);
	u8* fresh = nullptr;
	if (keep == 0 && c.arena2.base) { dev_free(c.device, c.arena2.base, c.arena2.cap); c.arena2.base = nullptr; c.arena2.cap = 0; }
	JP_TRY(dev_alloc(c.device, (void**)&fresh, want));
	if (c.arena2.base) {
		const cudaError_t e = cudaMemcpy(fresh, c.arena2.base, keep, cudaMemcpyDeviceToDevice);   // (synchronous: every stream of the call is idle here or ordered before it)
		dev_free(c.device, c.arena2.base, c.arena2.cap);
		if (e != cudaSuccess) { dev_free(c.device, fresh, want); c.arena2.base = nullptr; c.arena2.cap = 0; set_error_detail("arena2 carry-over: %s", cudaGetErrorString(e)); return JP_ERR_CUDA; }
	}
	c.arena2.base = fresh;
	c.arena2.cap = want;
	return JP_OK;
}

// ---- optional stage trace (JP_BWT_TRACE=1): per-direction totals printed to stderr at exit ---------------
// The reference's own progress line reports CPU time (clock(), jampack.cpp:202-229); this gives the wall-clock
// view of just this stage when it runs inside the reference's pipeline (BASELINE.json configs[3]).
struct Trace {
	std::atomic<long long> calls{0}, bytes{0}, wall_us{0}, dev_us{0};
	std::atomic<long long> first_us{-1}, last_us{0};
};
static Trace g_trace[2];
static std::atomic<int> g_trace_on{-1};
static bool g_trace_calls = false;             // JP_BWT_TRACE=2: one line per call as well
static long long now_us() { return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static std::atomic<long long>& hostreg_counter(int which);
static void trace_report()
{
	static const char* names[2] = {"forward", "inverse"};
	fprintf(stderr, "[jp_bwt trace] host blocks page-locked in the background after their first call: %lld (%.1f ms in cudaHostRegister)\n", hostreg_counter(0).load(), hostreg_counter(1).load() / 1e3);
	for (int d = 0; d < 2; d++) {
		const long long calls = g_trace[d].calls.load();
		if (!calls) continue;
		const double mb = g_trace[d].bytes.load() / 1e6, wall = g_trace[d].wall_us.load() / 1e6, dev = g_trace[d].dev_us.load() / 1e6;
		const double span = (g_trace[d].last_us.load() - g_trace[d].first_us.load()) / 1e6;
		fprintf(stderr, "[jp_bwt trace] %s: calls=%lld MB=%.1f sum_call_wall_s=%.4f sum_device_s=%.4f span_s=%.4f "
		        "MBps_per_call=%.1f MBps_over_span=%.1f devices=%d\n", names[d], calls, mb, wall, dev, span,
		        wall > 0 ? mb / wall : 0.0, span > 0 ? mb / span : 0.0, jp_bwt_device_count());
	}
}
static bool trace_enabled()
{
	int v = g_trace_on.load();
	if (v < 0) {
		const char* e = getenv("JP_BWT_TRACE");
		v = (e && *e && *e != '0') ? 1 : 0;
		if (v && *e == '2') g_trace_calls = true;
		int expect = -1;
		if (g_trace_on.compare_exchange_strong(expect, v) && v) atexit(trace_report);
	}
	return g_trace_on.load() == 1;
}
static void trace_add(int direction, long long bytes, long long t0, long long t1, float dev_ms)
{
	Trace& t = g_trace[direction];
	t.calls++; t.bytes += bytes; t.wall_us += (t1 - t0); t.dev_us += (long long)(dev_ms * 1e3f);
	long long expect = -1;
	t.first_us.compare_exchange_strong(expect, t0);
	long long prev = t.last_us.load();
	while (prev < t1 && !t.last_us.compare_exchange_weak(prev, t1)) {}
}

// ---- context pool -----------------------------------------------------------------------------------
// blocks in flight per device; more hides more latency but multiplies the workspaces (JP_BWT_MAX_CTX to override)
static int max_ctx_per_device()
{
	static int v = 0;
	if (v == 0) { const char* e = getenv("JP_BWT_MAX_CTX"); v = e ? atoi(e) : 4; if (v < 1) v = 1; if (v > 32) v = 32; }
	return v;
}
#define MAX_CTX_PER_DEVICE max_ctx_per_device()

struct Pool {
	std::mutex mu;
	std::condition_variable cv;
	std::vector<int> devices;
	bool devices_ready = false;
	std::vector<std::unique_ptr<Ctx>> ctxs;
	std::vector<int> pending;          // contexts being created, per visible device (creation runs outside the lock)
	std::vector<int> state;            // per visible device: 0 cold, 1 primary context coming up, 2 usable
	unsigned rr = 0;
};
static Pool g_pool;
static bool g_shutting_down = false;      // (under g_pool.mu)

static int visible_devices()
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

static int init_devices_locked()
{
	if (g_pool.devices_ready) return g_pool.devices.empty() ? JP_ERR_NO_DEVICE : JP_OK;
	const int n = visible_devices();
	g_pool.devices.clear();
	if (const char* e = getenv("JP_BWT_DEVICES")) {
		const char* p = e;
		while (*p) {
			char* end; long v = strtol(p, &end, 10);
			if (end == p) break;
			if (v >= 0 && v < n) g_pool.devices.push_back((int)v);
			p = (*end == ',') ? end + 1 : end;
		}
	}
	if (g_pool.devices.empty()) for (int i = 0; i < n; i++) g_pool.devices.push_back(i);
	g_pool.devices_ready = true;
	// (a reset while calls are in flight keeps the counts of contexts being created and the devices already up)
	if (g_pool.pending.size() < (size_t)(n > 0 ? n : 1)) g_pool.pending.resize((size_t)(n > 0 ? n : 1), 0);
	if (g_pool.state.size() < (size_t)(n > 0 ? n : 1)) g_pool.state.resize((size_t)(n > 0 ? n : 1), 0);
	if (g_pool.devices.empty()) { set_error_detail("no CUDA device visible; this stage has no CPU path"); return JP_ERR_NO_DEVICE; }
	// Only the first configured device is brought up here. A primary context costs about a second on these parts
	// and the driver creates them one after the other (measured: first batch of a 4-GPU run waited 3-5 s), so the
	// other devices come up in the background the first time the usable ones are saturated (warm_next_device).
	if (g_pool.state[g_pool.devices[0]] == 0) g_pool.state[g_pool.devices[0]] = 2;
	return JP_OK;
}

static int create_ctx(int device, Ctx** out)
{
	std::unique_ptr<Ctx> c(new Ctx());
	c->device = device;
	JP_CUDA(cudaSetDevice(device));
	if (const char* e = getenv("JP_BWT_L2_FETCH")) {
		// Experiment switch only. cudaLimitMaxL2FetchGranularity = 32/64/128 was measured to change nothing on B200:
		// a random 4-byte gather costs ~2.7 DRAM sectors whatever the hint (profiles/README.md).
		const long v = atol(e);
		if ((v == 32 || v == 64 || v == 128) && cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)v) != cudaSuccess) cudaGetLastError();
	}
	JP_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
	JP_CUDA(cudaMallocHost(&c->h_small, 256 * sizeof(int)));
	for (auto& e : c->ev) JP_CUDA(cudaEventCreate(&e));
	JP_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
	c->busy = true;
	*out = c.get();
	std::lock_guard<std::mutex> lk(g_pool.mu);
	g_pool.ctxs.push_back(std::move(c));
	return JP_OK;
}

// Starts bringing up one more configured device (detached; joins the pool when its primary context exists).
static void warm_next_device_locked()
{
	for (int d : g_pool.devices) if (g_pool.state[d] == 1) return;          // one at a time: the driver serialises them anyway
	for (int d : g_pool.devices) {
		if (g_pool.state[d] != 0) continue;
		g_pool.state[d] = 1;
		static bool hooked = false;
		if (!hooked) {                                           // never tear the runtime down under a device that is coming up
			hooked = true;
			atexit([] {
				for (int spin = 0; spin < 10000; spin++) {
					{ std::lock_guard<std::mutex> lk(g_pool.mu); bool busy = false; for (int st : g_pool.state) busy |= (st == 1); if (!busy) return; }
					std::this_thread::sleep_for(std::chrono::milliseconds(1));
				}
			});
		}
		std::thread([d] {
			const bool ok = cudaSetDevice(d) == cudaSuccess && cudaFree(0) == cudaSuccess;
			{ std::lock_guard<std::mutex> lk(g_pool.mu); g_pool.state[d] = ok ? 2 : 3; }   // 3: unusable, never retried
			g_pool.cv.notify_all();
		}).detach();
		return;
	}
}

// device < 0: the configured device with the fewest blocks in flight (ties go round-robin) -- whole blocks are
// the unit of sharding. A new context is created outside the pool lock, so first calls on different devices do
// not queue behind each other.
static int acquire(int device, Ctx** out)
{
	std::unique_lock<std::mutex> lk(g_pool.mu);
	JP_TRY(init_devices_locked());
	const bool any = device < 0;
	bool waited = false;
	if (!any && device >= (int)g_pool.pending.size()) { set_error_detail("device %d not visible", device); return JP_ERR_NO_DEVICE; }
	for (;;) {
		if (any) {
			int best = -1, best_load = 1 << 30;
			const size_t nd = g_pool.devices.size();
			for (size_t k = 0; k < nd; k++) {
				const int d = g_pool.devices[(g_pool.rr + k) % nd];
				if (g_pool.state[d] != 2 && g_pool.state[d] != 4) continue;
				int load = g_pool.pending[d];
				for (auto& c : g_pool.ctxs) if (c->device == d && c->busy) load++;
				if (load < best_load) { best_load = load; best = d; }
			}
			if (best < 0) {
				// no configured device is usable yet: bring one up (or wait for the one on its way); none left -> fail
				bool coming = false, cold = false;
				for (int d : g_pool.devices) { coming |= g_pool.state[d] == 1; cold |= g_pool.state[d] == 0; }
				if (!coming && !cold) { set_error_detail("no configured CUDA device could be brought up"); return JP_ERR_NO_DEVICE; }
				if (!coming) warm_next_device_locked();
				g_pool.cv.wait_for(lk, std::chrono::milliseconds(50));
				continue;
			}
			device = best;
			if (best_load < MAX_CTX_PER_DEVICE) g_pool.rr++;
			else if (waited) warm_next_device_locked();          // every usable device has stayed full: widen the pool meanwhile
		} else if (g_pool.state[device] != 2 && g_pool.state[device] != 4) {   // an explicitly named device is brought up on the spot
			g_pool.state[device] = 2;
		}
		int have = g_pool.pending[device];
		for (auto& c : g_pool.ctxs) {
			if (c->device != device) continue;
			have++;
			if (!c->busy) { c->busy = true; *out = c.get(); lk.unlock(); JP_CUDA(cudaSetDevice(device)); return JP_OK; }
		}
		if (have < MAX_CTX_PER_DEVICE) {
			g_pool.pending[device]++;
			lk.unlock();
			const int rc = create_ctx(device, out);
			lk.lock();
			g_pool.pending[device]--;
			if (rc != JP_OK && rc != JP_ERR_OOM) g_pool.state[device] = 3;   // the device itself failed: the 'any' path stops choosing it
			lk.unlock();
			if (rc != JP_OK) g_pool.cv.notify_all();
			return rc;
		}
		// short blocks clear a full pool in milliseconds; only a pool that stays full is worth another device's start-up
		waited = g_pool.cv.wait_for(lk, std::chrono::milliseconds(50)) == std::cv_status::timeout || waited;
	}
}

static void release(Ctx* c)
{
	{ std::lock_guard<std::mutex> lk(g_pool.mu); c->busy = false; }
	g_pool.cv.notify_all();
}

struct CtxGuard {
	Ctx* c = nullptr;
	~CtxGuard() { if (c) release(c); }
};

static int ensure_io(Ctx& c, size_t bytes)
{
	bytes = Arena::align(bytes + 64);
	if (c.d_io_cap >= bytes) return JP_OK;
	dev_free(c.device, c.d_in, c.d_io_cap); c.d_in = nullptr;
	dev_free(c.device, c.d_out, c.d_io_cap); c.d_out = nullptr;
	c.d_io_cap = 0;
	JP_TRY(dev_alloc(c.device, (void**)&c.d_in, bytes));
	if (dev_alloc(c.device, (void**)&c.d_out, bytes) != JP_OK) { dev_free(c.device, c.d_in, bytes); c.d_in = nullptr; return JP_ERR_OOM; }
	c.d_io_cap = bytes;
	return JP_OK;
}

static void drop_memory(Ctx& c)
{
	dev_free(c.device, c.arena.base, c.arena.cap); c.arena.base = nullptr; c.arena.cap = 0; c.arena.off = 0;
	dev_free(c.device, c.arena2.base, c.arena2.cap); c.arena2.base = nullptr; c.arena2.cap = 0;
	dev_free(c.device, c.d_in, c.d_io_cap); c.d_in = nullptr;
	dev_free(c.device, c.d_out, c.d_io_cap); c.d_out = nullptr;
	c.d_io_cap = 0;
}

// A call on `self` ran out of device memory. Take the workspaces of idle contexts of the same device away; if
// there are none, wait for a busy one to finish. false: nothing left to wait for -- the block really does not fit.
static bool relieve_memory_pressure(Ctx& self)
{
	std::unique_lock<std::mutex> lk(g_pool.mu);
	std::vector<Ctx*> idle;
	bool others_busy = false;
	for (auto& c : g_pool.ctxs) {
		if (c.get() == &self || c->device != self.device) continue;
		if (c->busy) others_busy = true;
		else if (c->arena.base || c->arena2.base || c->d_in) { c->busy = true; idle.push_back(c.get()); }
	}
	if (!idle.empty()) {
		lk.unlock();
		for (Ctx* c : idle) drop_memory(*c);
		lk.lock();
		for (Ctx* c : idle) c->busy = false;
		lk.unlock();
		g_pool.cv.notify_all();
		return true;
	}
	if (!others_busy && g_pool.pending[self.device] == 0) return false;
	g_pool.cv.wait_for(lk, std::chrono::milliseconds(200));
	return true;
}

// ---- caller blocks: page-lock them once ---------------------------------------------
// The reference allocates its two ping-pong blocks once per Jampack instance (pageable calloc/realloc, jampack.cpp:74-76,
// :157-159) and hands the same pointers to the stage for every block of the run. A pageable block goes through the
// driver's bounce buffers (measured in the reference pipeline: 6-21 ms per 64 MiB copy). Page-locking it
// (cudaHostRegister) turns every later copy into direct DMA, but costs 10-40 ms itself and stalls other CUDA calls while
// it runs -- measured inside the calls of a one-batch run it made the run SLOWER. So the first call on a block copies
// it as it is, and a background thread page-locks it afterwards, while the caller is busy with its other stages; from the
// second call on the block is DMA'd directly. The cache is keyed by the page-aligned range; a block that grows is
// registered again; at most HOSTREG_MAX ranges are kept (oldest out). JP_BWT_HOST_REGISTER=0 turns it off. Blocks that
// are already page-locked (jp_bwt_host_alloc, a caller's own cudaHostRegister) are recognised and left alone.
struct HostReg { uintptr_t lo, hi; unsigned long long stamp; };
// (heap objects that are never destroyed: the worker thread is detached and sleeps on the condition variable, and
// destroying a condition variable with a waiter -- which static destruction at exit would do -- blocks forever)
static std::mutex& g_hostreg_mu = *new std::mutex;
static std::condition_variable& g_hostreg_cv = *new std::condition_variable;
static std::vector<HostReg>& g_hostreg = *new std::vector<HostReg>;            // registered ranges
static std::vector<HostReg>& g_hostreg_todo = *new std::vector<HostReg>;       // ranges waiting for the worker
static bool g_hostreg_worker = false, g_hostreg_busy = false;
static uintptr_t g_hostreg_busy_lo = 0, g_hostreg_busy_hi = 0;
static unsigned long long g_hostreg_clock = 0;
static std::atomic<long long> g_hostreg_count{0}, g_hostreg_us{0};
constexpr size_t HOSTREG_MAX = 64, HOSTREG_MIN_BYTES = 1u << 20;
static bool hostreg_enabled()
{
	static int v = -1;
	if (v < 0) { const char* e = getenv("JP_BWT_HOST_REGISTER"); v = (e && *e == '0') ? 0 : 1; }
	return v == 1;
}
static void hostreg_worker()
{
	std::unique_lock<std::mutex> lk(g_hostreg_mu);
	for (;;) {
		g_hostreg_cv.wait(lk, [] { return !g_hostreg_todo.empty(); });
		const HostReg job = g_hostreg_todo.front();
		g_hostreg_todo.erase(g_hostreg_todo.begin());
		bool covered = false;
		for (auto& r : g_hostreg) covered |= (r.lo <= job.lo && job.hi <= r.hi);
		if (covered) continue;
		// ranges of ours that overlap the new one (the block grew, or was freed and its pages handed out again) go first
		for (size_t i = 0; i < g_hostreg.size();) {
			if (g_hostreg[i].lo < job.hi && job.lo < g_hostreg[i].hi) {
				if (cudaHostUnregister((void*)g_hostreg[i].lo) != cudaSuccess) cudaGetLastError();
				g_hostreg.erase(g_hostreg.begin() + (long)i);
			} else i++;
		}
		if (g_hostreg.size() >= HOSTREG_MAX) {
			size_t old = 0;
			for (size_t i = 1; i < g_hostreg.size(); i++) if (g_hostreg[i].stamp < g_hostreg[old].stamp) old = i;
			if (cudaHostUnregister((void*)g_hostreg[old].lo) != cudaSuccess) cudaGetLastError();
			g_hostreg.erase(g_hostreg.begin() + (long)old);
		}
		g_hostreg_busy = true; g_hostreg_busy_lo = job.lo; g_hostreg_busy_hi = job.hi;
		lk.unlock();                                        // the page-locking itself runs outside the lock
		{
			// register through a device this process already uses (a fresh thread would otherwise open device 0)
			int dev = -1;
			{ std::lock_guard<std::mutex> pl(g_pool.mu); for (auto& cx : g_pool.ctxs) { dev = cx->device; break; } }
			if (dev >= 0 && cudaSetDevice(dev) != cudaSuccess) cudaGetLastError();
		}
		const long long t0 = now_us();
		cudaPointerAttributes attr;
		bool ok = cudaPointerGetAttributes(&attr, (void*)job.lo) == cudaSuccess && attr.type == cudaMemoryTypeUnregistered;
		if (ok) ok = cudaHostRegister((void*)job.lo, job.hi - job.lo, cudaHostRegisterPortable) == cudaSuccess;
		if (!ok) cudaGetLastError();                        // (the copies still work, staged by the driver)
		const long long t1 = now_us();
		lk.lock();
		g_hostreg_busy = false;
		if (ok) { g_hostreg_us += t1 - t0; g_hostreg_count++; g_hostreg.push_back(HostReg{job.lo, job.hi, ++g_hostreg_clock}); }
		g_hostreg_cv.notify_all();
	}
}
// called when a stage call on [p, p + bytes) has completed
static void hostreg_seen(const void* p, size_t bytes)
{
	if (!hostreg_enabled() || !p || bytes < HOSTREG_MIN_BYTES) return;
	const uintptr_t page = 4096, lo = (uintptr_t)p & ~(page - 1), hi = ((uintptr_t)p + bytes + page - 1) & ~(page - 1);
	std::lock_guard<std::mutex> lk(g_hostreg_mu);
	for (auto& r : g_hostreg) if (r.lo <= lo && hi <= r.hi) { r.stamp = ++g_hostreg_clock; return; }
	for (auto& r : g_hostreg_todo) if (r.lo <= lo && hi <= r.hi) return;
	if (g_hostreg_busy && g_hostreg_busy_lo <= lo && hi <= g_hostreg_busy_hi) return;
	g_hostreg_todo.push_back(HostReg{lo, hi, 0});
	if (!g_hostreg_worker) {
		g_hostreg_worker = true;
		std::thread(hostreg_worker).detach();
		atexit([] {                                         // let a registration in progress finish before the runtime goes away
			std::unique_lock<std::mutex> lk2(g_hostreg_mu);
			g_hostreg_todo.clear();
			g_hostreg_cv.wait_for(lk2, std::chrono::seconds(5), [] { return !g_hostreg_busy; });
		});
	}
	g_hostreg_cv.notify_all();
}
static std::atomic<long long>& hostreg_counter(int which) { return which == 0 ? g_hostreg_count : g_hostreg_us; }
// p == nullptr: every range; else the ranges that contain p
static void hostreg_release(const void* p)
{
	std::unique_lock<std::mutex> lk(g_hostreg_mu);
	const uintptr_t a = (uintptr_t)p;
	for (size_t i = 0; i < g_hostreg_todo.size();) {
		if (!p || (g_hostreg_todo[i].lo <= a && a < g_hostreg_todo[i].hi)) g_hostreg_todo.erase(g_hostreg_todo.begin() + (long)i); else i++;
	}
	g_hostreg_cv.wait(lk, [&] { return !g_hostreg_busy || (p && !(g_hostreg_busy_lo <= a && a < g_hostreg_busy_hi)); });
	for (size_t i = 0; i < g_hostreg.size();) {
		if (!p || (g_hostreg[i].lo <= a && a < g_hostreg[i].hi)) {
			if (cudaHostUnregister((void*)g_hostreg[i].lo) != cudaSuccess) cudaGetLastError();
			g_hostreg.erase(g_hostreg.begin() + (long)i);
		} else i++;
	}
}

static void begin_call(Ctx& c)
{
	c.launches = 0;
	c.arena.reset();
	c.arena.high = 0;
	c.arena2.high = 0;
	memset(&t_stats, 0, sizeof(t_stats));
	t_detail[0] = 0;
}

static int host_call_once(Ctx& c, int direction, const u8* in, i32 in_len, i32 len, i32 nlen, u8* out);

static int host_call(int direction, const u8* in, i32 in_len, u8* out, i32* out_len)
{
	if (!in || !out || !out_len || in_len < 0) { set_error_detail("null pointer or negative length"); return JP_ERR_ARG; }
	if (direction == 1 && in_len < JP_BWT_TRAILER_BYTES) { set_error_detail("inverse input shorter than its trailer"); return JP_ERR_ARG; }
	const i32 len = direction == 0 ? in_len : in_len - JP_BWT_TRAILER_BYTES;
	if ((i64)len > JP_BWT_MAX_CALL_LEN) { set_error_detail("block longer than %lld bytes", (long long)JP_BWT_MAX_CALL_LEN); return JP_ERR_ARG; }
	const i32 nlen = len - len % JP_BWT_UNITS;
	*out_len = direction == 0 ? len + JP_BWT_TRAILER_BYTES : len;          // bwt.cpp:27 / :78
	const bool tr = trace_enabled();
	const long long t_begin = tr ? now_us() : 0;
	CtxGuard g;
	JP_TRY(acquire(-1, &g.c));
	const long long t_acquired = tr ? now_us() : 0;
	Ctx& c = *g.c;
	int rc = JP_ERR_OOM;
	for (int attempt = 0; attempt < 256 && rc == JP_ERR_OOM; attempt++) {
		if (attempt > 0 && !relieve_memory_pressure(c)) break;
		rc = host_call_once(c, direction, in, in_len, len, nlen, out);
	}
	if (rc != JP_OK) return rc;
	if (tr) {
		const long long t_end = now_us();
		trace_add(direction, len, t_begin, t_end, t_stats.ms_total + t_stats.ms_h2d + t_stats.ms_d2h);
		if (g_trace_calls)
			fprintf(stderr, "[jp_bwt call] dir=%d len=%d dev=%d t0_us=%lld wait_us=%lld call_us=%lld h2d_ms=%.3f kernels_ms=%.3f d2h_ms=%.3f\n", direction, len,
			        c.device, t_begin, t_acquired - t_begin, t_end - t_begin, t_stats.ms_h2d, t_stats.ms_total, t_stats.ms_d2h);
	}
	return JP_OK;
}

static int host_call_once(Ctx& c, int direction, const u8* in, i32 in_len, i32 len, i32 nlen, u8* out)
{
	begin_call(c);
	cudaStream_t s = c.own_stream;
	JP_TRY(ensure_io(c, (size_t)len + JP_BWT_TRAILER_BYTES));
	const size_t in_bytes = (size_t)in_len;
	// forward: the trailer exists only when something was transformed (bwt.cpp:35)
	const size_t out_bytes = direction == 0 ? (size_t)len + (nlen > 0 ? JP_BWT_TRAILER_BYTES : 0) : (size_t)len;
	JP_CUDA(cudaEventRecord(c.ev[8], s));
	if (in_bytes) JP_CUDA(cudaMemcpyAsync(c.d_in, in, in_bytes, cudaMemcpyHostToDevice, s));
	JP_CUDA(cudaEventRecord(c.ev[9], s));
	int rc = direction == 0 ? forward_device(c, c.d_in, len, c.d_out, s, &t_stats)
	                        : inverse_device(c, c.d_in, in_len, c.d_out, s, &t_stats, c.d_in);   // our own copy: free scratch
	if (rc != JP_OK) return rc;
	JP_CUDA(cudaEventRecord(c.ev[10], s));
	if (out_bytes) JP_CUDA(cudaMemcpyAsync(out, c.d_out, out_bytes, cudaMemcpyDeviceToHost, s));
	JP_CUDA(cudaEventRecord(c.ev[11], s));
	JP_CUDA(cudaStreamSynchronize(s));
	JP_CUDA(cudaEventElapsedTime(&t_stats.ms_h2d, c.ev[8], c.ev[9]));
	JP_CUDA(cudaEventElapsedTime(&t_stats.ms_d2h, c.ev[10], c.ev[11]));
	t_stats.kernel_launches = c.launches;
	t_stats.device_bytes = c.arena.high + c.arena2.high;
	hostreg_seen(in, in_bytes);                             // page-locked in the background for the calls to come
	hostreg_seen(out, (size_t)len + JP_BWT_TRAILER_BYTES);
	return JP_OK;
}

static int device_call(int direction, const u8* d_in, i32 in_len, u8* d_out, int device, void* stream, bool consume_in = false)
{
	if (!d_in || !d_out || in_len < 0 || device < 0) { set_error_detail("null pointer, negative length or device"); return JP_ERR_ARG; }
	if (direction == 1 && in_len < JP_BWT_TRAILER_BYTES) { set_error_detail("inverse input shorter than its trailer"); return JP_ERR_ARG; }
	if (((uintptr_t)d_in & 15) || ((uintptr_t)d_out & 15)) { set_error_detail("device blocks must be 16-byte aligned"); return JP_ERR_ARG; }
	if ((i64)in_len - (direction == 1 ? JP_BWT_TRAILER_BYTES : 0) > JP_BWT_MAX_CALL_LEN) { set_error_detail("block longer than %lld bytes", (long long)JP_BWT_MAX_CALL_LEN); return JP_ERR_ARG; }
	CtxGuard g;
	JP_TRY(acquire(device, &g.c));
	Ctx& c = *g.c;
	cudaStream_t s = stream ? (cudaStream_t)stream : c.own_stream;
	int rc = JP_ERR_OOM;
	for (int attempt = 0; attempt < 256 && rc == JP_ERR_OOM; attempt++) {
		if (attempt > 0 && !relieve_memory_pressure(c)) break;
		begin_call(c);
		rc = direction == 0 ? forward_device(c, d_in, in_len, d_out, s, &t_stats)
		                    : inverse_device(c, d_in, in_len, d_out, s, &t_stats, consume_in ? const_cast<u8*>(d_in) : nullptr);
	}
	t_stats.kernel_launches = c.launches;
	t_stats.device_bytes = c.arena.high + c.arena2.high;
	return rc;
}

} // namespace jp

using namespace jp;

extern "C" {

int jp_bwt_forward(const uint8_t* in, int32_t len, uint8_t* out, int32_t* out_len) { return host_call(0, in, len, out, out_len); }
int jp_bwt_inverse(const uint8_t* in, int32_t len_with_trailer, uint8_t* out, int32_t* out_len) { return host_call(1, in, len_with_trailer, out, out_len); }
int jp_bwt_forward_device(const uint8_t* d_in, int32_t len, uint8_t* d_out, int device, void* stream) { return device_call(0, d_in, len, d_out, device, stream); }
int jp_bwt_inverse_device(const uint8_t* d_in, int32_t len_with_trailer, uint8_t* d_out, int device, void* stream) { return device_call(1, d_in, len_with_trailer, d_out, device, stream); }
int jp_bwt_inverse_device_consume(uint8_t* d_in, int32_t len_with_trailer, uint8_t* d_out, int device, void* stream) { return device_call(1, d_in, len_with_trailer, d_out, device, stream, true); }

int jp_src_rle0_device(const uint8_t* d_in, int32_t len, int32_t* d_freq, uint16_t* d_rle, int32_t* d_rlen, int device, void* stream)
{
	if (!d_in || !d_freq || !d_rle || !d_rlen || len < 0 || device < 0) { set_error_detail("null pointer, negative length or device"); return JP_ERR_ARG; }
	if ((uintptr_t)d_in & 15) { set_error_detail("device blocks must be 16-byte aligned"); return JP_ERR_ARG; }
	CtxGuard g;
	JP_TRY(acquire(device, &g.c));
	Ctx& c = *g.c;
	cudaStream_t s = stream ? (cudaStream_t)stream : c.own_stream;
	int rc = JP_ERR_OOM;
	for (int attempt = 0; attempt < 256 && rc == JP_ERR_OOM; attempt++) {
		if (attempt > 0 && !relieve_memory_pressure(c)) break;
		begin_call(c);
		rc = src_rle0_device(c, d_in, len, d_freq, d_rle, d_rlen, s, &t_stats);
	}
	t_stats.kernel_launches = c.launches;
	return rc;
}

int jp_src_rle0(const uint8_t* in, int32_t len, int32_t* freq, uint16_t* rle, int32_t* rlen)
{
	if (!in || !freq || !rle || !rlen || len < 0) { set_error_detail("null pointer or negative length"); return JP_ERR_ARG; }
	if (len == 0) return JP_OK;
	CtxGuard g;
	JP_TRY(acquire(-1, &g.c));
	Ctx& c = *g.c;
	cudaStream_t s = c.own_stream;
	const size_t n = (size_t)len, nchunk = (n + JP_ANS_CHUNK - 1) / JP_ANS_CHUNK;
	JP_TRY(ensure_io(c, n));                                            // d_in: the block
	// outputs: one allocation per call (this entry point is a convenience for hosts without device buffers)
	u8* d_o = nullptr;
	const size_t b_rle = Arena::align(n * 2), b_freq = Arena::align(nchunk * 256 * 4), b_rlen = Arena::align(nchunk * 4);
	JP_TRY(dev_alloc(c.device, (void**)&d_o, b_rle + b_freq + b_rlen));
	struct Free { int dev; void* p; size_t b; ~Free() { dev_free(dev, p, b); } } fr{c.device, d_o, b_rle + b_freq + b_rlen};
	u16* d_rle = (u16*)d_o; i32* d_freq = (i32*)(d_o + b_rle); i32* d_rlen = (i32*)(d_o + b_rle + b_freq);
	JP_CUDA(cudaMemcpyAsync(c.d_in, in, n, cudaMemcpyHostToDevice, s));
	int rc = JP_ERR_OOM;
	for (int attempt = 0; attempt < 256 && rc == JP_ERR_OOM; attempt++) {
		if (attempt > 0 && !relieve_memory_pressure(c)) break;
		begin_call(c);
		rc = src_rle0_device(c, c.d_in, len, d_freq, d_rle, d_rlen, s, &t_stats);
	}
	if (rc != JP_OK) return rc;
	JP_CUDA(cudaMemcpyAsync(freq, d_freq, nchunk * 256 * 4, cudaMemcpyDeviceToHost, s));
	JP_CUDA(cudaMemcpyAsync(rlen, d_rlen, nchunk * 4, cudaMemcpyDeviceToHost, s));
	JP_CUDA(cudaMemcpyAsync(rle, d_rle, n * 2, cudaMemcpyDeviceToHost, s));   // (chunk k's symbols start at rle[k * JP_ANS_CHUNK])
	JP_CUDA(cudaStreamSynchronize(s));
	t_stats.kernel_launches = c.launches;
	return JP_OK;
}

int jp_bwt_set_devices(const int* ids, int n)
{
	std::lock_guard<std::mutex> lk(g_pool.mu);
	if (n <= 0 || !ids) { g_pool.devices_ready = false; return init_devices_locked(); }
	const int vis = visible_devices();
	if (vis == 0) { set_error_detail("no CUDA device visible; this stage has no CPU path"); return JP_ERR_NO_DEVICE; }
	std::vector<int> d;
	for (int i = 0; i < n; i++) { if (ids[i] < 0 || ids[i] >= vis) { set_error_detail("device %d not visible", ids[i]); return JP_ERR_ARG; } d.push_back(ids[i]); }
	g_pool.devices = d;
	g_pool.devices_ready = true;
	if (g_pool.pending.size() < (size_t)vis) g_pool.pending.resize((size_t)vis, 0);
	if (g_pool.state.size() < (size_t)vis) g_pool.state.resize((size_t)vis, 0);
	if (g_pool.state[d[0]] == 0) g_pool.state[d[0]] = 2;
	g_pool.rr = 0;
	return JP_OK;
}

int jp_bwt_warmup_async(void)
{
	// Starts bringing up primary contexts in the background: called by the shim's static initialiser, so that the
	// start-up (0.3-2 s per device, and the driver creates them one after the other) overlaps the reference's file read
	// and LZ77 of the first batch instead of the first stage calls. By default only the FIRST configured device is
	// started here and the others follow on demand, when the usable ones stay saturated (warm_next_device_locked):
	// measured through the reference CLI, a 1 GiB job keeps one B200 busy for a fraction of a second per batch, and
	// every further context costs about a second to create and as much to tear down at exit -- with all of 4 devices
	// started up front the decompression of that job took 7.9 s against 4.8 s on one (profiles/pipeline_*_r02.json).
	// JP_BWT_WARM=all starts every configured device at load (long jobs on many GPUs).
	{
		std::lock_guard<std::mutex> lk(g_pool.mu);
		const int rc = init_devices_locked();
		if (rc != JP_OK) return rc;
		static bool started = false;
		if (started) return JP_OK;
		started = true;
		const char* mode = getenv("JP_BWT_WARM");
		const bool all = mode && strcmp(mode, "all") == 0;
		bool first = true;
		for (int d : g_pool.devices) {
			if (!all && !first) break;
			if (g_pool.state[d] == 0 || g_pool.state[d] == 2) g_pool.state[d] = g_pool.state[d] == 2 ? 4 : 1;   // 4: usable, context not forced yet
			first = false;
		}
		static bool hooked = false;
		if (!hooked) {
			hooked = true;
			atexit([] {                                  // never tear the runtime down under a device that is coming up
				const long long t0 = now_us();
				{ std::lock_guard<std::mutex> lk2(g_pool.mu); g_shutting_down = true; }
				struct Done { long long t0; ~Done() { if (g_trace_on.load() == 1) fprintf(stderr, "[jp_bwt warmup] exit waited %.3f s for start-ups in progress\n", (now_us() - t0) / 1e6); } } done{t0};
				for (int spin = 0; spin < 20000; spin++) {
					{ std::lock_guard<std::mutex> lk2(g_pool.mu); bool busy = false; for (int st : g_pool.state) busy |= (st == 1); if (!busy) return; }
					std::this_thread::sleep_for(std::chrono::milliseconds(1));
				}
			});
		}
	}
	std::thread([] {
		std::vector<int> devs;
		{ std::lock_guard<std::mutex> lk(g_pool.mu); devs = g_pool.devices; }
		const long long t_load = now_us();
		for (int d : devs) {
			{
				std::lock_guard<std::mutex> lk(g_pool.mu);
				if (g_pool.state[d] != 1 && g_pool.state[d] != 4) continue;
				if (g_shutting_down) { if (g_pool.state[d] == 1) g_pool.state[d] = 0; continue; }   // the process is leaving: no more start-ups
			}
			const long long t0 = now_us();
			const bool ok = cudaSetDevice(d) == cudaSuccess && cudaFree(0) == cudaSuccess;
			{ std::lock_guard<std::mutex> lk(g_pool.mu); g_pool.state[d] = ok ? 2 : 3; }
			g_pool.cv.notify_all();
			if (trace_enabled()) fprintf(stderr, "[jp_bwt warmup] device %d %s after %.3f s (+%.3f s since load)\n", d, ok ? "up" : "FAILED", (now_us() - t0) / 1e6, (now_us() - t_load) / 1e6);
		}
	}).detach();
	return JP_OK;
}

int jp_bwt_debug_copy(const uint8_t* in, int32_t in_len, uint8_t* out, int32_t out_len)
{
	// the host<->device copies of a stage call and nothing else: the ceiling of the end-to-end figure (bench.py)
	if (!in || !out || in_len < 0 || out_len < 0) return JP_ERR_ARG;
	CtxGuard g; JP_TRY(acquire(-1, &g.c));
	Ctx& c = *g.c;
	begin_call(c);
	cudaStream_t s = c.own_stream;
	JP_TRY(ensure_io(c, (size_t)std::max(in_len, out_len)));
	if (in_len) JP_CUDA(cudaMemcpyAsync(c.d_in, in, (size_t)in_len, cudaMemcpyHostToDevice, s));
	if (out_len) JP_CUDA(cudaMemcpyAsync(out, c.d_out, (size_t)out_len, cudaMemcpyDeviceToHost, s));
	JP_CUDA(cudaStreamSynchronize(s));
	hostreg_seen(in, (size_t)in_len);
	hostreg_seen(out, (size_t)out_len);
	return JP_OK;
}

int jp_bwt_device_count(void)
{
	std::lock_guard<std::mutex> lk(g_pool.mu);
	if (init_devices_locked() != JP_OK) return 0;
	return (int)g_pool.devices.size();
}

void* jp_bwt_host_alloc(uint64_t bytes)
{
	void* p = nullptr;
	if (cudaMallocHost(&p, bytes ? bytes : 1) == cudaSuccess) return p;
	cudaGetLastError();
	hostreg_release(nullptr);                       // page-locked memory is a limited resource: give the cached caller blocks back and retry
	if (cudaMallocHost(&p, bytes ? bytes : 1) == cudaSuccess) return p;
	cudaGetLastError();
	return nullptr;
}
void jp_bwt_host_release(const void* p) { hostreg_release(p); }
void jp_bwt_host_free(void* p) { if (p) cudaFreeHost(p); }

int jp_bwt_last_stats(jp_bwt_stats* out) { if (!out) return JP_ERR_ARG; *out = t_stats; return JP_OK; }

const char* jp_bwt_strerror(int rc)
{
	switch (rc) {
	case JP_OK: return "ok";
	case JP_ERR_ARG: return "Bwt :: invalid argument";
	case JP_ERR_NO_DEVICE: return "Bwt :: no CUDA device (this stage has no CPU path)";
	case JP_ERR_CUDA: return "Bwt :: CUDA error";
	case JP_ERR_OOM: return "Bwt :: out of device or pinned memory";
	case JP_ERR_BAD_INDEX: return "Bwt :: stored primary index out of range";
	case JP_ERR_CORRUPT: return "Bwt :: inconsistent BWT block";
	case JP_ERR_INTERNAL: return "Bwt :: internal invariant failed";
	default: return "Bwt :: unknown error";
	}
}
const char* jp_bwt_last_error_detail(void) { return t_detail; }
const char* jp_bwt_version(void) { return "jampack-bwt-b200 0.1 (sm_100a)"; }

int jp_bwt_debug_lf(const uint8_t* in, int32_t nlen, int32_t* lf, int32_t* ctable)
{
	if (!in || !lf || !ctable) return JP_ERR_ARG;
	CtxGuard g; JP_TRY(acquire(-1, &g.c)); begin_call(*g.c);
	return debug_lf(*g.c, in, nlen, lf, ctable);
}
int jp_bwt_suffix_array(const uint8_t* in, int32_t n, int32_t* sa)
{
	if (n < 0 || (n > 0 && (!in || !sa))) { set_error_detail("null pointer or negative length"); return JP_ERR_ARG; }
	if ((i64)n > JP_BWT_MAX_CALL_LEN) { set_error_detail("block longer than %lld bytes", (long long)JP_BWT_MAX_CALL_LEN); return JP_ERR_ARG; }
	if (n == 0) return JP_OK;
	CtxGuard g; JP_TRY(acquire(-1, &g.c));
	// same out-of-memory policy as the stage entry points: -m2 blocks run from the same OpenMP team (lz77.cpp:141)
	int rc = JP_ERR_OOM;
	for (int attempt = 0; attempt < 256 && rc == JP_ERR_OOM; attempt++) {
		if (attempt > 0 && !relieve_memory_pressure(*g.c)) break;
		begin_call(*g.c);
		rc = debug_suffix_array(*g.c, in, n, sa);
	}
	t_stats.kernel_launches = g.c->launches;
	t_stats.device_bytes = g.c->arena.high + g.c->arena2.high;
	return rc;
}
double jp_bwt_debug_gather_rate(uint64_t table_bytes, int32_t chains, int32_t steps, int dependent)
{
	CtxGuard g; int rc = acquire(-1, &g.c); if (rc != JP_OK) return rc; begin_call(*g.c);
	return debug_gather_rate(*g.c, table_bytes, chains, steps, dependent);
}

}
