// simt_runtime.hpp -- TEST INFRASTRUCTURE ONLY: a small SIMT emulator that runs the product's CUDA kernels, unmodified
// apart from mechanical rewriting of the launch syntax and the inline PTX (tests/simt/gen.py), on the CPU.
//
// It is NOT a CPU path of the product (nothing under jampack_b200/ knows about it, the C-ABI library never links it):
// it exists so that `pytest -m "not gpu"` exercises the kernels' LOGIC -- warp collectives, shared memory, tickets,
// scans, the replay of the single-walk inverse -- against the oracle on small blocks where no GPU is at hand.
//
// Model: blocks run one after the other; the threads of a block are fibers (ucontext) scheduled round-robin; a fiber
// runs until it reaches a warp collective or a block barrier, where it waits for the other live threads of its warp /
// block. Full-mask collectives must be reached by every live lane of the warp (as on the hardware); lanes that have
// returned from the kernel no longer count. Atomics are plain operations (one OS thread). There is no memory model
// to speak of: racecheck on the GPU is the tool for that. Kernels that spin on other blocks would deadlock here.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <ucontext.h>
#include <vector>

// ---- CUDA vocabulary -------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __shared__ static
#define CUDART_CB

struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }
struct ulonglong2 { unsigned long long x, y; };
static inline ulonglong2 make_ulonglong2(unsigned long long a, unsigned long long b) { return ulonglong2{a, b}; }

namespace simt {
extern uint3 g_threadIdx, g_blockIdx;
extern dim3 g_blockDim, g_gridDim;
}
#define threadIdx (simt::g_threadIdx)
#define blockIdx (simt::g_blockIdx)
#define blockDim (simt::g_blockDim)
#define gridDim (simt::g_gridDim)
#define warpSize 32

using std::min;
using std::max;

// ---- runtime API shims: device memory is host memory, streams are immediate --------------------------------------
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum cudaError_t { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorUnknown = 999 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "simt emulation"; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
template <typename T> static inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <typename F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 2; return cudaSuccess; }

// ---- the emulator ------------------------------------------------------------------------------------------------
namespace simt {

enum Op { OP_NONE, OP_BALLOT, OP_SHFL, OP_SHFL_UP, OP_SHFL_DOWN, OP_SHFL_XOR, OP_MATCH, OP_SYNCWARP };

struct Warp {
	unsigned live = 0, arrived = 0, gen = 0;
	Op op = OP_NONE;
	uint64_t val[32]; int arg[32]; uint64_t out[32];
};

struct Fiber { ucontext_t ctx; void* sp = nullptr; char* stack = nullptr; bool done = true; };

struct Block {
	std::vector<Fiber> fibers;
	std::vector<Warp> warps;
	unsigned nthreads = 0, live_threads = 0, bar_arrived = 0, bar_gen = 0, bar_or = 0, bar_or_out = 0;
	int current = -1;
	ucontext_t main_ctx;
	const std::function<void()>* body = nullptr;
};

extern Block g_block;
extern unsigned char* g_dynamic_smem;
extern long long g_launches;
constexpr size_t STACK_BYTES = 256 << 10;

void yield();
void launch(dim3 grid, dim3 block, size_t smem, cudaStream_t, const std::function<void()>& body);
// `kernel<<<grid, block[, smem[, stream]]>>>(args)` is rewritten to `simt::launch_cfg(grid, block, ...)([&]{ kernel(args); })`
struct LaunchCfg {
	dim3 grid, block; size_t smem;
	void operator()(const std::function<void()>& body) const { launch(grid, block, smem, nullptr, body); }
};
inline LaunchCfg launch_cfg(dim3 g, dim3 b, size_t smem = 0, cudaStream_t = nullptr) { return LaunchCfg{g, b, smem}; }
void fail(const char* what);

inline unsigned lane() { return g_threadIdx.x & 31; }
inline Warp& my_warp() { return g_block.warps[g_threadIdx.x >> 5]; }
inline unsigned lanemask_lt() { return (1u << lane()) - 1u; }

void complete(Warp& w);

// every live lane of the warp meets here; `v`/`a` are the lane's operands, the return value its result
inline uint64_t collective(Op op, uint64_t v, int a)
{
	Warp& w = my_warp();
	const unsigned l = lane();
	if (w.arrived == 0) w.op = op; else if (w.op != op) fail("lanes of one warp reached different collectives (divergent full-mask collective)");
	w.val[l] = v; w.arg[l] = a; w.arrived |= 1u << l;
	const unsigned mygen = w.gen;
	if (w.arrived == w.live) complete(w);
	while (w.gen == mygen) yield();
	return w.out[l];
}

} // namespace simt

static inline void __syncwarp(unsigned = 0xffffffffu) { simt::collective(simt::OP_SYNCWARP, 0, 0); }
static inline unsigned __ballot_sync(unsigned, int pred) { return (unsigned)simt::collective(simt::OP_BALLOT, pred != 0, 0); }
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, !pred) == 0; }
template <typename T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) { uint64_t x = 0; memcpy(&x, &v, sizeof(T)); x = simt::collective(simt::OP_SHFL, x, src); T r; memcpy(&r, &x, sizeof(T)); return r; }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) { uint64_t x = 0; memcpy(&x, &v, sizeof(T)); x = simt::collective(simt::OP_SHFL_UP, x, (int)d); T r; memcpy(&r, &x, sizeof(T)); return r; }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) { uint64_t x = 0; memcpy(&x, &v, sizeof(T)); x = simt::collective(simt::OP_SHFL_DOWN, x, (int)d); T r; memcpy(&r, &x, sizeof(T)); return r; }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { uint64_t x = 0; memcpy(&x, &v, sizeof(T)); x = simt::collective(simt::OP_SHFL_XOR, x, m); T r; memcpy(&r, &x, sizeof(T)); return r; }
template <typename T> static inline unsigned __match_any_sync(unsigned, T v) { uint64_t x = 0; memcpy(&x, &v, sizeof(T)); return (unsigned)simt::collective(simt::OP_MATCH, x, 0); }

static inline unsigned __reduce_add_sync(unsigned m, unsigned v) { for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(m, v, o); return v; }

void __syncthreads();
int __syncthreads_or(int pred);
static inline int __syncthreads_and(int pred) { return !__syncthreads_or(!pred); }

static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) { sh &= 31; return sh ? (hi << sh) | (lo >> (32 - sh)) : hi; }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcs(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }

template <typename T, typename U> static inline T atomicAdd(T* p, U v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <typename T, typename U> static inline T atomicOr(T* p, U v) { T o = *p; *p = (T)(o | (T)v); return o; }
template <typename T, typename U> static inline T atomicAnd(T* p, U v) { T o = *p; *p = (T)(o & (T)v); return o; }
template <typename T, typename U> static inline T atomicMax(T* p, U v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <typename T, typename U> static inline T atomicMin(T* p, U v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <typename T, typename U> static inline T atomicExch(T* p, U v) { T o = *p; *p = (T)v; return o; }
template <typename T, typename U, typename V> static inline T atomicCAS(T* p, U cmp, V v) { T o = *p; if (o == (T)cmp) *p = (T)v; return o; }
