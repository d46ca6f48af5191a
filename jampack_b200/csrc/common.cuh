// common.cuh -- shared types, error plumbing and warp/block primitives for the sm_100a BWT kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

typedef uint8_t  u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef int32_t  i32;
typedef uint64_t u64;
typedef int64_t  i64;

namespace jp {

// ---- device-side error flag ---------------------------------------------------------------------
// Kernels never trap: a failed check records the first failing code and the kernel carries on with
// in-bounds (possibly meaningless) data; the host reads the flag once per call.
enum DevErr : int {
	DE_NONE = 0,
	DE_BAD_INDEX = 1,      // inverse: stored index outside [1, nlen] or anchors not distinct
	DE_CHAIN_LEN = 2,      // inverse: a decode unit's chain is not exactly `step` long
	DE_CHAIN_RANGE = 3,    // inverse: a sub-chain would write outside its block
	DE_RANK_LOOP = 4,      // inverse: list ranking did not reach an anchor (cyclic garbage input)
	DE_FWD_RANGE = 5,      // forward: an active suffix asked for a rank beyond the end (invariant)
	DE_FWD_ROUNDS = 6,     // forward: doubling did not converge within the round limit
	DE_STREAM_OVERFLOW = 7, // inverse, single-walk path: stream space or a length field ran out (host reruns the two-pass path)
};

__device__ __forceinline__ void dev_fail(int* flag, int code) { atomicCAS(flag, 0, code); }

// ---- small helpers ------------------------------------------------------------------------------
__host__ __device__ __forceinline__ u32 mix32(u32 x)
{
	x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
	return x;
}

__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ u32 lanemask_lt() { u32 m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

__device__ __forceinline__ u32 warp_incl_sum(u32 v)
{
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, v, o); if (lane_id() >= (u32)o) v += t; }
	return v;
}
__device__ __forceinline__ i32 warp_incl_max(i32 v)
{
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { i32 t = __shfl_up_sync(0xffffffffu, v, o); if (lane_id() >= (u32)o) v = max(v, t); }
	return v;
}
__device__ __forceinline__ u32 warp_sum(u32 v)
{
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}
__device__ __forceinline__ i32 warp_max(i32 v)
{
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}

// Block-wide inclusive sum over blockDim.x (multiple of 32, <= 1024) threads.
// `ws` = 32 u32 of shared scratch. Returns inclusive prefix; *total = block sum. Two barriers.
__device__ __forceinline__ u32 block_incl_sum(u32 v, u32* ws, u32* total)
{
	const u32 lane = lane_id(), w = threadIdx.x >> 5, nw = blockDim.x >> 5;
	u32 inc = warp_incl_sum(v);
	if (lane == 31) ws[w] = inc;
	__syncthreads();
	u32 wv = (lane < nw) ? ws[lane] : 0;
	u32 winc = warp_incl_sum(wv);
	u32 wprefix = __shfl_sync(0xffffffffu, winc, w) - __shfl_sync(0xffffffffu, wv, w);
	*total = __shfl_sync(0xffffffffu, winc, 31);
	__syncthreads();
	return inc + wprefix;
}

// Lanes holding the same 8-bit value, from eight ballots. match.any gives the same mask in one instruction, but
// it runs on the ADU pipe at a cost proportional to the number of DISTINCT values in the warp (ncu: k_rs_hist
// 0.43 ms with ~30 distinct digits per warp against 0.18 ms with a few; pipe_adu 97 % busy). This form costs the
// same 16 vote/logic instructions whatever the data. `v` must be < 256 in every participating lane.
__device__ __forceinline__ u32 match_any8(u32 v)
{
	u32 peers = 0xffffffffu;
	#pragma unroll
	for (int b = 0; b < 8; b++) {
		const bool bit = (v >> b) & 1u;
		const u32 m = __ballot_sync(0xffffffffu, bit);
		peers &= bit ? m : ~m;
	}
	return peers;
}

// Warp-uniform choice between the two: neighbouring lanes that differ are a cheap proxy for the number of distinct
// values (sorted or run-heavy data has few boundaries and match.any is then the faster form).
template <int MAX_BOUNDARIES = 24>
__device__ __forceinline__ u32 match_any8_adaptive(u32 v)
{
	const u32 prev = __shfl_up_sync(0xffffffffu, v, 1);
	const u32 boundaries = __popc(__ballot_sync(0xffffffffu, v != prev));
	return boundaries <= MAX_BOUNDARIES ? __match_any_sync(0xffffffffu, v) : match_any8(v);
}

__host__ __device__ __forceinline__ int bit_length(u64 x) { int b = 0; while (x) { b++; x >>= 1; } return b; }

} // namespace jp

// ---- host-side CUDA call checking ----------------------------------------------------------------
namespace jp { void set_error_detail(const char* fmt, ...); }

#define JP_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { \
	jp::set_error_detail("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
	return (e__ == cudaErrorMemoryAllocation) ? JP_ERR_OOM : JP_ERR_CUDA; } } while (0)

#define JP_TRY(expr) do { int rc__ = (expr); if (rc__ != JP_OK) return rc__; } while (0)
