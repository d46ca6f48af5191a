// radix_sort.cuh -- hand-written LSD radix sort of (64-bit key, 32-bit value) pairs for sm_100a.
//
// This is the "radix-bucketed" half of the forward transform: it orders suffixes by their packed
// prefix keys and, in every doubling round, (group, rank-of-continuation) keys. 8-bit digits, stable.
// One pass = per-tile digit histogram -> per-digit exclusive scan over tiles -> ranked scatter. The scatter
// ranks keys inside a warp with match.any (no atomics, stable), re-orders the tile in shared memory so that
// every digit's run leaves as coalesced stores, and adds the tile's global offsets.
// No inter-block dependency (no look-back spin) -- a pass cannot hang. (A one-sweep variant with decoupled look-back was
// built and measured in round 1: 6.2 ms against 5.1 ms for 8 passes over 64 M pairs; removed.)
#pragma once
#include "common.cuh"

namespace jp {

#ifndef RS_THREADS_CFG
#define RS_THREADS_CFG 256
#endif
constexpr int RS_THREADS = RS_THREADS_CFG;          // 256 (16 pairs per thread) measured best; 512 x 8 is the A/B variant
constexpr int RS_WARPS   = RS_THREADS / 32;
#ifndef RS_ITEMS_CFG
#define RS_ITEMS_CFG 16
#endif
constexpr int RS_ITEMS   = RS_ITEMS_CFG;            // pairs per thread (measured: profiles/radix_items_r02.md)
constexpr int RS_TILE    = RS_THREADS * RS_ITEMS;   // pairs per block
static_assert(RS_TILE % 32 == 0 && RS_TILE % 16 == 0, "tiles are whole words of the run bitmap and whole 16-byte digit loads");
constexpr size_t RS_SMEM_SCATTER = (size_t)RS_TILE * (8 + 4);

__device__ __forceinline__ u32 rs_digit(u64 k, int shift) { return (u32)(k >> shift) & 255u; }
__device__ __forceinline__ u32 rs_digit(u32 k, int shift) { return (k >> shift) & 255u; }

// Histogram layout is digit-major, tile_hist[digit * stride + tile] (stride = tiles rounded up to 4), so that
// the per-digit scan over tiles reads and writes whole rows (the tile-major layout made that scan as
// expensive as the scatter itself: 0.31 ms per pass at 16 K tiles, ncu round 1).
__host__ __device__ __forceinline__ u32 rs_stride(u32 tiles) { return (tiles + 3u) & ~3u; }

template <typename K>
__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const K* __restrict__ keys, u32 n, int shift, u32* __restrict__ tile_hist, u32 stride)
{
	__shared__ u32 h[RS_WARPS][256];
	const int t = threadIdx.x, w = t >> 5, lane = t & 31;
	for (int i = t; i < RS_WARPS * 256; i += RS_THREADS) (&h[0][0])[i] = 0;
	__syncthreads();
	const u32 base = blockIdx.x * RS_TILE + w * (32 * RS_ITEMS);
	u32* hw = h[w];
	#pragma unroll
	for (int i = 0; i < RS_ITEMS; i++) {
		const u32 p = base + i * 32 + lane;
		// keys of neighbouring suffixes often share a digit: aggregate inside the warp before the atomic
		const bool ok = p < n;
		const u32 d = ok ? rs_digit(keys[p], shift) : 0u;
		// few distinct digits in the warp: aggregate with match.any and add once per digit; many: one shared atomic
		// per lane (hardly any two lanes collide, and neither match.any nor eight ballots is cheaper than that)
		const u32 prev = __shfl_up_sync(0xffffffffu, d, 1);
		const u32 boundaries = __popc(__ballot_sync(0xffffffffu, d != prev));
		if (boundaries <= 24) {
			const u32 peers = __match_any_sync(0xffffffffu, d) & __ballot_sync(0xffffffffu, ok);
			if (ok && (peers & lanemask_lt()) == 0) atomicAdd(&hw[d], (u32)__popc(peers));
		} else if (ok) atomicAdd(&hw[d], 1u);
	}
	__syncthreads();
	if (t < 256) {
		u32 s = 0;
		#pragma unroll
		for (int k = 0; k < RS_WARPS; k++) s += h[k][t];
		tile_hist[(size_t)t * stride + blockIdx.x] = s;
	}
}

// The same tile histogram from a one-byte-per-key digit array. Every scatter pass leaves such an array for the NEXT
// digit next to the pairs it writes (1 extra byte per pair), so the histogram of passes 2.. reads 1 byte per key
// instead of 8 (ncu: k_rs_hist was DRAM-bound on the 8-byte keys, 0.18-0.21 ms per 64 M keys).
__global__ void __launch_bounds__(RS_THREADS) k_rs_hist_bytes(const u8* __restrict__ digits, u32 n, u32* __restrict__ tile_hist, u32 stride)
{
	__shared__ u32 h[RS_WARPS][256];
	const int t = threadIdx.x, w = t >> 5;
	for (int i = t; i < RS_WARPS * 256; i += RS_THREADS) (&h[0][0])[i] = 0;
	__syncthreads();
	u32* hw = h[w];
	const u32 p = blockIdx.x * RS_TILE + t * 16;        // one 16-byte load per thread (the first RS_TILE / 16 threads)
	if (t * 16 >= RS_TILE) {}
	else if (p + 16 <= n) {
		const uint4 q = *reinterpret_cast<const uint4*>(digits + p);
		const u32 wd[4] = {q.x, q.y, q.z, q.w};
		u32 cur = wd[0] & 255u, run = 0;                // runs of equal digits (sorted data) fold into one atomic
		#pragma unroll
		for (int k = 0; k < 16; k++) {
			const u32 c = (wd[k >> 2] >> ((k & 3) * 8)) & 255u;
			if (c != cur) { atomicAdd(&hw[cur], run); cur = c; run = 0; }
			run++;
		}
		atomicAdd(&hw[cur], run);
	} else {
		for (u32 k = p; k < n && k < p + 16; k++) atomicAdd(&hw[digits[k]], 1u);
	}
	__syncthreads();
	if (t < 256) {
		u32 s = 0;
		#pragma unroll
		for (int k = 0; k < RS_WARPS; k++) s += h[k][t];
		tile_hist[(size_t)t * stride + blockIdx.x] = s;
	}
}

// block d: total of digit d over all tiles (one coalesced row)
__global__ void __launch_bounds__(256) k_rs_totals(const u32* __restrict__ tile_hist, u32 tiles, u32 stride, u32* __restrict__ totals)
{
	__shared__ u32 ws[32];
	const u32 d = blockIdx.x, t = threadIdx.x;
	const u32* row = tile_hist + (size_t)d * stride;
	u32 s = 0;
	for (u32 k = t; k < tiles; k += 256) s += row[k];
	u32 total;
	block_incl_sum(s, ws, &total);
	if (t == 0) totals[d] = total;
}

// block d: row d <- global offset of each tile's first key with digit d (exclusive scan + digit base)
__global__ void __launch_bounds__(256) k_rs_scan(u32* __restrict__ tile_hist, u32 tiles, u32 stride, const u32* __restrict__ totals)
{
	__shared__ u32 ws[32];
	const u32 d = blockIdx.x, t = threadIdx.x;
	u32 carry;
	block_incl_sum(t < d ? totals[t] : 0u, ws, &carry);    // carry = sum of totals[0..d)
	u32* row = tile_hist + (size_t)d * stride;
	for (u32 base = 0; base < tiles; base += 1024) {
		const u32 k = base + t * 4;
		uint4 v = make_uint4(0, 0, 0, 0);
		if (k + 3 < tiles) v = *reinterpret_cast<const uint4*>(row + k);   // rows are padded to a multiple of 4 entries
		else {                                                           // the padding itself is never read
			if (k < tiles) v.x = row[k];
			if (k + 1 < tiles) v.y = row[k + 1];
			if (k + 2 < tiles) v.z = row[k + 2];
		}
		const u32 sum = v.x + v.y + v.z + v.w;
		u32 total;
		const u32 inc = block_incl_sum(sum, ws, &total);
		u32 run = carry + inc - sum;
		uint4 o;
		o.x = run; run += v.x; o.y = run; run += v.y; o.z = run; run += v.z; o.w = run;
		if (k < tiles) *reinterpret_cast<uint4*>(row + k) = o;
		carry += total;
	}
}

#ifndef RS_SCATTER_MIN_BLOCKS
#define RS_SCATTER_MIN_BLOCKS 2
#endif
#ifndef RS_SCATTER_MIN_BLOCKS_KEYS
#define RS_SCATTER_MIN_BLOCKS_KEYS 4          // keys-only records: no value registers, 64 registers without spills, 32 KB of tile
#endif
// PAIRS = false: the words ARE the records (a key in the high bits, the payload below bit `shift` of the first pass): no
// value arrays, two thirds of the traffic and of the shared memory.
template <typename K, bool PAIRS = true>
__global__ void __launch_bounds__(RS_THREADS, PAIRS ? RS_SCATTER_MIN_BLOCKS : RS_SCATTER_MIN_BLOCKS_KEYS) k_rs_scatter(const K* __restrict__ kin, const u32* __restrict__ vin,
                                                           K* __restrict__ kout, u32* __restrict__ vout,
                                                           const u32* __restrict__ tile_off, u32 stride, u32 n, int shift,
                                                           u8* __restrict__ dnext = nullptr, int next_shift = 0)
{
	extern __shared__ __align__(16) u8 rs_smem[];
	K* skey = reinterpret_cast<K*>(rs_smem);
	u32* sval = reinterpret_cast<u32*>(rs_smem + (size_t)RS_TILE * sizeof(K));
	__shared__ u32 wcnt[RS_WARPS][256];
	__shared__ u32 bin_start[256];
	__shared__ u32 g_off[256];
	__shared__ u32 ws[32];

	const int t = threadIdx.x, w = t >> 5, lane = t & 31;
	const u32 lt = lanemask_lt();
	const u32 tile_base = blockIdx.x * RS_TILE;
	const u32 valid = min((u32)RS_TILE, n - tile_base);
	for (int i = t; i < RS_WARPS * 256; i += RS_THREADS) (&wcnt[0][0])[i] = 0;

	K key[RS_ITEMS];
	u32 val[RS_ITEMS];
	const u32 wbase = tile_base + w * (32 * RS_ITEMS);
	#pragma unroll
	for (int i = 0; i < RS_ITEMS; i++) {
		const u32 p = wbase + i * 32 + lane;
		const bool ok = p < n;
		key[i] = ok ? kin[p] : (K)~(K)0;  // padding sorts last inside the tile and is never written out
		val[i] = (PAIRS && ok) ? vin[p] : 0u;
	}
	__syncthreads();

	u32 rank[RS_ITEMS];
	u32* mycnt = wcnt[w];
	#pragma unroll
	for (int i = 0; i < RS_ITEMS; i++) {
		const u32 d = rs_digit(key[i], shift);
		const u32 peers = match_any8_adaptive(d);
		const u32 below = __popc(peers & lt);
		u32 before = 0;
		if (below == 0) { before = mycnt[d]; mycnt[d] = before + __popc(peers); }
		before = __shfl_sync(0xffffffffu, before, __ffs(peers) - 1);
		rank[i] = before + below;
		__syncwarp();
	}
	__syncthreads();

	// digit t: exclusive scan over warps, then over digits
	u32 run = 0;
	if (t < 256) {
		#pragma unroll
		for (int k = 0; k < RS_WARPS; k++) { const u32 v = wcnt[k][t]; wcnt[k][t] = run; run += v; }
	}
	u32 total;
	const u32 inc = block_incl_sum(run, ws, &total);
	if (t < 256) {
		bin_start[t] = inc - run;
		g_off[t] = tile_off[(size_t)t * stride + blockIdx.x] - (inc - run);
	}
	__syncthreads();

	#pragma unroll
	for (int i = 0; i < RS_ITEMS; i++) {
		const u32 d = rs_digit(key[i], shift);
		const u32 pos = bin_start[d] + mycnt[d] + rank[i];
		skey[pos] = key[i];
		if (PAIRS) sval[pos] = val[i];
	}
	__syncthreads();

	for (u32 j = t; j < valid; j += RS_THREADS) {
		const K k = skey[j];
		const u32 dst = g_off[rs_digit(k, shift)] + j;
		kout[dst] = k;
		if (PAIRS) vout[dst] = sval[j];
		if (dnext) dnext[dst] = (u8)rs_digit(k, next_shift);   // the next pass counts this byte instead of re-reading the key
	}
}

struct RadixBuffers {
	u64* k[2]; u32* v[2];
	u32* tile_hist;   // 256 rows of rs_stride(ceil(n / RS_TILE)) entries
	u32* totals;      // 256
	u8*  dnext;       // n bytes (or null): digit of the next pass, written by the scatter, read by k_rs_hist_bytes
	int* err;
};

inline size_t radix_tiles(size_t n) { return (n + RS_TILE - 1) / RS_TILE; }

// Sorts pairs in (k[cur], v[cur]) by key bits [bit_lo, bit_hi); returns the index (0/1) of the buffers
// holding the result, or a negative error code. launches is incremented per kernel.
// first_hist_ready: b.tile_hist already holds the tile histogram of the first digit (the producer of the keys made it)
inline int radix_sort_pairs(RadixBuffers& b, int cur, u32 n, int bit_lo, int bit_hi, cudaStream_t s, int* launches,
                            bool first_hist_ready = false, bool keys_only = false)
{
	if (n == 0) return cur;
	if (keys_only) {
		// the words carry their payload in the bits below bit_lo: classic passes without the value arrays
		const size_t smem = (size_t)RS_TILE * 8;
		if (cudaFuncSetAttribute(k_rs_scatter<u64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
		const u32 tiles = (u32)radix_tiles(n), stride = rs_stride(tiles);
		for (int shift = bit_lo; shift < bit_hi; shift += 8) {
			if (shift == bit_lo) {
				if (!first_hist_ready) { k_rs_hist<u64><<<tiles, RS_THREADS, 0, s>>>(b.k[cur], n, shift, b.tile_hist, stride); *launches += 1; }
			} else if (b.dnext) { k_rs_hist_bytes<<<tiles, RS_THREADS, 0, s>>>(b.dnext, n, b.tile_hist, stride); *launches += 1; }
			else { k_rs_hist<u64><<<tiles, RS_THREADS, 0, s>>>(b.k[cur], n, shift, b.tile_hist, stride); *launches += 1; }
			k_rs_totals<<<256, 256, 0, s>>>(b.tile_hist, tiles, stride, b.totals);
			k_rs_scan<<<256, 256, 0, s>>>(b.tile_hist, tiles, stride, b.totals);
			const bool more = shift + 8 < bit_hi;
			k_rs_scatter<u64, false><<<tiles, RS_THREADS, smem, s>>>(b.k[cur], nullptr, b.k[cur ^ 1], nullptr, b.tile_hist, stride, n, shift,
			                                                         more ? b.dnext : nullptr, shift + 8);
			*launches += 3;
			cur ^= 1;
		}
		return cur;
	}
	// function attributes are per device; setting one is a host-only call
	if (cudaFuncSetAttribute(k_rs_scatter<u64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS_SMEM_SCATTER) != cudaSuccess) return -1;
	const u32 tiles = (u32)radix_tiles(n), stride = rs_stride(tiles);
	for (int shift = bit_lo; shift < bit_hi; shift += 8) {
		if (shift == bit_lo) {
			if (!first_hist_ready) { k_rs_hist<u64><<<tiles, RS_THREADS, 0, s>>>(b.k[cur], n, shift, b.tile_hist, stride); *launches += 1; }
		} else if (b.dnext) { k_rs_hist_bytes<<<tiles, RS_THREADS, 0, s>>>(b.dnext, n, b.tile_hist, stride); *launches += 1; }
		else { k_rs_hist<u64><<<tiles, RS_THREADS, 0, s>>>(b.k[cur], n, shift, b.tile_hist, stride); *launches += 1; }
		k_rs_totals<<<256, 256, 0, s>>>(b.tile_hist, tiles, stride, b.totals);
		k_rs_scan<<<256, 256, 0, s>>>(b.tile_hist, tiles, stride, b.totals);
		const bool more = shift + 8 < bit_hi;
		k_rs_scatter<u64><<<tiles, RS_THREADS, RS_SMEM_SCATTER, s>>>(b.k[cur], b.v[cur], b.k[cur ^ 1], b.v[cur ^ 1], b.tile_hist, stride, n, shift,
		                                                              more ? b.dnext : nullptr, shift + 8);
		*launches += 3;
		cur ^= 1;
	}
	return cur;
}

// LSD sort of (32-bit key, 32-bit value) pairs by key bits [0, bits): k[cur], v[cur] -> returns the index of the buffers
// holding the result (or a negative error code). Used where ranks, not text prefixes, are the keys.
inline int radix_sort_pairs32(u32* const k[2], u32* const v[2], int cur, u32 n, int bits, u32* tile_hist, u32* totals, cudaStream_t s, int* launches)
{
	if (n == 0) return cur;
	const size_t smem = (size_t)RS_TILE * (4 + 4);
	if (cudaFuncSetAttribute(k_rs_scatter<u32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
	const u32 tiles = (u32)radix_tiles(n), stride = rs_stride(tiles);
	for (int shift = 0; shift < bits; shift += 8) {
		k_rs_hist<u32><<<tiles, RS_THREADS, 0, s>>>(k[cur], n, shift, tile_hist, stride);
		k_rs_totals<<<256, 256, 0, s>>>(tile_hist, tiles, stride, totals);
		k_rs_scan<<<256, 256, 0, s>>>(tile_hist, tiles, stride, totals);
		k_rs_scatter<u32><<<tiles, RS_THREADS, smem, s>>>(k[cur], v[cur], k[cur ^ 1], v[cur ^ 1], tile_hist, stride, n, shift);
		*launches += 4;
		cur ^= 1;
	}
	return cur;
}

// One stable 8-bit partition pass over (32-bit key, 32-bit value) pairs: (kin, vin) -> (kout, vout) ordered by
// digit (key >> shift) & 255. Used to bucket (suffix, rank) pairs by ISA region before they are scattered.
inline int radix_partition_u32(const u32* kin, const u32* vin, u32* kout, u32* vout, u32 n, int shift,
                               u32* tile_hist, u32* totals, cudaStream_t s, int* launches)
{
	if (n == 0) return 0;
	const size_t smem = (size_t)RS_TILE * (4 + 4);
	if (cudaFuncSetAttribute(k_rs_scatter<u32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
	const u32 tiles = (u32)radix_tiles(n), stride = rs_stride(tiles);
	k_rs_hist<u32><<<tiles, RS_THREADS, 0, s>>>(kin, n, shift, tile_hist, stride);
	k_rs_totals<<<256, 256, 0, s>>>(tile_hist, tiles, stride, totals);
	k_rs_scan<<<256, 256, 0, s>>>(tile_hist, tiles, stride, totals);
	k_rs_scatter<u32><<<tiles, RS_THREADS, smem, s>>>(kin, vin, kout, vout, tile_hist, stride, n, shift);
	*launches += 4;
	return 0;
}

} // namespace jp
