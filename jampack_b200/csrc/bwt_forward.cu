// bwt_forward.cu -- forward BWT for sm_100a: GPU suffix sort + BWT emission + the 120 sampled indices.
//
// Replaces BlockSort::Bwt::ForwardBwt (reference bwt.cpp:22-65) and the divsufsort() call under it
// (divsufsort.cpp:1721; contract divsufsort.hpp:37-45: plain suffix array, a proper prefix sorts first).
// Nothing of divsufsort's induced-copying design is kept: a serial induce pass has no place on 148 SMs.
//
//   1. symbol remap   present byte values -> dense codes 1..sigma (0 = end of string), b = ceil(log2(sigma+1)) bits
//   2. initial keys   key(i) = the first d = floor(64/b) codes of suffix i, zero padded: the padding IS the
//                     sentinel, so "shorter sorts first" needs no tie-break and no suffix shorter than the
//                     current depth is ever in a group with another one
//   3. radix bucket   LSD radix sort of (key, i)                                      [radix_sort.cuh]
//   4. ranks          group heads = key changes; rank = 1 + position of the group's head; singletons retire
//                     into SA at once, the rest are compacted into the active set
//   5. doubling       while any group is unsorted: key = (dense group id, ISA[s + h]) -> radix sort ->
//                     heads/ranks/retire/compact; h doubles. Only the active set is touched (Larsson-Sadakane
//                     discarding); ISA[nlen] = 0 is the empty suffix.
//   6. emit           bwt[o] = T[SA[row]-1] with the row of suffix 0 skipped (bwt.cpp:50-56), the sampled
//                     indices ISA[k*step] (bwt.cpp:44-48,57-61), the raw tail (bwt.cpp:32-33).
#include "bwt_internal.cuh"
#include "radix_sort.cuh"
#include <chrono>

namespace jp {

struct FwdMeta {
	u32 code[256];
	i32 sigma, bits, depth, key_bits;   // bits: per symbol (reported); depth: symbols per key; key_bits: bit length of the largest key
	u32 hist[256];
};

// ---- 1. symbol histogram and dense codes -------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fwd_symhist(const u8* __restrict__ T, i32 n, FwdMeta* __restrict__ meta)
{
	__shared__ u32 h[8][256];
	const int t = threadIdx.x, w = t >> 5;
	for (int i = t; i < 8 * 256; i += 256) (&h[0][0])[i] = 0;
	__syncthreads();
	const i64 stride = (i64)gridDim.x * 256 * 16;
	for (i64 p = ((i64)blockIdx.x * 256 + t) * 16; p < n; p += stride) {
		if (p + 16 <= n) {
			const uint4 q = __ldg(reinterpret_cast<const uint4*>(T + p));
			const u32 wd[4] = {q.x, q.y, q.z, q.w};
			#pragma unroll
			for (int k = 0; k < 16; k++) atomicAdd(&h[w][(wd[k >> 2] >> ((k & 3) * 8)) & 255], 1u);
		} else for (i64 q = p; q < n; q++) atomicAdd(&h[w][T[q]], 1u);
	}
	__syncthreads();
	u32 s = 0;
	#pragma unroll
	for (int k = 0; k < 8; k++) s += h[k][t];
	if (s) atomicAdd(&meta->hist[t], s);
}

__global__ void __launch_bounds__(256) k_fwd_codes(FwdMeta* __restrict__ meta)
{
	__shared__ u32 ws[32];
	const int t = threadIdx.x;
	const u32 present = meta->hist[t] ? 1u : 0u;
	u32 total;
	const u32 inc = block_incl_sum(present, ws, &total);
	meta->code[t] = present ? inc : 0u;            // codes 1..sigma in byte order
	if (t == 0) {
		// Keys are mixed-radix numbers in base sigma+1 (digit 0 = end of string): as many symbols as fit below 2^63.
		// Bit fields would waste the gap between sigma+1 and the next power of two -- 65 code values in 7-bit fields
		// give 9 symbols per key, base 65 gives 10 (and 27 instead of 21 for a 4-symbol block). Lexicographic order of
		// the symbol strings is the numeric order of the keys either way.
		const u64 base = (u64)total + 1;
		int depth = 0; u64 span = 1;                // span = base^depth
		while (depth < 63 && span <= (((u64)1 << 63) - 1) / base) { span *= base; depth++; }
		meta->sigma = (i32)total;
		meta->bits = bit_length((u64)total);
		meta->depth = depth;
		meta->key_bits = bit_length(span - 1);
	}
}

// ---- 2. initial keys -------------------------------------------------------------------------------
// One block builds the keys of one radix tile (RS_TILE positions) and, having them in hand, also counts the
// lowest digit: the first radix pass starts from this tile histogram instead of re-reading the keys.
constexpr int KEY_TILE = RS_TILE;
__global__ void __launch_bounds__(256) k_fwd_keys(const u8* __restrict__ T, i32 n, const FwdMeta* __restrict__ meta,
                                                  u64* __restrict__ keys, u32* __restrict__ vals,
                                                  u32* __restrict__ tile_hist, u32 stride)
{
	__shared__ u16 sc[KEY_TILE + 64];
	__shared__ u32 part[KEY_TILE + 32];       // value of the q = depth/2 symbols starting at each position (fits 32 bits)
	__shared__ u16 code[256];
	__shared__ u32 h[8][256];
	const int t = threadIdx.x, w = t >> 5;
	code[t] = (u16)meta->code[t];
	for (int i = t; i < 8 * 256; i += 256) (&h[0][0])[i] = 0;
	const int depth = meta->depth, q = depth >> 1;
	const u32 radix = (u32)meta->sigma + 1;
	__syncthreads();
	const i64 base = (i64)blockIdx.x * KEY_TILE;
	for (int i = t; i < KEY_TILE + 64; i += 256) {
		const i64 p = base + i;
		sc[i] = p < n ? code[T[p]] : (u16)0;
	}
	__syncthreads();
	// key(i) = part(i) * radix^(depth-q) + [part(i+q) or part(i+q) * radix + symbol(i+2q)]: q 32-bit multiply-adds per
	// position and one or two 64-bit ones, instead of `depth` 64-bit multiply-adds
	for (int i = t; i < KEY_TILE + 32; i += 256) {
		u32 v = 0;
		for (int d = 0; d < q; d++) v = v * radix + sc[i + d];
		part[i] = v;
	}
	u64 hi_scale = 1;
	for (int d = 0; d < depth - q; d++) hi_scale *= radix;
	__syncthreads();
	#pragma unroll 4
	for (int j = 0; j < KEY_TILE / 256; j++) {
		const int li = j * 256 + t;
		const i64 p = base + li;
		if (p < n) {
			u64 tail = part[li + q];
			if (depth & 1) tail = tail * radix + sc[li + 2 * q];
			const u64 k = (u64)part[li] * hi_scale + tail;
			keys[p] = k;
			vals[p] = (u32)p;
			atomicAdd(&h[w][(u32)k & 255u], 1u);           // low digits of text-order keys are spread: lanes rarely collide
		}
	}
	__syncthreads();
	u32 sum = 0;
	#pragma unroll
	for (int k = 0; k < 8; k++) sum += h[k][t];
	tile_hist[(size_t)t * stride + blockIdx.x] = sum;
}

// ---- 4/5. group heads -> ranks, retire singletons, compact the rest ----------------------------------
// Scan element: (P of the last group head so far, #survivors, #surviving group heads).
constexpr int GS_THREADS = 256;
constexpr int GS_SUB     = 8;
constexpr int GS_TILE    = GS_THREADS * GS_SUB;
struct GAgg { i32 mh; u32 ns; u32 ng; u32 pad; };

// Group boundaries come either from the sorted keys (a key change) or, after the shared-memory segmented sort,
// from the one-byte head flags it wrote.
struct GFlags { bool head, nhead; };
__device__ __forceinline__ GFlags group_flags(const u64* __restrict__ K, const u8* __restrict__ F, u32 j, u32 A)
{
	GFlags f;
	if (F) {
		f.head = F[j] != 0;
		f.nhead = (j + 1 == A) || (F[j + 1] != 0);
	} else {
		const u64 kj = K[j];
		f.head = (j == 0) || (K[j - 1] != kj);
		f.nhead = (j + 1 == A) || (K[j + 1] != kj);
	}
	return f;
}

__global__ void __launch_bounds__(GS_THREADS) k_grp_reduce(const u64* __restrict__ K, const u8* __restrict__ F,
                                                           const u32* __restrict__ P, u32 A, GAgg* __restrict__ agg)
{
	__shared__ i32 smh[8];
	__shared__ u32 sns[8], sng[8];
	const int t = threadIdx.x;
	const u32 base = blockIdx.x * GS_TILE;
	i32 mh = -1; u32 ns = 0, ng = 0;
	#pragma unroll 4
	for (int s = 0; s < GS_SUB; s++) {
		const u32 j = base + s * GS_THREADS + t;
		if (j < A) {
			const GFlags f = group_flags(K, F, j, A);
			if (f.head) mh = max(mh, (i32)(P ? P[j] : j));
			ns += !(f.head && f.nhead);
			ng += (f.head && !f.nhead);
		}
	}
	mh = warp_max(mh); ns = warp_sum(ns); ng = warp_sum(ng);
	if ((t & 31) == 0) { smh[t >> 5] = mh; sns[t >> 5] = ns; sng[t >> 5] = ng; }
	__syncthreads();
	if (t == 0) {
		for (int k = 1; k < 8; k++) { mh = max(mh, smh[k]); ns += sns[k]; ng += sng[k]; }
		GAgg a; a.mh = mh; a.ns = ns; a.ng = ng; a.pad = 0;
		agg[blockIdx.x] = a;
	}
}

// single block: exclusive scan of the tile aggregates in place; totals -> out[0] = survivors, out[1] = groups.
// Each warp owns a contiguous range and walks it 32 aggregates at a time (coalesced 512 B rows, one warp scan per
// row); the 32 warp totals are stitched through shared memory.
__device__ __forceinline__ GAgg gagg_combine(const GAgg& a, const GAgg& b) { GAgg r; r.mh = max(a.mh, b.mh); r.ns = a.ns + b.ns; r.ng = a.ng + b.ng; r.pad = 0; return r; }
__device__ __forceinline__ GAgg gagg_shfl_up(const GAgg& a, int o)
{
	GAgg r; r.mh = __shfl_up_sync(0xffffffffu, a.mh, o); r.ns = __shfl_up_sync(0xffffffffu, a.ns, o); r.ng = __shfl_up_sync(0xffffffffu, a.ng, o); r.pad = 0; return r;
}
__device__ __forceinline__ GAgg gagg_warp_incl(GAgg v)
{
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const GAgg t = gagg_shfl_up(v, o); if ((int)lane_id() >= o) v = gagg_combine(t, v); }
	return v;
}
__global__ void __launch_bounds__(1024) k_grp_scan_tiles(GAgg* __restrict__ agg, int tiles, u32* __restrict__ out)
{
	__shared__ GAgg wtot[32];
	const int t = threadIdx.x, lane = t & 31, w = t >> 5;
	const int rows = (tiles + 31) / 32;                  // rows of 32 aggregates
	const int rows_per_warp = (rows + 31) / 32;
	const int r0 = min(rows, w * rows_per_warp), r1 = min(rows, r0 + rows_per_warp);
	GAgg ident; ident.mh = -1; ident.ns = 0; ident.ng = 0; ident.pad = 0;
	GAgg acc = ident;
	for (int r = r0; r < r1; r++) { const int k = r * 32 + lane; if (k < tiles) acc = gagg_combine(acc, agg[k]); }
	acc = gagg_warp_incl(acc);
	if (lane == 31) wtot[w] = acc;
	__syncthreads();
	GAgg carry = ident;
	for (int k = 0; k < w; k++) carry = gagg_combine(carry, wtot[k]);
	for (int r = r0; r < r1; r++) {
		const int k = r * 32 + lane;
		const GAgg v = k < tiles ? agg[k] : ident;
		const GAgg inc = gagg_warp_incl(v);
		GAgg ex = gagg_shfl_up(inc, 1);
		if (lane == 0) ex = ident;
		ex = gagg_combine(carry, ex);
		if (k < tiles) agg[k] = ex;
		GAgg rowtot; rowtot.mh = __shfl_sync(0xffffffffu, inc.mh, 31); rowtot.ns = __shfl_sync(0xffffffffu, inc.ns, 31); rowtot.ng = __shfl_sync(0xffffffffu, inc.ng, 31); rowtot.pad = 0;
		carry = gagg_combine(carry, rowtot);
	}
	if (t == 1023) {
		GAgg all = ident;
		for (int k = 0; k < 32; k++) all = gagg_combine(all, wtot[k]);
		out[0] = all.ns; out[1] = all.ng;
	}
}

// All GS_SUB sub-tiles of a tile are loaded up front (GS_SUB independent loads per array in flight), scanned
// inside each warp, and stitched together with ONE barrier through a [sub-tile][warp] table -- the earlier
// one-sub-tile-at-a-time version exposed a full load latency and two barriers per 256 elements (ncu: 93
// long-scoreboard + 50 barrier stall cycles per issue).
__global__ void __launch_bounds__(GS_THREADS) k_grp_apply(const u64* __restrict__ K, const u8* __restrict__ F, const u32* __restrict__ V,
                                                          const u32* __restrict__ P, u32 A, const GAgg* __restrict__ agg,
                                                          int rank_bits, u32* __restrict__ ISA, u32* __restrict__ R, u32* __restrict__ SA,
                                                          u64* __restrict__ Kn, u32* __restrict__ Vn, u32* __restrict__ Pn)
{
	__shared__ i32 wmh[GS_SUB][8];
	__shared__ u32 wpk[GS_SUB][8];
	const int t = threadIdx.x, lane = t & 31, w = t >> 5;
	const u32 base = blockIdx.x * GS_TILE;
	const GAgg carry0 = agg[blockIdx.x];

	u32 v[GS_SUB], p[GS_SUB], fl[GS_SUB];     // fl: bit0 valid, bit1 head, bit2 survivor, bit3 surviving head
	#pragma unroll
	for (int s = 0; s < GS_SUB; s++) {
		const u32 j = base + s * GS_THREADS + t;
		v[s] = 0; p[s] = 0; fl[s] = 0;
		if (j < A) {
			const GFlags f = group_flags(K, F, j, A);
			fl[s] = 1u | (f.head ? 2u : 0u) | (!(f.head && f.nhead) ? 4u : 0u) | ((f.head && !f.nhead) ? 8u : 0u);
			v[s] = V[j]; p[s] = P ? P[j] : j;
		}
	}
	i32 imh[GS_SUB]; u32 ipk[GS_SUB];
	#pragma unroll
	for (int s = 0; s < GS_SUB; s++) {
		const i32 mh = (fl[s] & 2u) ? (i32)p[s] : -1;
		const u32 packed = ((fl[s] >> 2) & 1u) | (((fl[s] >> 3) & 1u) << 16);
		imh[s] = warp_incl_max(mh);
		ipk[s] = warp_incl_sum(packed);
		if (lane == 31) { wmh[s][w] = imh[s]; wpk[s][w] = ipk[s]; }
	}
	__syncthreads();
	i32 run_mh = carry0.mh; u32 run_pk = 0;
	#pragma unroll
	for (int s = 0; s < GS_SUB; s++) {
		i32 pm = run_mh, tm = run_mh; u32 pp = run_pk, tp = run_pk;
		#pragma unroll
		for (int k = 0; k < 8; k++) {
			const i32 a = wmh[s][k]; const u32 b = wpk[s][k];
			if (k < w) { pm = max(pm, a); pp += b; }
			tm = max(tm, a); tp += b;
		}
		if (fl[s] & 1u) {
			const i32 fmh = max(imh[s], pm);
			const u32 fpk = ipk[s] + pp;
			if (R) R[base + s * GS_THREADS + t] = (u32)fmh + 1u;   // ranks leave in slot order; k_isa_scatter places them
			else ISA[v[s]] = (u32)fmh + 1u;
			if (fl[s] & 4u) {
				const u32 q = carry0.ns + (fpk & 0xffffu) - 1;
				Vn[q] = v[s]; Pn[q] = p[s]; Kn[q] = (u64)(carry0.ng + (fpk >> 16) - 1) << rank_bits;
			} else SA[p[s]] = v[s];
		}
		run_mh = tm; run_pk = tp;
	}
}

// ISA[V[j]] = R[j] is a random 4-byte scatter; done naively every store dirties one sector that DRAM later has to
// read-modify-write (ncu: 2.7 ms for 64 M ranks, 24 G stores/s against 72 G/s for gathers). Instead the slots are
// streamed once per REGION of ISA (2^region_log2 entries, sized to stay L2-resident): a pass only stores the ranks
// that fall in its region, so the sectors fill up in L2 and go to DRAM once, complete. Blocks are ordered by
// region, so the passes follow each other inside one launch.
__global__ void __launch_bounds__(256) k_isa_scatter(const u32* __restrict__ V, const u32* __restrict__ R, u32 A, u32* __restrict__ ISA,
                                                     int region_log2, u32 tiles)
{
	const u32 region = blockIdx.x / tiles, tile = blockIdx.x % tiles;
	const u32 base = tile * 2048 + threadIdx.x;
	u32 v[8], r[8];
	#pragma unroll
	for (int i = 0; i < 8; i++) {
		const u32 j = base + i * 256;
		v[i] = 0xffffffffu; r[i] = 0;
		if (j < A) { v[i] = __ldcs(V + j); r[i] = __ldcs(R + j); }
	}
	#pragma unroll
	for (int i = 0; i < 8; i++) if (v[i] != 0xffffffffu && (region_log2 >= 32 || (v[i] >> region_log2) == region)) ISA[v[i]] = r[i];
}

// ---- 4b. run skip (JP_BWT_FWD_RUNSKIP=1; off by default until measured) --------------------------------
// Prefix doubling pays one round per doubling of the longest repeat, and the cheapest way to make a long repeat is a
// run of one symbol (zero pages, padding): a block of n equal bytes takes log2(n / depth) rounds over the whole block.
// Runs have an exact shortcut. Let r(v) be the length of the run of T[v] that starts at v, and call v a RUN SUFFIX
// when r(v) >= depth, the number of symbols in the initial key: its key is c^depth, so after the initial sort the run
// suffixes of a symbol c form one group, and inside it suffix v = c^r(v) X(v), where X(v) starts with a symbol other
// than c (or is empty). Comparing c^r X with c^r' X', r < r', is decided at position r: X's first symbol against c.
// Hence the order inside the group is: first the suffixes whose run is followed by a SMALLER symbol (or by the end of
// the text), by ascending r; then those followed by a larger one, by descending r; ties (equal class and r -- they
// come from different runs) by the order of the X's. So
//   * pass A (the first round, h = depth): a run suffix takes key2 = r (first class) or 2n + 1 - r (second class)
//     instead of ISA[v + h] -- a block of equal bytes is sorted by that one pass -- which leaves its group with equal
//     (class, r);
//   * pass B (h still = depth) and every later round: a run suffix with r(v) >= h takes ISA[v + r(v)], the rank of
//     X(v); one with r(v) < h the ordinary ISA[v + h]. Either way the key is uniform inside a group (its members have
//     the same r) and the round leaves every suffix at least 2h-ordered, which is the invariant the ordinary keys of
//     the NEXT round rely on when they read the rank of a position inside a run: c^r X keyed on an h-ordered rank of X
//     is (r + h)-ordered, and r >= h. (Keying on ISA[v + r] regardless of h looks tempting and is wrong: a run suffix
//     with r < h would then be less refined than its neighbours assume. Found by the emulated tests.)
// All other suffixes are untouched. key2 needs one more bit (values up to 2n), which the host accounts for in
// rank_bits. RL = r(v) for every position, `bits` = bitmap of the run suffixes (read through L2 by the gathers, so
// only run suffixes pay for the RL gather).
struct RunSkip { const u32* rl; const u32* bits; const u8* T; u32 first; };

template <bool RS>
__device__ __forceinline__ u32 key2_of(u32 v, u32 h, u32 n, const u32* __restrict__ ISA, const RunSkip& rs, int* __restrict__ err)
{
	if (RS) {
		if ((__ldg(&rs.bits[v >> 5]) >> (v & 31)) & 1u) {
			const u32 r = __ldg(&rs.rl[v]);                    // v + r <= n by construction
			if (rs.first) {
				const bool smaller_follows = (v + r >= n) || rs.T[v + r] < rs.T[v];
				return smaller_follows ? r : 2u * n + 1u - r;
			}
			if (r >= h) return __ldg(&ISA[v + r]);
		}
	}
	u32 p = v + h;
	if (p > n) { dev_fail(err, DE_FWD_RANGE); p = n; }
	return __ldg(&ISA[p]);
}

constexpr int RUN_TILE = 2048;
constexpr u32 RUN_NONE = 0xffffffffu;
__device__ __forceinline__ u32 warp_min_u32(u32 v)
{
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}

// last position of the first run that ENDS inside the tile (T[j] != T[j+1], or j == n-1); RUN_NONE if none does
__global__ void __launch_bounds__(256) k_run_first(const u8* __restrict__ T, u32 n, u32* __restrict__ tile_first)
{
	__shared__ u32 wbest[8];
	const u32 base = blockIdx.x * RUN_TILE;
	u32 best = RUN_NONE;
	#pragma unroll
	for (int i = 0; i < RUN_TILE / 256; i++) {
		const u32 j = base + i * 256 + threadIdx.x;
		if (j < n && (j == n - 1 || T[j] != T[j + 1])) best = min(best, j);
	}
	best = warp_min_u32(best);
	if ((threadIdx.x & 31) == 0) wbest[threadIdx.x >> 5] = best;
	__syncthreads();
	if (threadIdx.x == 0) {
		#pragma unroll
		for (int k = 1; k < 8; k++) best = min(best, wbest[k]);
		tile_first[blockIdx.x] = best;
	}
}

// single block: tile_next[t] = first run end in any LATER tile (exclusive suffix minimum)
__global__ void __launch_bounds__(1024) k_run_scan(const u32* __restrict__ tile_first, u32 tiles, u32* __restrict__ tile_next)
{
	__shared__ u32 part[1024];
	const u32 t = threadIdx.x;
	const u32 per = (tiles + 1023) / 1024;
	const u32 lo = min(tiles, t * per), hi = min(tiles, lo + per);
	u32 m = RUN_NONE;
	for (u32 k = lo; k < hi; k++) m = min(m, tile_first[k]);
	part[t] = m;
	__syncthreads();
	u32 run = RUN_NONE;
	for (u32 k = t + 1; k < 1024; k++) run = min(run, part[k]);
	for (u32 k = hi; k > lo; k--) { tile_next[k - 1] = run; run = min(run, tile_first[k - 1]); }
}

// RL[j] = length of the run of T[j] starting at j; bits = (RL[j] >= depth); *count += number of run suffixes
__global__ void __launch_bounds__(256) k_run_fill(const u8* __restrict__ T, u32 n, const u32* __restrict__ tile_next, u32 depth,
                                                  u32* __restrict__ RL, u32* __restrict__ bits, u32* __restrict__ count)
{
	__shared__ u32 end_at[RUN_TILE];            // first run end at or after each position of the tile
	__shared__ u32 wfirst[8];
	const u32 t = threadIdx.x, lane = t & 31, w = t >> 5;
	const u32 base = blockIdx.x * RUN_TILE;
	// each thread owns 8 consecutive positions: its first run end, if any
	u32 mine = RUN_NONE;
	#pragma unroll
	for (int k = 7; k >= 0; k--) {
		const u32 j = base + t * 8 + k;
		if (j < n && (j == n - 1 || T[j] != T[j + 1])) mine = j;
	}
	// exclusive suffix minimum over the threads: inside the warp by shuffles, across the 8 warps through shared memory
	u32 incl = mine;
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const u32 x = __shfl_down_sync(0xffffffffu, incl, o); if (lane + o < 32) incl = min(incl, x); }
	u32 excl = __shfl_down_sync(0xffffffffu, incl, 1);
	if (lane == 31) excl = RUN_NONE;
	if (lane == 0) wfirst[w] = incl;
	__syncthreads();
	u32 carry = min(tile_next[blockIdx.x], excl);
	for (u32 k = 7; k > w; k--) carry = min(carry, wfirst[k]);
	#pragma unroll
	for (int k = 7; k >= 0; k--) {
		const u32 j = base + t * 8 + k;
		if (j < n && (j == n - 1 || T[j] != T[j + 1])) carry = j;
		end_at[t * 8 + k] = carry;
	}
	__syncthreads();
	u32 cnt = 0;
	#pragma unroll
	for (int i = 0; i < RUN_TILE / 256; i++) {
		const u32 j = base + i * 256 + t;
		u32 r = 0;
		if (j < n) { r = end_at[i * 256 + t] - j + 1; RL[j] = r; }
		const u32 m = __ballot_sync(0xffffffffu, j < n && r >= depth);
		if (lane == 0 && base + i * 256 + (t & ~31u) < n) { bits[(base + i * 256 + t) >> 5] = m; cnt += __popc(m); }
	}
	if (lane == 0 && cnt) atomicAdd(count, cnt);
}

// ---- 5a. doubling round, small groups: gather + segmented sort in shared memory -------------------------
// Groups are contiguous in the active set, so refining them is a SEGMENTED sort. Block c owns the groups whose
// head lies in its window of SG_WIN slots; they end within two windows unless the last one is "large". The block
// gathers key2 = ISA[s + h] for its (<= 4096) elements, sorts (local group id, key2) with an LSD radix sort that
// never leaves shared memory (warp match.any ranking, same scheme as k_rs_scatter), writes the suffix ids back in
// place and one head flag per element. One global read and one global write of 4 bytes per active suffix
// replace the 7 global radix passes over 12-byte pairs of the composite-key route.
constexpr int SG_THREADS = 256;
constexpr int SG_ITEMS   = 16;
constexpr int SG_CAP     = SG_THREADS * SG_ITEMS;   // 4096 elements sorted per block
constexpr int SG_WIN     = SG_CAP / 2;              // window of group heads per block
constexpr size_t SG_SMEM = (size_t)SG_CAP * (8 + 4);
constexpr u32 SG_PAIR_MAX = 1024;                    // longest group the all-pairs rank refinement takes on

struct SegTile { u32 start, len; };

// Slot range [start, start+len) of the groups whose head lies in window `win` (len == 0: nothing to do).
// When `win_first` is given, the block also records its first head slot and the head of the group it had to
// leave to the large-group route (0xffffffff = none); k_large_collect turns those into the list of large groups.
__device__ __forceinline__ SegTile seg_range(u32 win, const u64* __restrict__ K, u32 A, int rank_bits,
                                             u32* __restrict__ sh /*[3]*/, u32* __restrict__ has_large,
                                             u32* __restrict__ win_first = nullptr, u32* __restrict__ win_large = nullptr)
{
	const int t = threadIdx.x;
	const u32 w0 = win * SG_WIN;
	const u32 L = min(w0 + (u32)SG_WIN, A);
	SegTile r; r.start = 0; r.len = 0;
	if (t == 0) { sh[0] = 0xffffffffu; sh[1] = 0; sh[2] = 0xffffffffu; }
	__syncthreads();
	#pragma unroll
	for (int i = 0; i < SG_WIN / SG_THREADS; i++) {
		const u32 j = w0 + i * SG_THREADS + t;
		if (j < L) {
			const u64 g = K[j] >> rank_bits;
			if (j == 0 || (K[j - 1] >> rank_bits) != g) { atomicMin(&sh[0], j); atomicMax(&sh[1], j); }
		}
	}
	__syncthreads();
	const u32 start = sh[0];
	if (t == 0 && win_first) { win_first[win] = start; win_large[win] = 0xffffffffu; }
	if (start == 0xffffffffu) return r;               // the window lies inside a group owned by an earlier block
	u32 end;
	if (L == A) end = A;
	else {
		const u64 gl = K[L - 1] >> rank_bits;
		const u32 lim = min(w0 + 2u * SG_WIN, A);
		#pragma unroll
		for (int i = 0; i < SG_WIN / SG_THREADS; i++) {
			const u32 j = L + i * SG_THREADS + t;
			if (j < lim && (K[j] >> rank_bits) != gl) atomicMin(&sh[2], j);
		}
		__syncthreads();
		end = sh[2];
		if (end == 0xffffffffu) {
			if (lim == A) end = A;
			else {                                        // last group spans > a window: not ours
				end = sh[1];
				if (t == 0 && has_large) { *has_large = 1u; if (win_large) win_large[win] = end; }
			}
		}
	}
	if (end <= start) return r;
	r.start = start; r.len = end - start;             // <= SG_CAP by construction
	return r;
}

__device__ __forceinline__ u32 warp_rev_incl_min(u32 v)
{
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_down_sync(0xffffffffu, v, o); if (lane_id() + o < 32) v = min(v, t); }
	return v;
}

// Light kernel: tiles whose groups are all short are finished here by warp-level rank refinement; the others
// are queued for k_seg_sort_radix. Every element learns its group's slot range [gs, ge) from two warp-level
// scans (last head at or before me / first group end at or after me) stitched across the block through a small
// table with a single barrier; then it counts the members of its group that sort before it.
constexpr size_t SG_SMEM_LIGHT = (size_t)SG_CAP * (4 + 4 + 1);
template <bool RS>
__global__ void __launch_bounds__(SG_THREADS, 5) k_seg_sort(const u64* __restrict__ K, u32* __restrict__ V, u32 A,
                                                         const u32* __restrict__ ISA, u32 h, u32 n, int rank_bits,
                                                         u8* __restrict__ F, u32* __restrict__ counters /*[2]=has_large [3]=queued*/,
                                                         u32* __restrict__ queue, u32* __restrict__ win_first, u32* __restrict__ win_large,
                                                         int* __restrict__ err, RunSkip rs)
{
	extern __shared__ __align__(16) u8 sg_smem[];
	u32* skey = reinterpret_cast<u32*>(sg_smem);        // key2 = ISA[s + h]; later the refined suffix ids
	u32* sval = skey + SG_CAP;                          // suffix ids in slot order
	u8* sflag = reinterpret_cast<u8*>(sval + SG_CAP);   // head flags of the refined order
	__shared__ u32 sh[3];
	__shared__ u32 wf[SG_ITEMS][SG_THREADS / 32], wb[SG_ITEMS][SG_THREADS / 32];
	const int t = threadIdx.x, lane = t & 31, w = t >> 5;
	const SegTile tile = seg_range(blockIdx.x, K, A, rank_bits, sh, counters + 2, win_first, win_large);
	const u32 len = tile.len, start = tile.start;
	if (len == 0) return;

	// per element, packed: bits 0-12 last head at or before me (-> group start), bits 13-25 first group end at or
	// after me, bit 26 head, bit 27 last of its group, bit 28 valid
	u32 pk[SG_ITEMS];
	#pragma unroll 4
	for (int i = 0; i < SG_ITEMS; i++) {
		const u32 e = i * SG_THREADS + t;
		u32 hl = 0;
		if (e < len) {
			const u32 j = start + e;
			const u32 v = V[j];
			const u64 g = K[j] >> rank_bits;
			const bool head = (e == 0) || (K[j - 1] >> rank_bits) != g;
			const bool last = (e == len - 1) || (K[j + 1] >> rank_bits) != g;
			sval[e] = v;
			skey[e] = key2_of<RS>(v, h, n, ISA, rs, err);
			hl = (head ? 1u : 0u) | (last ? 2u : 0u) | 4u;
		}
		const u32 f = warp_incl_max((hl & 1u) ? (i32)e : 0);
		const u32 b = warp_rev_incl_min((hl & 2u) ? e + 1 : 0x1fffu);
		if (lane == 31) wf[i][w] = f;
		if (lane == 0) wb[i][w] = b;
		pk[i] = f | (b << 13) | (hl << 26);
	}
	__syncthreads();
	int big = 0;
	{
		u32 run = 0;
		#pragma unroll
		for (int i = 0; i < SG_ITEMS; i++) {
			u32 pm = run, tm = run;
			#pragma unroll
			for (int k = 0; k < SG_THREADS / 32; k++) { const u32 a = wf[i][k]; if (k < w) pm = max(pm, a); tm = max(tm, a); }
			pk[i] = (pk[i] & ~0x1fffu) | max(pk[i] & 0x1fffu, pm);
			run = tm;
		}
		run = 0x1fffu;
		#pragma unroll
		for (int i = SG_ITEMS - 1; i >= 0; i--) {
			u32 pm = run, tm = run;
			#pragma unroll
			for (int k = SG_THREADS / 32 - 1; k >= 0; k--) { const u32 a = wb[i][k]; if (k > w) pm = min(pm, a); tm = min(tm, a); }
			const u32 ge = min((pk[i] >> 13) & 0x1fffu, pm);
			pk[i] = (pk[i] & ~(0x1fffu << 13)) | (ge << 13);
			if ((pk[i] >> 28) & 1u) big |= (ge - (pk[i] & 0x1fffu) > SG_PAIR_MAX);
			run = tm;
		}
	}
	if (__syncthreads_or(big)) {                        // a long group: the radix kernel takes this tile
		if (t == 0) queue[atomicAdd(counters + 3, 1u)] = blockIdx.x;
		return;
	}
	// ---- warp-level rank refinement: new slot = group start + #smaller + #equal-and-earlier (stable); an
	// element opens a new sub-group iff no equal key precedes it. Groups here hold a few to a few dozen suffixes.
	#pragma unroll 2
	for (int i = 0; i < SG_ITEMS; i++) {
		const u32 e = i * SG_THREADS + t;
		if ((pk[i] >> 28) & 1u) {
			const u32 mine = skey[e];
			const u32 gs = pk[i] & 0x1fffu, ge = (pk[i] >> 13) & 0x1fffu;
			u32 cnt = 0, eqb = 0;
			for (u32 k = gs; k < e; k++) { const u32 o = skey[k]; cnt += (o <= mine); eqb += (o == mine); }
			for (u32 k = e + 1; k < ge; k++) cnt += (skey[k] < mine);
			pk[i] = (gs + cnt) | (eqb == 0 ? 0x10000u : 0u) | (1u << 28);
		}
	}
	__syncthreads();                                    // every key2 has been read: skey becomes the output staging
	#pragma unroll
	for (int i = 0; i < SG_ITEMS; i++) {
		const u32 e = i * SG_THREADS + t;
		if ((pk[i] >> 28) & 1u) { const u32 pos = pk[i] & 0xffffu; skey[pos] = sval[e]; sflag[pos] = (u8)((pk[i] >> 16) & 1u); }
	}
	__syncthreads();
	#pragma unroll
	for (int i = 0; i < SG_ITEMS; i++) {
		const u32 e = i * SG_THREADS + t;
		if (e < len) { V[start + e] = skey[e]; F[start + e] = sflag[e]; }
	}
}

// Radix route for the queued tiles: LSD radix sort of (local group id, key2) that never leaves shared memory.
template <bool RS>
__global__ void __launch_bounds__(SG_THREADS) k_seg_sort_radix(const u64* __restrict__ K, u32* __restrict__ V, u32 A,
                                                               const u32* __restrict__ ISA, u32 h, u32 n, int rank_bits,
                                                               u8* __restrict__ F, const u32* __restrict__ queue, int* __restrict__ err, RunSkip rs)
{
	extern __shared__ __align__(16) u8 sg_smem[];
	u64* skey = reinterpret_cast<u64*>(sg_smem);
	u32* sval = reinterpret_cast<u32*>(sg_smem + (size_t)SG_CAP * 8);
	__shared__ u32 wcnt[SG_THREADS / 32][256];
	__shared__ u32 bin_start[256];
	__shared__ u32 ws[32];
	__shared__ u32 sh[3];
	const int t = threadIdx.x, w = t >> 5, lane = t & 31;
	const u32 lt = lanemask_lt();
	const SegTile tile = seg_range(queue[blockIdx.x], K, A, rank_bits, sh, nullptr);
	const u32 len = tile.len, start = tile.start;
	if (len == 0) return;
	{
		const u64 g0 = K[start] >> rank_bits;
		#pragma unroll 4
		for (int i = 0; i < SG_ITEMS; i++) {
			const u32 e = i * SG_THREADS + t;
			if (e < len) {
				const u32 j = start + e;
				const u32 v = V[j];
				const u32 k2 = key2_of<RS>(v, h, n, ISA, rs, err);
				sval[e] = v;
				skey[e] = (((K[j] >> rank_bits) - g0) << 32) | (u64)k2;
			}
		}
		__syncthreads();
	}
	const u32 gmax = (u32)(skey[len - 1] >> 32);
	const int bits = rank_bits + bit_length((u64)gmax);
	u64 key[SG_ITEMS];
	u32 val[SG_ITEMS];
	#pragma unroll
	for (int i = 0; i < SG_ITEMS; i++) {
		const u32 e = w * (32 * SG_ITEMS) + i * 32 + lane;
		key[i] = ~0ull; val[i] = 0;
		if (e < len) { const u64 c = skey[e]; key[i] = ((c >> 32) << rank_bits) | (c & 0xffffffffull); val[i] = sval[e]; }
	}
	__syncthreads();

	for (int shift = 0; shift < bits; shift += 8) {
		for (int i = t; i < (SG_THREADS / 32) * 256; i += SG_THREADS) (&wcnt[0][0])[i] = 0;
		__syncthreads();
		u32 rank[SG_ITEMS];
		u32* mycnt = wcnt[w];
		#pragma unroll
		for (int i = 0; i < SG_ITEMS; i++) {
			const u32 d = rs_digit(key[i], shift);
			const u32 peers = __match_any_sync(0xffffffffu, d);
			const u32 below = __popc(peers & lt);
			u32 before = 0;
			if (below == 0) { before = mycnt[d]; mycnt[d] = before + __popc(peers); }
			before = __shfl_sync(0xffffffffu, before, __ffs(peers) - 1);
			rank[i] = before + below;
			__syncwarp();
		}
		__syncthreads();
		u32 run = 0;
		#pragma unroll
		for (int k = 0; k < SG_THREADS / 32; k++) { const u32 v = wcnt[k][t]; wcnt[k][t] = run; run += v; }
		u32 total;
		const u32 inc = block_incl_sum(run, ws, &total);
		bin_start[t] = inc - run;
		__syncthreads();
		#pragma unroll
		for (int i = 0; i < SG_ITEMS; i++) {
			const u32 d = rs_digit(key[i], shift);
			const u32 pos = bin_start[d] + mycnt[d] + rank[i];
			skey[pos] = key[i];
			sval[pos] = val[i];
		}
		__syncthreads();
		#pragma unroll
		for (int i = 0; i < SG_ITEMS; i++) {
			const u32 e = w * (32 * SG_ITEMS) + i * 32 + lane;
			key[i] = skey[e];
			val[i] = sval[e];
		}
		__syncthreads();
	}
	// after the loop the registers hold the sorted sequence in slot order and skey still holds the same data
	#pragma unroll
	for (int i = 0; i < SG_ITEMS; i++) {
		const u32 e = w * (32 * SG_ITEMS) + i * 32 + lane;
		if (e < len) {
			V[start + e] = val[i];
			F[start + e] = (e == 0 || skey[e - 1] != key[i]) ? 1 : 0;
		}
	}
}

// ---- 5b. doubling round, large groups -------------------------------------------------------------------
// A group longer than a window (common prefixes of real text, periodic data) cannot be sorted inside one block.
// The windows report where such groups start; k_large_collect finds where they end (the next group head of any
// later window) and lays them out back to back; their suffixes are extracted with key (large-group id, ISA[s+h]),
// sorted by the global radix sort, and written back in place with their head flags. Everything else in the
// active set stays on the shared-memory route.
__global__ void __launch_bounds__(1024) k_large_collect(u32* __restrict__ win_first, const u32* __restrict__ win_large, u32 nwin, u32 A,
                                                        u32* __restrict__ lg_head, u32* __restrict__ lg_off, u32* __restrict__ counters /*[4] groups [5] elements*/)
{
	__shared__ u32 ws[32];
	__shared__ u32 cmin[1024];
	const u32 t = threadIdx.x;
	const u32 per = (nwin + 1023) / 1024;
	const u32 lo = min(nwin, t * per), hi = min(nwin, lo + per);
	// next group head after each window: exclusive suffix minimum of win_first (in place)
	u32 m = 0xffffffffu;
	for (u32 c = lo; c < hi; c++) m = min(m, win_first[c]);
	cmin[t] = m;
	__syncthreads();
	u32 run = 0xffffffffu;
	for (u32 k = t + 1; k < 1024; k++) run = min(run, cmin[k]);      // 1024 x 1024 shared reads: negligible next to the sort
	u32 cnt = 0, sum = 0;
	for (u32 c = hi; c-- > lo;) {
		const u32 f = win_first[c];
		const u32 nf = min(run, A);
		win_first[c] = nf;
		run = min(run, f);
		const u32 h = win_large[c];
		if (h != 0xffffffffu) { cnt++; sum += nf - h; }
	}
	u32 tot_c, tot_s;
	const u32 ic = block_incl_sum(cnt, ws, &tot_c);
	const u32 is = block_incl_sum(sum, ws, &tot_s);
	u32 g = ic - cnt, off = is - sum;
	for (u32 c = lo; c < hi; c++) {
		const u32 h = win_large[c];
		if (h != 0xffffffffu) { lg_head[g] = h; lg_off[g] = off; off += win_first[c] - h; g++; }
	}
	if (t == 1023) { lg_off[tot_c] = tot_s; counters[4] = tot_c; counters[5] = tot_s; }
}

__device__ __forceinline__ u32 large_group_of(const u32* __restrict__ lg_off, u32 ng, u32 x)
{
	u32 lo = 0, hi = ng;                                  // last g with lg_off[g] <= x
	while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (lg_off[mid] <= x) lo = mid; else hi = mid; }
	return lo;
}

template <bool RS>
__global__ void __launch_bounds__(256) k_large_extract(const u32* __restrict__ V, const u32* __restrict__ ISA, u32 h, u32 n, int rank_bits,
                                                       const u32* __restrict__ lg_head, const u32* __restrict__ lg_off, u32 ng, u32 total,
                                                       u64* __restrict__ LK, u32* __restrict__ LV, int* __restrict__ err, RunSkip rs)
{
	const u32 x = blockIdx.x * 256 + threadIdx.x;
	if (x >= total) return;
	const u32 g = large_group_of(lg_off, ng, x);
	const u32 v = V[lg_head[g] + (x - lg_off[g])];
	const u32 k2 = key2_of<RS>(v, h, n, ISA, rs, err);
	LK[x] = ((u64)g << rank_bits) | (u64)k2;
	LV[x] = v;
}

__global__ void __launch_bounds__(256) k_large_writeback(const u64* __restrict__ LK, const u32* __restrict__ LV, int rank_bits,
                                                         const u32* __restrict__ lg_head, const u32* __restrict__ lg_off, u32 total,
                                                         u32* __restrict__ V, u8* __restrict__ F)
{
	const u32 x = blockIdx.x * 256 + threadIdx.x;
	if (x >= total) return;
	const u64 k = LK[x];
	const u32 g = (u32)(k >> rank_bits);
	const u32 o = lg_off[g];
	const u32 slot = lg_head[g] + (x - o);
	V[slot] = LV[x];
	F[slot] = (x == o || LK[x - 1] != k) ? 1 : 0;
}

// ---- 5c. doubling round, composite-key route for the whole active set (A/B reference: JP_BWT_FWD_GLOBAL=1) ---------------------------------------------------
template <bool RS>
__global__ void __launch_bounds__(256) k_fwd_gather(u64* __restrict__ K, const u32* __restrict__ V, u32 A,
                                                    const u32* __restrict__ ISA, u32 h, u32 n, int* __restrict__ err, RunSkip rs)
{
	const u32 j = blockIdx.x * 256 + threadIdx.x;
	if (j >= A) return;
	K[j] |= (u64)key2_of<RS>(V[j], h, n, ISA, rs, err);
}

// ---- 6. emission ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fwd_emit(const u8* __restrict__ T, const u32* __restrict__ SA,
                                                  const u32* __restrict__ ISA, i32 n, u8* __restrict__ out)
{
	const i64 o0 = ((i64)blockIdx.x * 256 + threadIdx.x) * 4;
	if (o0 >= n) return;
	const i64 idx0 = (i64)ISA[0] - 1;                 // SA position of suffix 0 (bwt.cpp:51)
	u32 acc = 0; int cnt = 0;
	#pragma unroll
	for (int b = 0; b < 4; b++) {
		const i64 o = o0 + b;
		if (o < n) {
			u32 c;
			if (o == 0) c = T[n - 1];                 // bwt.cpp:50
			else {
				const i64 i = (o <= idx0) ? o - 1 : o; // bwt.cpp:53-56
				c = T[SA[i] - 1];
			}
			acc |= c << (8 * b); cnt++;
		}
	}
	if (cnt == 4) *reinterpret_cast<u32*>(out + o0) = acc;
	else for (int b = 0; b < cnt; b++) out[o0 + b] = (u8)(acc >> (8 * b));
}

__global__ void k_fwd_trailer(const u8* __restrict__ T, const u32* __restrict__ ISA, i32 n, i32 len, u8* __restrict__ out)
{
	const int t = threadIdx.x;
	const i32 step = n / JP_BWT_UNITS;                 // bwt.cpp:44
	if (t < JP_BWT_UNITS) {
		const u32 v = ISA[(i64)t * step];              // = Indicies[t] + 1 (bwt.cpp:46-48,57-58)
		u8* p = out + len + 4 * t;                     // unaligned, native-endian (bwt.cpp:60-61)
		p[0] = (u8)v; p[1] = (u8)(v >> 8); p[2] = (u8)(v >> 16); p[3] = (u8)(v >> 24);
	}
	for (int i = t; i < len - n; i += blockDim.x) out[n + i] = T[n + i];   // bwt.cpp:32-33
}

// ---- host driver ---------------------------------------------------------------------------------------
struct FwdBuffers {
	RadixBuffers rb;
	u32* P[2];
	u32* ISA; u32* SA; u32* R;
	u32* RL; u32* run_bits; u32* run_first; u32* run_next; u32* run_count;   // run skip (null when off)
	int isa_region_log2;
	u8* F;
	u32* queue;
	u32* win_first; u32* win_large; u32* lg_head; u32* lg_off;
	GAgg* agg;
	FwdMeta* meta;
	u32* counters;
	int* err;
};

static bool runskip_enabled()
{
	const char* e = getenv("JP_BWT_FWD_RUNSKIP");
	return e != nullptr && atoi(e) != 0;
}

static int fwd_alloc(Ctx& c, i32 n, FwdBuffers& b)
{
	const size_t N = (size_t)n;
	const size_t rtiles = radix_tiles(N), gtiles = (N + GS_TILE - 1) / GS_TILE;
	size_t total = 2 * Arena::align(N * 8) + 4 * Arena::align(N * 4) + Arena::align((N + 1) * 4) + 2 * Arena::align(N * 4) +
	               Arena::align((rtiles + 4) * 256 * 4) + Arena::align(256 * 4) + Arena::align((8 * 256 + 64) * 4) + Arena::align(N + 64) + Arena::align(gtiles * sizeof(GAgg)) +
	               Arena::align(sizeof(FwdMeta)) + Arena::align(64) + Arena::align(64) + Arena::align(N + 16) + 5 * Arena::align((N / SG_WIN + 16) * 4);
	const bool runskip = runskip_enabled();
	const size_t run_tiles = (N + RUN_TILE - 1) / RUN_TILE;
	if (runskip) total += Arena::align(N * 4) + Arena::align((N / 32 + 2) * 4) + 2 * Arena::align((run_tiles + 1) * 4) + Arena::align(64);
	JP_TRY(arena_reserve(c, total));
	b.rb.k[0] = arena_take<u64>(c, N); b.rb.k[1] = arena_take<u64>(c, N);
	b.rb.v[0] = arena_take<u32>(c, N); b.rb.v[1] = arena_take<u32>(c, N);
	b.P[0] = arena_take<u32>(c, N); b.P[1] = arena_take<u32>(c, N);
	b.ISA = arena_take<u32>(c, N + 1); b.SA = arena_take<u32>(c, N); b.R = arena_take<u32>(c, N);
	b.isa_region_log2 = 24;                            // 2^24 ranks = 64 MiB of ISA per pass (measured best of 2^21..2^25)
	if (const char* e = getenv("JP_BWT_ISA_REGION_LOG2")) b.isa_region_log2 = atoi(e);
	b.rb.tile_hist = arena_take<u32>(c, (rtiles + 4) * 256);
	b.rb.totals = arena_take<u32>(c, 256);
	b.rb.os_state = arena_take<u32>(c, 8 * 256 + 64);
	b.rb.dnext = getenv("JP_BWT_RADIX_NO_DIGIT_BYTES") ? nullptr : arena_take<u8>(c, N + 64);
	// Measured on B200 (64 M pairs, 8 passes): three-kernel passes 5.07 ms, one-sweep (all digits counted in one read of
	// the keys + 8 look-back passes) 6.18 ms -- with ~440 tiles in flight the per-digit look-back chain costs more than
	// the key re-read it saves, so the classic pass is the default.
	b.rb.classic = getenv("JP_BWT_RADIX_ONESWEEP") == nullptr;
	b.F = arena_take<u8>(c, N + 16);
	b.queue = arena_take<u32>(c, N / SG_WIN + 16);
	b.win_first = arena_take<u32>(c, N / SG_WIN + 16); b.win_large = arena_take<u32>(c, N / SG_WIN + 16);
	b.lg_head = arena_take<u32>(c, N / SG_WIN + 16); b.lg_off = arena_take<u32>(c, N / SG_WIN + 16);
	b.agg = arena_take<GAgg>(c, gtiles);
	b.meta = arena_take<FwdMeta>(c, 1);
	b.counters = arena_take<u32>(c, 16);
	b.err = arena_take<int>(c, 16);
	b.rb.err = b.err;
	b.RL = b.run_bits = b.run_first = b.run_next = b.run_count = nullptr;
	if (runskip) {
		b.RL = arena_take<u32>(c, N); b.run_bits = arena_take<u32>(c, N / 32 + 2);
		b.run_first = arena_take<u32>(c, run_tiles + 1); b.run_next = arena_take<u32>(c, run_tiles + 1);
		b.run_count = arena_take<u32>(c, 16);
	}
	return JP_OK;
}

// One grouping step over the sorted pairs in rb.k/v[cur]; survivors land in rb.k/v[cur^1] and P[pc^1].
static int group_step(Ctx& c, FwdBuffers& b, int cur, int pc, bool identity_pos, bool use_flags, u32 A, int rank_bits, cudaStream_t s)
{
	const int tiles = (int)((A + GS_TILE - 1) / GS_TILE);
	const u32* P = identity_pos ? nullptr : b.P[pc];
	const u8* F = use_flags ? b.F : nullptr;
	k_grp_reduce<<<tiles, GS_THREADS, 0, s>>>(b.rb.k[cur], F, P, A, b.agg); JP_LAUNCH(c);
	k_grp_scan_tiles<<<1, 1024, 0, s>>>(b.agg, tiles, b.counters); JP_LAUNCH(c);
	// small active sets (and the A/B switch region_log2 <= 0) scatter straight from the apply kernel
	const bool staged = b.isa_region_log2 > 0 && A > (1u << 20);
	k_grp_apply<<<tiles, GS_THREADS, 0, s>>>(b.rb.k[cur], F, b.rb.v[cur], P, A, b.agg, rank_bits, b.ISA, staged ? b.R : nullptr, b.SA,
	                                          b.rb.k[cur ^ 1], b.rb.v[cur ^ 1], b.P[pc ^ 1]); JP_LAUNCH(c);
	if (staged) {
		const u32 n_entries = c.cur_n + 1, regions = (n_entries + (1u << b.isa_region_log2) - 1) >> b.isa_region_log2;
		const u32 stiles = (A + 2047) / 2048;
		if (regions <= 4 || regions > 256) {
			// few regions: stream the slots once per region and keep the ranks that fall in it
			k_isa_scatter<<<regions * stiles, 256, 0, s>>>(b.rb.v[cur], b.R, A, b.ISA, b.isa_region_log2, stiles); JP_LAUNCH(c);
		} else {
			// many regions (blocks over 64 MiB): one radix partition pass buckets the (suffix, rank) pairs by region -- the
			// sorted keys of this step are dead, their buffer holds the bucketed pairs -- then a single ordered sweep
			u32* pv = reinterpret_cast<u32*>(b.rb.k[cur]);
			u32* pr = pv + A;
			if (radix_partition_u32(b.rb.v[cur], b.R, pv, pr, A, b.isa_region_log2, b.rb.tile_hist, b.rb.totals, s, &c.launches) != 0) { set_error_detail("radix partition setup failed"); return JP_ERR_CUDA; }
			k_isa_scatter<<<stiles, 256, 0, s>>>(pv, pr, A, b.ISA, 32, stiles); JP_LAUNCH(c);
		}
	}
	JP_KCHECK();
	JP_CUDA(cudaMemcpyAsync(c.h_small + 8, b.counters, 2 * sizeof(u32), cudaMemcpyDeviceToHost, s));
	JP_CUDA(cudaMemcpyAsync(c.h_small, b.err, sizeof(int), cudaMemcpyDeviceToHost, s));
	JP_CUDA(cudaStreamSynchronize(s));
	return JP_OK;
}

// Builds SA and ISA (ranks 1..n; ISA[n] = 0) of T[0..n) in b. Events ev[1..4] mark the phase boundaries.
static int suffix_sort(Ctx& c, const u8* d_T, i32 n, FwdBuffers& b, cudaStream_t s, jp_bwt_stats* st)
{
	c.cur_n = (u32)n;
	JP_CUDA(cudaMemsetAsync(b.meta, 0, sizeof(FwdMeta), s));
	JP_CUDA(cudaMemsetAsync(b.err, 0, 64, s));
	JP_CUDA(cudaMemsetAsync(b.ISA + n, 0, sizeof(u32), s));             // the empty suffix ranks below everything
	const i64 hwant = ((i64)n + 4095) / 4096, hcap = (i64)c.sm_count * 8;
	const int hblocks = (int)(hwant < hcap ? hwant : hcap);
	k_fwd_symhist<<<hblocks, 256, 0, s>>>(d_T, n, b.meta); JP_LAUNCH(c);
	k_fwd_codes<<<1, 256, 0, s>>>(b.meta); JP_LAUNCH(c);
	JP_KCHECK();
	JP_CUDA(cudaMemcpyAsync(c.h_small + 16, &b.meta->sigma, 4 * sizeof(i32), cudaMemcpyDeviceToHost, s)); // sigma, bits, depth, key_bits
	JP_CUDA(cudaStreamSynchronize(s));
	const int bits = c.h_small[17], depth = c.h_small[18], key_bits0 = c.h_small[19];
	if (bits < 1 || bits > 9 || depth < 7 || depth > 63 || key_bits0 < 1 || key_bits0 > 63) { set_error_detail("symbol remap gave bits=%d depth=%d key bits=%d", bits, depth, key_bits0); return JP_ERR_INTERNAL; }
	st->symbol_bits = bits; st->initial_depth = depth;

	k_fwd_keys<<<(n + KEY_TILE - 1) / KEY_TILE, 256, 0, s>>>(d_T, n, b.meta, b.rb.k[0], b.rb.v[0], b.rb.tile_hist,
	                                                         rs_stride((u32)radix_tiles((size_t)n))); JP_LAUNCH(c);
	JP_KCHECK();
	const bool runskip = b.RL != nullptr;                                // JP_BWT_FWD_RUNSKIP=1 (see "run skip" above)
	if (runskip) {
		const u32 run_tiles = (u32)(((size_t)n + RUN_TILE - 1) / RUN_TILE);
		JP_CUDA(cudaMemsetAsync(b.run_count, 0, 64, s));
		k_run_first<<<run_tiles, 256, 0, s>>>(d_T, (u32)n, b.run_first); JP_LAUNCH(c);
		k_run_scan<<<1, 1024, 0, s>>>(b.run_first, run_tiles, b.run_next); JP_LAUNCH(c);
		k_run_fill<<<run_tiles, 256, 0, s>>>(d_T, (u32)n, b.run_next, (u32)depth, b.RL, b.run_bits, b.run_count); JP_LAUNCH(c);
		JP_KCHECK();
		JP_CUDA(cudaMemcpyAsync(c.h_small + 14, b.run_count, sizeof(u32), cudaMemcpyDeviceToHost, s));   // read after the first group step's sync
	}
	JP_CUDA(cudaEventRecord(c.ev[1], s));
	const bool force_global = getenv("JP_BWT_FWD_GLOBAL") != nullptr;    // A/B switch: composite-key route for every round
	if (cudaFuncSetAttribute(k_seg_sort<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG_SMEM_LIGHT) != cudaSuccess ||
	    cudaFuncSetAttribute(k_seg_sort_radix<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG_SMEM) != cudaSuccess ||
	    cudaFuncSetAttribute(k_seg_sort<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG_SMEM_LIGHT) != cudaSuccess ||
	    cudaFuncSetAttribute(k_seg_sort_radix<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG_SMEM) != cudaSuccess) { set_error_detail("k_seg_sort smem attribute"); return JP_ERR_CUDA; }
	int cur = radix_sort_pairs(b.rb, 0, (u32)n, 0, key_bits0, s, &c.launches, /*first_hist_ready=*/true);
	if (cur < 0) { set_error_detail("radix sort setup failed"); return JP_ERR_CUDA; }
	JP_KCHECK();
	JP_CUDA(cudaEventRecord(c.ev[2], s));

	const int rank_bits = bit_length(runskip ? 2 * (u64)n + 1 : (u64)n);   // key2 is a rank <= n, or a run key <= 2n
	int pc = 0;
	JP_TRY(group_step(c, b, cur, pc, true, false, (u32)n, rank_bits, s));
	const bool use_rs = runskip && c.h_small[14] != 0;                   // no run suffix in this block: the plain kernels
	RunSkip rs; rs.rl = b.RL; rs.bits = b.run_bits; rs.T = d_T; rs.first = 1;
	JP_CUDA(cudaEventRecord(c.ev[3], s));
	int act = cur ^ 1; pc ^= 1;
	u32 A = (u32)c.h_small[8], G = (u32)c.h_small[9];
	u64 sectors = 2ull * (u64)n;
	i64 h = depth;
	int rounds = 0;
	double large_frac_prev = 0.0;
	const bool trace_rounds = getenv("JP_BWT_TRACE_ROUNDS") != nullptr;      // one stderr line per doubling round (host wall time)
	double t_round = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
	while (A > 0) {
		if (c.h_small[0] != 0) return map_dev_err(c.h_small[0]);
		rs.first = rounds == 0 ? 1u : 0u;                                // run skip: pass A, then pass B at the same h
		if (rounds >= JP_BWT_MAX_ROUNDS || h > (i64)n) { set_error_detail("doubling stuck: round %d h=%lld active=%u", rounds, (long long)h, A); return JP_ERR_INTERNAL; }
		st->active_fraction[rounds] = (float)((double)A / (double)n);
		sectors += 2ull * A;
		double large_frac_now = 0.0;
		// a round that had most of the block in over-long groups (periodic data, one-symbol runs) is followed by more
		// of the same: skip the shared-memory kernels and sort the whole active set by composite key
		if (!force_global && large_frac_prev <= 0.5) {
			// short groups: fused gather + warp-level rank refinement in shared memory; longer ones are queued for the
			// shared-memory radix kernel; groups longer than a window are reported for the large-group route
			const u32 nwin = (A + SG_WIN - 1) / SG_WIN;
			JP_CUDA(cudaMemsetAsync(b.counters + 2, 0, 4 * sizeof(u32), s));
			if (use_rs) k_seg_sort<true><<<nwin, SG_THREADS, SG_SMEM_LIGHT, s>>>(b.rb.k[act], b.rb.v[act], A, b.ISA, (u32)h, (u32)n, rank_bits, b.F,
			                                                   b.counters, b.queue, b.win_first, b.win_large, b.err, rs);
			else k_seg_sort<false><<<nwin, SG_THREADS, SG_SMEM_LIGHT, s>>>(b.rb.k[act], b.rb.v[act], A, b.ISA, (u32)h, (u32)n, rank_bits, b.F,
			                                                   b.counters, b.queue, b.win_first, b.win_large, b.err, rs);
			JP_LAUNCH(c);
			JP_KCHECK();
			JP_CUDA(cudaMemcpyAsync(c.h_small + 10, b.counters + 2, 2 * sizeof(u32), cudaMemcpyDeviceToHost, s));
			JP_CUDA(cudaStreamSynchronize(s));
			const bool large = c.h_small[10] != 0;
			const u32 queued = (u32)c.h_small[11];
			if (queued) {
				if (use_rs) k_seg_sort_radix<true><<<queued, SG_THREADS, SG_SMEM, s>>>(b.rb.k[act], b.rb.v[act], A, b.ISA, (u32)h, (u32)n,
				                                                    rank_bits, b.F, b.queue, b.err, rs);
				else k_seg_sort_radix<false><<<queued, SG_THREADS, SG_SMEM, s>>>(b.rb.k[act], b.rb.v[act], A, b.ISA, (u32)h, (u32)n,
				                                                    rank_bits, b.F, b.queue, b.err, rs);
				JP_LAUNCH(c);
				JP_KCHECK();
				st->ms_phase[6] += (float)queued;       // tiles that needed the shared-memory radix route
			}
			if (large) {
				k_large_collect<<<1, 1024, 0, s>>>(b.win_first, b.win_large, nwin, A, b.lg_head, b.lg_off, b.counters); JP_LAUNCH(c);
				JP_KCHECK();
				JP_CUDA(cudaMemcpyAsync(c.h_small + 12, b.counters + 4, 2 * sizeof(u32), cudaMemcpyDeviceToHost, s));
				JP_CUDA(cudaStreamSynchronize(s));
				const u32 ng = (u32)c.h_small[12], total = (u32)c.h_small[13];
				st->ms_phase[5] += (float)((double)total / (double)n);   // fraction of the block that went through the large-group route, summed over rounds
				large_frac_now = (double)total / (double)A;
				if (ng == 0 || total == 0 || total > A) { set_error_detail("large-group list inconsistent: %u groups, %u suffixes, %u active", ng, total, A); return JP_ERR_INTERNAL; }
				// the gidx keys of the active set are dead once the shared-memory kernels are done: sort in the spare buffers
				RadixBuffers lb = b.rb;
				lb.k[0] = b.rb.k[act ^ 1]; lb.k[1] = b.rb.k[act];
				lb.v[0] = b.rb.v[act ^ 1]; lb.v[1] = b.R;
				if (use_rs) k_large_extract<true><<<(total + 255) / 256, 256, 0, s>>>(b.rb.v[act], b.ISA, (u32)h, (u32)n, rank_bits, b.lg_head, b.lg_off, ng, total,
				                                                    lb.k[0], lb.v[0], b.err, rs);
				else k_large_extract<false><<<(total + 255) / 256, 256, 0, s>>>(b.rb.v[act], b.ISA, (u32)h, (u32)n, rank_bits, b.lg_head, b.lg_off, ng, total,
				                                                    lb.k[0], lb.v[0], b.err, rs);
				JP_LAUNCH(c);
				const int key_bits = rank_bits + bit_length((u64)(ng - 1));
				const int lc = radix_sort_pairs(lb, 0, total, 0, key_bits, s, &c.launches);
				if (lc < 0) { set_error_detail("radix sort setup failed"); return JP_ERR_CUDA; }
				k_large_writeback<<<(total + 255) / 256, 256, 0, s>>>(lb.k[lc], lb.v[lc], rank_bits, b.lg_head, b.lg_off, total, b.rb.v[act], b.F); JP_LAUNCH(c);
				JP_KCHECK();
			}
			cur = act;
			JP_TRY(group_step(c, b, cur, pc, false, true, A, rank_bits, s));
		} else {
			st->ms_phase[5] += (float)((double)A / (double)n);
			large_frac_now = G * (u64)SG_WIN >= A ? 0.0 : 1.0;       // back to the segmented route once groups average under a window
			if (use_rs) k_fwd_gather<true><<<(A + 255) / 256, 256, 0, s>>>(b.rb.k[act], b.rb.v[act], A, b.ISA, (u32)h, (u32)n, b.err, rs);
			else k_fwd_gather<false><<<(A + 255) / 256, 256, 0, s>>>(b.rb.k[act], b.rb.v[act], A, b.ISA, (u32)h, (u32)n, b.err, rs);
			JP_LAUNCH(c);
			const int key_bits = rank_bits + bit_length((u64)(G > 0 ? G - 1 : 0));
			cur = radix_sort_pairs(b.rb, act, A, 0, key_bits, s, &c.launches);
			if (cur < 0) { set_error_detail("radix sort setup failed"); return JP_ERR_CUDA; }
			JP_TRY(group_step(c, b, cur, pc, false, false, A, rank_bits, s));
		}
		if (trace_rounds) {
			const double t1 = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
			fprintf(stderr, "[jp_bwt round] %d h=%lld active=%u groups=%u -> active=%u groups=%u large_frac=%.4f %.3f ms\n", rounds, (long long)h, A, G,
			        (u32)c.h_small[8], (u32)c.h_small[9], large_frac_now, t1 - t_round);
			t_round = t1;
		}
		act = cur ^ 1; pc ^= 1;
		A = (u32)c.h_small[8]; G = (u32)c.h_small[9];
		large_frac_prev = large_frac_now;
		if (!(use_rs && rounds == 0)) h *= 2;
		rounds++;
	}
	if (c.h_small[0] != 0) return map_dev_err(c.h_small[0]);
	JP_CUDA(cudaEventRecord(c.ev[4], s));
	st->rounds = rounds;
	st->random_sectors = sectors;
	return JP_OK;
}

int forward_device(Ctx& c, const u8* d_in, i32 len, u8* d_out, cudaStream_t s, jp_bwt_stats* st)
{
	const i32 nlen = len - len % JP_BWT_UNITS;                          // bwt.cpp:29-30
	st->direction = 0; st->len = len; st->nlen = nlen; st->device = c.device;
	JP_CUDA(cudaEventRecord(c.ev[0], s));
	if (nlen == 0) {                                                    // bwt.cpp:35: tail only, trailer untouched
		if (len > 0) JP_CUDA(cudaMemcpyAsync(d_out, d_in, (size_t)len, cudaMemcpyDeviceToDevice, s));
		JP_CUDA(cudaEventRecord(c.ev[1], s));
		JP_CUDA(cudaStreamSynchronize(s));
		JP_CUDA(cudaEventElapsedTime(&st->ms_total, c.ev[0], c.ev[1]));
		return JP_OK;
	}
	FwdBuffers b;
	JP_TRY(fwd_alloc(c, nlen, b));
	JP_TRY(suffix_sort(c, d_in, nlen, b, s, st));
	k_fwd_emit<<<(int)(((i64)nlen + 1023) / 1024), 256, 0, s>>>(d_in, b.SA, b.ISA, nlen, d_out); JP_LAUNCH(c);
	k_fwd_trailer<<<1, 128, 0, s>>>(d_in, b.ISA, nlen, len, d_out); JP_LAUNCH(c);
	JP_KCHECK();
	JP_CUDA(cudaEventRecord(c.ev[5], s));
	JP_CUDA(cudaStreamSynchronize(s));
	for (int i = 0; i < 5; i++) JP_CUDA(cudaEventElapsedTime(&st->ms_phase[i], c.ev[i], c.ev[i + 1]));
	JP_CUDA(cudaEventElapsedTime(&st->ms_total, c.ev[0], c.ev[5]));
	st->device_bytes = c.arena.high;
	return JP_OK;
}

int debug_suffix_array(Ctx& c, const u8* h_in, i32 n, i32* h_sa)
{
	if (n == 0) return JP_OK;
	cudaStream_t s = c.own_stream;
	c.arena.reset();
	FwdBuffers b;
	const size_t N = (size_t)n;
	JP_TRY(arena_reserve(c, N * 56 + (8u << 20)));
	u8* d_T = arena_take<u8>(c, N + 16);
	JP_TRY(fwd_alloc(c, n, b));
	JP_CUDA(cudaMemcpyAsync(d_T, h_in, N, cudaMemcpyHostToDevice, s));
	JP_CUDA(cudaEventRecord(c.ev[0], s));
	jp_bwt_stats st = {};
	JP_TRY(suffix_sort(c, d_T, n, b, s, &st));
	JP_CUDA(cudaMemcpyAsync(h_sa, b.SA, N * 4, cudaMemcpyDeviceToHost, s));
	JP_CUDA(cudaStreamSynchronize(s));
	return JP_OK;
}

} // namespace jp
