// bwt_forward.cu -- forward BWT for sm_100a: GPU suffix sort + BWT emission + the 120 sampled indices.
//
// Replaces BlockSort::Bwt::ForwardBwt (reference bwt.cpp:22-65) and the divsufsort() call under it
// (divsufsort.cpp:1721; contract divsufsort.hpp:37-45: plain suffix array, a proper prefix sorts first).
// Nothing of divsufsort's induced-copying design is kept: a serial induce pass has no place on 148 SMs.
//
//   1. symbol remap   present byte values -> dense codes 1..sigma (0 = end of string)
//   2. initial keys   key(i) = the first d codes of suffix i as a base-(sigma+1) number, zero padded: the padding IS
//                     the sentinel, so "shorter sorts first" needs no tie-break
//   3. radix bucket   LSD radix sort of (key, i)                                      [radix_sort.cuh]
//   4. ranks          the sorted suffix ids ARE the suffix array, in place; group heads = key changes; rank = 1 +
//                     position of the group's head; singletons are final, the rest form the ACTIVE SET: one 32-bit
//                     word per unsorted suffix = its SA position, bit 31 = first of its group
//   5. doubling       while any group is unsorted: every active suffix s takes key2 = ISA[s + h]; groups are refined in
//                     place in SA by key2 (shared-memory kernels for groups up to a tile, a global radix sort for longer
//                     ones), ranks are updated, sorted suffixes leave the active set; h doubles. Only the active set is
//                     touched (Larsson-Sadakane discarding); ISA[nlen] = 0 is the empty suffix.
//   6. emit           bwt[o] = T[SA[row]-1] with the row of suffix 0 skipped (bwt.cpp:50-56), the sampled
//                     indices ISA[k*step] (bwt.cpp:44-48,57-61), the raw tail (bwt.cpp:32-33).
//
// Device memory: six units of 4(N+2) bytes -- two (key, id) buffer pairs during the initial sort; SA, ISA, two
// active-set buffers, a staging unit and a spare during the rounds -- plus N/4 of radix histograms: 24.3 N. The
// one-byte arrays (next-digit bytes of the radix passes, head flags of the rounds) live in the caller's output block,
// which is dead until the emission.
#include "bwt_internal.cuh"
#include "radix_sort.cuh"
#include <algorithm>
#include <chrono>

namespace jp {

constexpr u32 AP_HEAD = 0x80000000u;     // active-set word: this slot is the first of its group
constexpr u32 AP_POS  = 0x3fffffffu;     // ... and its position in SA (blocks stay below 2^30, bwt_internal.cuh)

struct FwdMeta {
	u32 code[256];
	i32 sigma, bits, depth, key_bits;   // bits: per symbol (reported); depth: symbols per key; key_bits: bit length of the largest key
	u32 hist[256];
	u32 eq4;                            // aligned 4-byte words of one repeated byte: a cheap screen for single-symbol runs
};

// ---- 1. symbol histogram and dense codes -------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fwd_symhist(const u8* __restrict__ T, i32 n, FwdMeta* __restrict__ meta)
{
	__shared__ u32 h[8][256];
	__shared__ u32 s_eq4;
	const int t = threadIdx.x, w = t >> 5;
	for (int i = t; i < 8 * 256; i += 256) (&h[0][0])[i] = 0;
	if (t == 0) s_eq4 = 0;
	__syncthreads();
	const i64 stride = (i64)gridDim.x * 256 * 16;
	u32 eq4 = 0;
	for (i64 p = ((i64)blockIdx.x * 256 + t) * 16; p < n; p += stride) {
		if (p + 16 <= n) {
			const uint4 q = __ldg(reinterpret_cast<const uint4*>(T + p));
			const u32 wd[4] = {q.x, q.y, q.z, q.w};
			#pragma unroll
			for (int k = 0; k < 4; k++) eq4 += (wd[k] == (wd[k] & 255u) * 0x01010101u);
			#pragma unroll
			for (int k = 0; k < 16; k++) atomicAdd(&h[w][(wd[k >> 2] >> ((k & 3) * 8)) & 255], 1u);
		} else for (i64 q = p; q < n; q++) atomicAdd(&h[w][T[q]], 1u);
	}
	eq4 = warp_sum(eq4);
	if ((t & 31) == 0 && eq4) atomicAdd(&s_eq4, eq4);
	__syncthreads();
	u32 s = 0;
	#pragma unroll
	for (int k = 0; k < 8; k++) s += h[k][t];
	if (s) atomicAdd(&meta->hist[t], s);
	if (t == 0 && s_eq4) atomicAdd(&meta->eq4, s_eq4);
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

// ---- 4. group heads -> ranks, retire singletons, compact the rest -------------------------------------
// The suffix array is refined IN PLACE: a slot of the active set names a position of SA, positions of a group are
// consecutive, and a suffix that has become a singleton simply stays where it is. What the grouping step produces per
// slot is (a) the rank of its suffix = 1 + position of the group's head and (b), for slots of groups that are still
// unsorted, the next active set. Scan element: (position of the last group head so far, #survivors, #surviving heads).
constexpr int GS_THREADS = 256;
constexpr int GS_SUB     = 8;
constexpr int GS_TILE    = GS_THREADS * GS_SUB;
struct GAgg { i32 mh; u32 ns; u32 ng; u32 pad; };

// head flags of the initial order: a key change
__global__ void __launch_bounds__(256) k_fwd_flags(const u64* __restrict__ K, u32 n, u8* __restrict__ F)
{
	const u32 j0 = (blockIdx.x * 256 + threadIdx.x) * 4;
	if (j0 >= n) return;
	u64 prev = j0 ? K[j0 - 1] : ~K[0];
	u32 w = 0;
	#pragma unroll
	for (int b = 0; b < 4; b++) {
		const u32 j = j0 + b;
		if (j < n) { const u64 k = K[j]; w |= (k != prev ? 1u : 0u) << (8 * b); prev = k; }
	}
	if (j0 + 4 <= n) *reinterpret_cast<u32*>(F + j0) = w;
	else for (int b = 0; b < 4 && j0 + b < n; b++) F[j0 + b] = (u8)(w >> (8 * b));
}

__global__ void __launch_bounds__(GS_THREADS) k_grp_reduce(const u8* __restrict__ F, const u32* __restrict__ AP, u32 A, GAgg* __restrict__ agg)
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
			const bool head = F[j] != 0, nhead = (j + 1 == A) || (F[j + 1] != 0);
			if (head) mh = max(mh, (i32)(AP ? (AP[j] & AP_POS) : j));
			ns += !(head && nhead);
			ng += (head && !nhead);
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
// inside each warp, and stitched together with ONE barrier through a [sub-tile][warp] table.
//   AP == nullptr: the slots are the positions 0..A-1 themselves (the step after the initial sort).
//   R  != nullptr: ranks leave in slot order (and the suffix ids in VS, if given) for k_isa_scatter to place; R may be
//                  AP itself -- every thread reads its own slots before it writes them.
__global__ void __launch_bounds__(GS_THREADS) k_grp_apply(const u8* __restrict__ F, const u32* AP, const u32* __restrict__ SA, u32 A,
                                                          const GAgg* __restrict__ agg, u32* __restrict__ ISA, u32* R, u32* __restrict__ VS,
                                                          u32* __restrict__ APn)
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
			const bool head = F[j] != 0, nhead = (j + 1 == A) || (F[j + 1] != 0);
			fl[s] = 1u | (head ? 2u : 0u) | (!(head && nhead) ? 4u : 0u) | ((head && !nhead) ? 8u : 0u);
			p[s] = AP ? (AP[j] & AP_POS) : j;
			v[s] = SA[p[s]];
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
			const u32 j = base + s * GS_THREADS + t;
			if (R) { R[j] = (u32)fmh + 1u; if (VS) VS[j] = v[s]; }      // ranks leave in slot order; k_isa_scatter places them
			else ISA[v[s]] = (u32)fmh + 1u;
			if (fl[s] & 4u) APn[carry0.ns + (fpk & 0xffffu) - 1] = p[s] | ((fl[s] & 2u) ? AP_HEAD : 0u);
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

// ---- 5. doubling rounds ----------------------------------------------------------------------------------
// key2 of suffix v in the round with offset h. (The template parameter selects the periodic-run keys added below.)
struct PerSkip { const u32* rl; const u32* bits; const u8* T; u32 p; u32 first; };

template <bool PS>
__device__ __forceinline__ u32 key2_of(u32 v, u32 h, u32 n, const u32* __restrict__ ISA, const PerSkip& ps, int* __restrict__ err)
{
	if (PS) {
		if ((__ldg(&ps.bits[v >> 5]) >> (v & 31)) & 1u) {
			const u32 r = __ldg(&ps.rl[v]);                    // v + r <= n by construction, r >= p
			if (ps.first) {
				const bool smaller_follows = (v + r >= n) || ps.T[v + r] < ps.T[v + r - ps.p];
				return smaller_follows ? r : 2u * n + 1u - r;
			}
			if (r >= h) return __ldg(&ISA[v + r]);
		}
	}
	u32 p = v + h;
	if (p > n) { dev_fail(err, DE_FWD_RANGE); p = n; }
	return __ldg(&ISA[p]);
}

// ---- 5a. small groups: gather + segmented sort in shared memory -----------------------------------------
// Groups are contiguous in the active set, so refining them is a SEGMENTED sort. Block c owns the groups whose
// head lies in its window of SG_WIN slots; they end within two windows unless the last one is "large". The block
// gathers key2 for its (<= 4096) elements, orders every group by key2 in shared memory, writes the suffix ids back
// to their group's positions of SA and one head flag per slot.
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
__device__ __forceinline__ SegTile seg_range(u32 win, const u32* __restrict__ AP, u32 A,
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
		if (j < L && (j == 0 || (AP[j] & AP_HEAD))) { atomicMin(&sh[0], j); atomicMax(&sh[1], j); }
	}
	__syncthreads();
	const u32 start = sh[0];
	if (t == 0 && win_first) { win_first[win] = start; win_large[win] = 0xffffffffu; }
	if (start == 0xffffffffu) return r;               // the window lies inside a group owned by an earlier block
	u32 end;
	if (L == A) end = A;
	else {
		const u32 lim = min(w0 + 2u * SG_WIN, A);
		#pragma unroll
		for (int i = 0; i < SG_WIN / SG_THREADS; i++) {
			const u32 j = L + i * SG_THREADS + t;
			if (j < lim && (AP[j] & AP_HEAD)) atomicMin(&sh[2], j);
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
template <bool PS>
__global__ void __launch_bounds__(SG_THREADS, 5) k_seg_sort(const u32* __restrict__ AP, u32* __restrict__ SA, u32 A,
                                                         const u32* __restrict__ ISA, u32 h, u32 n,
                                                         u8* __restrict__ F, u32* __restrict__ counters /*[2]=has_large [3]=queued*/,
                                                         u32* __restrict__ queue, u32* __restrict__ win_first, u32* __restrict__ win_large,
                                                         int* __restrict__ err, PerSkip ps)
{
	extern __shared__ __align__(16) u8 sg_smem[];
	u32* skey = reinterpret_cast<u32*>(sg_smem);        // key2; later the refined suffix ids
	u32* sval = skey + SG_CAP;                          // suffix ids in slot order
	u8* sflag = reinterpret_cast<u8*>(sval + SG_CAP);   // head flags of the refined order
	__shared__ u32 sh[3];
	__shared__ u32 wf[SG_ITEMS][SG_THREADS / 32], wb[SG_ITEMS][SG_THREADS / 32];
	const int t = threadIdx.x, lane = t & 31, w = t >> 5;
	const SegTile tile = seg_range(blockIdx.x, AP, A, sh, counters + 2, win_first, win_large);
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
			const u32 ap = AP[j];
			const bool head = (e == 0) || (ap & AP_HEAD);
			const bool last = (e == len - 1) || (AP[j + 1] & AP_HEAD);
			const u32 v = SA[ap & AP_POS];
			sval[e] = v;
			skey[e] = key2_of<PS>(v, h, n, ISA, ps, err);
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
		if (e < len) { SA[AP[start + e] & AP_POS] = skey[e]; F[start + e] = sflag[e]; }
	}
}

// Radix route for the queued tiles: LSD radix sort of (local group number, key2) that never leaves shared memory.
template <bool PS>
__global__ void __launch_bounds__(SG_THREADS) k_seg_sort_radix(const u32* __restrict__ AP, u32* __restrict__ SA, u32 A,
                                                               const u32* __restrict__ ISA, u32 h, u32 n, int rank_bits,
                                                               u8* __restrict__ F, const u32* __restrict__ queue, int* __restrict__ err, PerSkip ps)
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
	const SegTile tile = seg_range(queue[blockIdx.x], AP, A, sh, nullptr);
	const u32 len = tile.len, start = tile.start;
	if (len == 0) return;
	#pragma unroll 4
	for (int i = 0; i < SG_ITEMS; i++) {
		const u32 e = i * SG_THREADS + t;
		if (e < len) {
			const u32 ap = AP[start + e];
			const u32 v = SA[ap & AP_POS];
			sval[e] = v;
			skey[e] = ((u64)((e == 0 || (ap & AP_HEAD)) ? 1u : 0u) << 32) | (u64)key2_of<PS>(v, h, n, ISA, ps, err);
		}
	}
	__syncthreads();
	// local group number = heads at or before the element, minus one: each thread counts its 16 consecutive slots
	u32 gmax;
	{
		u32 mine = 0;
		#pragma unroll
		for (int k = 0; k < SG_ITEMS; k++) { const u32 e = t * SG_ITEMS + k; if (e < len) mine += (u32)(skey[e] >> 32); }
		u32 total;
		u32 g = block_incl_sum(mine, ws, &total) - mine;
		#pragma unroll
		for (int k = 0; k < SG_ITEMS; k++) {
			const u32 e = t * SG_ITEMS + k;
			if (e < len) { const u64 c = skey[e]; g += (u32)(c >> 32); skey[e] = ((u64)(g - 1) << 32) | (c & 0xffffffffull); }
		}
		gmax = total - 1;
		__syncthreads();
	}
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
			SA[AP[start + e] & AP_POS] = val[i];
			F[start + e] = (e == 0 || skey[e - 1] != key[i]) ? 1 : 0;
		}
	}
}

// ---- 5b. large groups -----------------------------------------------------------------------------------
// A group longer than a window (common prefixes of real text, periodic data) cannot be sorted inside one block.
// The windows report where such groups start; k_large_collect finds where they end (the next group head of any
// later window) and lays them out back to back; their suffixes are extracted with key (large-group number, key2),
// sorted by the global radix sort, and written back in place with their head flags.
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

template <bool PS>
__global__ void __launch_bounds__(256) k_large_extract(const u32* __restrict__ AP, const u32* __restrict__ SA, const u32* __restrict__ ISA, u32 h, u32 n, int rank_bits,
                                                       const u32* __restrict__ lg_head, const u32* __restrict__ lg_off, u32 ng, u32 total,
                                                       u64* __restrict__ LK, u32* __restrict__ LV, int* __restrict__ err, PerSkip ps)
{
	const u32 x = blockIdx.x * 256 + threadIdx.x;
	if (x >= total) return;
	const u32 g = large_group_of(lg_off, ng, x);
	const u32 v = SA[AP[lg_head[g] + (x - lg_off[g])] & AP_POS];
	const u32 k2 = key2_of<PS>(v, h, n, ISA, ps, err);
	LK[x] = ((u64)g << rank_bits) | (u64)k2;
	LV[x] = v;
}

__global__ void __launch_bounds__(256) k_large_writeback(const u64* __restrict__ LK, const u32* __restrict__ LV, int rank_bits,
                                                         const u32* __restrict__ lg_head, const u32* __restrict__ lg_off, u32 total,
                                                         const u32* __restrict__ AP, u32* __restrict__ SA, u8* __restrict__ F)
{
	const u32 x = blockIdx.x * 256 + threadIdx.x;
	if (x >= total) return;
	const u64 k = LK[x];
	const u32 g = (u32)(k >> rank_bits);
	const u32 o = lg_off[g];
	const u32 slot = lg_head[g] + (x - o);
	SA[AP[slot] & AP_POS] = LV[x];
	F[slot] = (x == o || LK[x - 1] != k) ? 1 : 0;
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
	size_t usz;                // bytes per unit: room for N + 2 32-bit words
	u8* unit[6];
	RadixBuffers rb;           // initial sort: k[0] = units 0-1, k[1] = units 2-3, v[0] = unit 4, v[1] = unit 5
	u32* SA; u32* ISA; u32* VS; u32* X; u32* AP[2];   // roles of the six units once the initial sort is done
	u8* F;                     // one byte per slot (next-digit bytes, then head flags): the caller's output block
	int isa_region_log2;
	u32* queue;
	u32* win_first; u32* win_large; u32* lg_head; u32* lg_off;
	GAgg* agg;
	FwdMeta* meta;
	u32* counters;
	int* err;
};

static size_t fwd_bytes(i32 n, bool own_flags)
{
	const size_t N = (size_t)n;
	const size_t rtiles = radix_tiles(N), gtiles = (N + GS_TILE - 1) / GS_TILE;
	return 6 * Arena::align((N + 2) * 4) + Arena::align((rtiles + 4) * 256 * 4) + Arena::align(256 * 4) + Arena::align((8 * 256 + 64) * 4) +
	       Arena::align(gtiles * sizeof(GAgg)) + Arena::align(sizeof(FwdMeta)) + Arena::align(64) + Arena::align(64) +
	       5 * Arena::align((N / SG_WIN + 16) * 4) + (own_flags ? Arena::align(N + 64) : 0);
}

// d_flags: N bytes of scratch that stay untouched until the suffix array is complete (the output block), or nullptr
static int fwd_alloc(Ctx& c, i32 n, FwdBuffers& b, u8* d_flags)
{
	const size_t N = (size_t)n;
	const size_t rtiles = radix_tiles(N), gtiles = (N + GS_TILE - 1) / GS_TILE;
	JP_TRY(arena_reserve(c, fwd_bytes(n, d_flags == nullptr)));
	b.usz = Arena::align((N + 2) * 4);
	for (int i = 0; i < 6; i++) b.unit[i] = arena_take<u8>(c, b.usz);
	b.rb.k[0] = reinterpret_cast<u64*>(b.unit[0]); b.rb.k[1] = reinterpret_cast<u64*>(b.unit[2]);
	b.rb.v[0] = reinterpret_cast<u32*>(b.unit[4]); b.rb.v[1] = reinterpret_cast<u32*>(b.unit[5]);
	b.isa_region_log2 = 24;                            // 2^24 ranks = 64 MiB of ISA per pass (measured best of 2^21..2^25)
	if (const char* e = getenv("JP_BWT_ISA_REGION_LOG2")) b.isa_region_log2 = atoi(e);
	b.rb.tile_hist = arena_take<u32>(c, (rtiles + 4) * 256);
	b.rb.totals = arena_take<u32>(c, 256);
	b.rb.os_state = arena_take<u32>(c, 8 * 256 + 64);
	b.F = d_flags ? d_flags : arena_take<u8>(c, N + 64);
	b.rb.dnext = getenv("JP_BWT_RADIX_NO_DIGIT_BYTES") ? nullptr : b.F;
	// Measured on B200 (64 M pairs, 8 passes): three-kernel passes 5.07 ms, one-sweep (all digits counted in one read of
	// the keys + 8 look-back passes) 6.18 ms -- with ~440 tiles in flight the per-digit look-back chain costs more than
	// the key re-read it saves, so the classic pass is the default.
	b.rb.classic = getenv("JP_BWT_RADIX_ONESWEEP") == nullptr;
	b.queue = arena_take<u32>(c, N / SG_WIN + 16);
	b.win_first = arena_take<u32>(c, N / SG_WIN + 16); b.win_large = arena_take<u32>(c, N / SG_WIN + 16);
	b.lg_head = arena_take<u32>(c, N / SG_WIN + 16); b.lg_off = arena_take<u32>(c, N / SG_WIN + 16);
	b.agg = arena_take<GAgg>(c, gtiles);
	b.meta = arena_take<FwdMeta>(c, 1);
	b.counters = arena_take<u32>(c, 16);
	b.err = arena_take<int>(c, 16);
	b.rb.err = b.err;
	b.SA = b.ISA = b.VS = b.X = b.AP[0] = b.AP[1] = nullptr;
	return JP_OK;
}

// ISA[V[j]] = R[j] for the A slots, staged by region (k_isa_scatter). `pv`/`pr`: room for `pcap` entries each, used when the
// block has too many regions for per-region sweeps: the pairs are bucketed by region first, `pcap` slots at a time.
static int place_ranks(Ctx& c, FwdBuffers& b, const u32* V, const u32* R, u32 A, u32* pv, u32* pr, u32 pcap, cudaStream_t s)
{
	const u32 n_entries = c.cur_n + 1, regions = (n_entries + (1u << b.isa_region_log2) - 1) >> b.isa_region_log2;
	if (regions <= 4 || regions > 256) {
		// few regions: stream the slots once per region and keep the ranks that fall in it
		const u32 stiles = (A + 2047) / 2048;
		k_isa_scatter<<<regions * stiles, 256, 0, s>>>(V, R, A, b.ISA, b.isa_region_log2, stiles); JP_LAUNCH(c);
		return JP_OK;
	}
	// many regions (blocks over 64 MiB): one radix partition pass buckets the (suffix, rank) pairs by region, then a
	// single ordered sweep
	for (u32 off = 0; off < A; off += pcap) {
		const u32 cnt = std::min(pcap, A - off);
		if (radix_partition_u32(V + off, R + off, pv, pr, cnt, b.isa_region_log2, b.rb.tile_hist, b.rb.totals, s, &c.launches) != 0) { set_error_detail("radix partition setup failed"); return JP_ERR_CUDA; }
		const u32 stiles = (cnt + 2047) / 2048;
		k_isa_scatter<<<stiles, 256, 0, s>>>(pv, pr, cnt, b.ISA, 32, stiles); JP_LAUNCH(c);
	}
	return JP_OK;
}

// One grouping step over the A slots of APin (nullptr: the N positions themselves) with head flags b.F; survivors land in
// APout. Leaves the survivor and group counts in h_small[8..9] and the device error flag in h_small[0].
static int group_step(Ctx& c, FwdBuffers& b, const u32* APin, u32* APout, u32 A, cudaStream_t s)
{
	const int tiles = (int)((A + GS_TILE - 1) / GS_TILE);
	k_grp_reduce<<<tiles, GS_THREADS, 0, s>>>(b.F, APin, A, b.agg); JP_LAUNCH(c);
	k_grp_scan_tiles<<<1, 1024, 0, s>>>(b.agg, tiles, b.counters); JP_LAUNCH(c);
	// small active sets (and the A/B switch region_log2 <= 0) scatter straight from the apply kernel
	const u32 stage_min = getenv("JP_BWT_ISA_STAGE_MIN") ? (u32)atol(getenv("JP_BWT_ISA_STAGE_MIN")) : (1u << 20);   // (tests lower it)
	const bool staged = b.isa_region_log2 > 0 && A > stage_min;
	if (!staged) { k_grp_apply<<<tiles, GS_THREADS, 0, s>>>(b.F, APin, b.SA, A, b.agg, b.ISA, nullptr, nullptr, APout); JP_LAUNCH(c); }
	else if (APin == nullptr) {
		// initial step: the suffix ids are SA itself; ranks go to the spare unit; VS and the second active buffer are free
		k_grp_apply<<<tiles, GS_THREADS, 0, s>>>(b.F, nullptr, b.SA, A, b.agg, b.ISA, b.X, nullptr, APout); JP_LAUNCH(c);
		JP_TRY(place_ranks(c, b, b.SA, b.X, A, b.VS, APout == b.AP[0] ? b.AP[1] : b.AP[0], (u32)(b.usz / 4), s));
	} else {
		// a round: ranks overwrite the (dead) input slots, suffix ids are staged in VS; the spare unit holds the bucketed pairs
		u32* R = const_cast<u32*>(APin);
		k_grp_apply<<<tiles, GS_THREADS, 0, s>>>(b.F, APin, b.SA, A, b.agg, b.ISA, R, b.VS, APout); JP_LAUNCH(c);
		const u32 half = (u32)(b.usz / 8);
		JP_TRY(place_ranks(c, b, b.VS, R, A, b.X, b.X + half, half, s));
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
	JP_CUDA(cudaEventRecord(c.ev[1], s));
	if (cudaFuncSetAttribute(k_seg_sort<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG_SMEM_LIGHT) != cudaSuccess ||
	    cudaFuncSetAttribute(k_seg_sort_radix<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG_SMEM) != cudaSuccess ||
	    cudaFuncSetAttribute(k_seg_sort<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG_SMEM_LIGHT) != cudaSuccess ||
	    cudaFuncSetAttribute(k_seg_sort_radix<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG_SMEM) != cudaSuccess) { set_error_detail("k_seg_sort smem attribute"); return JP_ERR_CUDA; }
	const int cur = radix_sort_pairs(b.rb, 0, (u32)n, 0, key_bits0, s, &c.launches, /*first_hist_ready=*/true);
	if (cur < 0) { set_error_detail("radix sort setup failed"); return JP_ERR_CUDA; }
	JP_KCHECK();
	JP_CUDA(cudaEventRecord(c.ev[2], s));

	// The sorted suffix ids are the suffix array from here on; the key buffers and the other id buffer take new roles
	// (the keys are read one last time, for the head flags).
	b.SA = b.rb.v[cur];
	k_fwd_flags<<<(u32)(((size_t)n + 1023) / 1024), 256, 0, s>>>(b.rb.k[cur], (u32)n, b.F); JP_LAUNCH(c);
	b.ISA = reinterpret_cast<u32*>(b.unit[cur ? 0 : 2]);      // first half of the other key buffer
	b.X = reinterpret_cast<u32*>(b.unit[cur ? 1 : 3]);        // ... and its second half
	b.AP[0] = reinterpret_cast<u32*>(b.unit[cur ? 2 : 0]);    // the sorted keys' own buffer, dead once the flags exist
	b.AP[1] = reinterpret_cast<u32*>(b.unit[cur ? 3 : 1]);
	b.VS = b.rb.v[cur ^ 1];
	JP_CUDA(cudaMemsetAsync(b.ISA + n, 0, sizeof(u32), s));             // the empty suffix ranks below everything
	const int rank_bits = bit_length((u64)n);                           // key2 is a rank <= n
	JP_TRY(group_step(c, b, nullptr, b.AP[0], (u32)n, s));
	PerSkip ps; ps.rl = nullptr; ps.bits = nullptr; ps.T = d_T; ps.p = 1; ps.first = 0;
	JP_CUDA(cudaEventRecord(c.ev[3], s));
	int act = 0;
	u32 A = (u32)c.h_small[8], G = (u32)c.h_small[9];
	u64 sectors = 2ull * (u64)n;
	i64 h = depth;
	int rounds = 0;
	const bool trace_rounds = getenv("JP_BWT_TRACE_ROUNDS") != nullptr;      // one stderr line per doubling round (host wall time)
	double t_round = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
	while (A > 0) {
		if (c.h_small[0] != 0) return map_dev_err(c.h_small[0]);
		if (rounds >= JP_BWT_MAX_ROUNDS || h > (i64)n) { set_error_detail("doubling stuck: round %d h=%lld active=%u", rounds, (long long)h, A); return JP_ERR_INTERNAL; }
		st->active_fraction[rounds] = (float)((double)A / (double)n);
		sectors += 2ull * A;
		double large_frac_now = 0.0;
		const u32* AP = b.AP[act];
		// short groups: fused gather + warp-level rank refinement in shared memory; longer ones are queued for the
		// shared-memory radix kernel; groups longer than a window are reported for the large-group route
		const u32 nwin = (A + SG_WIN - 1) / SG_WIN;
		JP_CUDA(cudaMemsetAsync(b.counters + 2, 0, 4 * sizeof(u32), s));
		k_seg_sort<false><<<nwin, SG_THREADS, SG_SMEM_LIGHT, s>>>(AP, b.SA, A, b.ISA, (u32)h, (u32)n, b.F, b.counters, b.queue, b.win_first, b.win_large, b.err, ps);
		JP_LAUNCH(c);
		JP_KCHECK();
		JP_CUDA(cudaMemcpyAsync(c.h_small + 10, b.counters + 2, 2 * sizeof(u32), cudaMemcpyDeviceToHost, s));
		JP_CUDA(cudaStreamSynchronize(s));
		const bool large = c.h_small[10] != 0;
		const u32 queued = (u32)c.h_small[11];
		if (queued) {
			k_seg_sort_radix<false><<<queued, SG_THREADS, SG_SMEM, s>>>(AP, b.SA, A, b.ISA, (u32)h, (u32)n, rank_bits, b.F, b.queue, b.err, ps);
			JP_LAUNCH(c);
			JP_KCHECK();
			st->radix_tiles += (i32)queued;
		}
		if (large) {
			k_large_collect<<<1, 1024, 0, s>>>(b.win_first, b.win_large, nwin, A, b.lg_head, b.lg_off, b.counters); JP_LAUNCH(c);
			JP_KCHECK();
			JP_CUDA(cudaMemcpyAsync(c.h_small + 12, b.counters + 4, 2 * sizeof(u32), cudaMemcpyDeviceToHost, s));
			JP_CUDA(cudaStreamSynchronize(s));
			const u32 ng = (u32)c.h_small[12], total = (u32)c.h_small[13];
			st->large_fraction += (float)((double)total / (double)n);   // share of the block on the large-group route, summed over rounds
			large_frac_now = (double)total / (double)A;
			if (ng == 0 || total == 0 || total > A) { set_error_detail("large-group list inconsistent: %u groups, %u suffixes, %u active", ng, total, A); return JP_ERR_INTERNAL; }
			// scratch of the sort: the spare unit, the staging unit and the idle active buffer hold (8 + 8 + 2 x 4) bytes for
			// up to N/2 suffixes; beyond that (blocks that are mostly long repeats) the second arena provides
			RadixBuffers lb = b.rb;
			lb.dnext = nullptr;                             // the flag bytes of this round are live in b.F
			if ((size_t)total * 8 <= b.usz) {
				lb.k[0] = reinterpret_cast<u64*>(b.X); lb.k[1] = reinterpret_cast<u64*>(b.VS);
				lb.v[0] = b.AP[act ^ 1]; lb.v[1] = b.AP[act ^ 1] + b.usz / 8;
			} else {
				const size_t T8 = Arena::align((size_t)total * 8), T4 = Arena::align((size_t)total * 4);
				JP_TRY(arena2_reserve(c, 2 * T8 + 2 * T4));
				lb.k[0] = reinterpret_cast<u64*>(c.arena2.base); lb.k[1] = reinterpret_cast<u64*>(c.arena2.base + T8);
				lb.v[0] = reinterpret_cast<u32*>(c.arena2.base + 2 * T8); lb.v[1] = reinterpret_cast<u32*>(c.arena2.base + 2 * T8 + T4);
			}
			k_large_extract<false><<<(total + 255) / 256, 256, 0, s>>>(AP, b.SA, b.ISA, (u32)h, (u32)n, rank_bits, b.lg_head, b.lg_off, ng, total,
			                                                    lb.k[0], lb.v[0], b.err, ps);
			JP_LAUNCH(c);
			const int key_bits = rank_bits + bit_length((u64)(ng - 1));
			const int lc = radix_sort_pairs(lb, 0, total, 0, key_bits, s, &c.launches);
			if (lc < 0) { set_error_detail("radix sort setup failed"); return JP_ERR_CUDA; }
			k_large_writeback<<<(total + 255) / 256, 256, 0, s>>>(lb.k[lc], lb.v[lc], rank_bits, b.lg_head, b.lg_off, total, AP, b.SA, b.F); JP_LAUNCH(c);
			JP_KCHECK();
		}
		JP_TRY(group_step(c, b, AP, b.AP[act ^ 1], A, s));
		if (trace_rounds) {
			const double t1 = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
			fprintf(stderr, "[jp_bwt round] %d h=%lld active=%u groups=%u -> active=%u groups=%u large_frac=%.4f %.3f ms\n", rounds, (long long)h, A, G,
			        (u32)c.h_small[8], (u32)c.h_small[9], large_frac_now, t1 - t_round);
			t_round = t1;
		}
		act ^= 1;
		A = (u32)c.h_small[8]; G = (u32)c.h_small[9];
		h *= 2;
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
	JP_TRY(fwd_alloc(c, nlen, b, d_out));                               // the output block is scratch until the emission
	JP_TRY(suffix_sort(c, d_in, nlen, b, s, st));
	k_fwd_emit<<<(int)(((i64)nlen + 1023) / 1024), 256, 0, s>>>(d_in, b.SA, b.ISA, nlen, d_out); JP_LAUNCH(c);
	k_fwd_trailer<<<1, 128, 0, s>>>(d_in, b.ISA, nlen, len, d_out); JP_LAUNCH(c);
	JP_KCHECK();
	JP_CUDA(cudaEventRecord(c.ev[5], s));
	JP_CUDA(cudaStreamSynchronize(s));
	for (int i = 0; i < 5; i++) JP_CUDA(cudaEventElapsedTime(&st->ms_phase[i], c.ev[i], c.ev[i + 1]));
	JP_CUDA(cudaEventElapsedTime(&st->ms_total, c.ev[0], c.ev[5]));
	st->device_bytes = c.arena.high + c.arena2.cap;
	return JP_OK;
}

int debug_suffix_array(Ctx& c, const u8* h_in, i32 n, i32* h_sa)
{
	if (n == 0) return JP_OK;
	cudaStream_t s = c.own_stream;
	c.arena.reset();
	FwdBuffers b;
	const size_t N = (size_t)n;
	JP_TRY(arena_reserve(c, Arena::align(N + 16) + fwd_bytes(n, true)));   // text + workspace in one reservation
	u8* d_T = arena_take<u8>(c, N + 16);
	JP_TRY(fwd_alloc(c, n, b, nullptr));
	JP_CUDA(cudaMemcpyAsync(d_T, h_in, N, cudaMemcpyHostToDevice, s));
	JP_CUDA(cudaEventRecord(c.ev[0], s));
	jp_bwt_stats st = {};
	JP_TRY(suffix_sort(c, d_T, n, b, s, &st));
	JP_CUDA(cudaMemcpyAsync(h_sa, b.SA, N * 4, cudaMemcpyDeviceToHost, s));
	JP_CUDA(cudaStreamSynchronize(s));
	return JP_OK;
}

} // namespace jp
