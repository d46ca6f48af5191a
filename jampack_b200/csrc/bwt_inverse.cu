// bwt_inverse.cu -- inverse BWT for sm_100a.
//
// Replaces BlockSort::Bwt::InverseBwt (reference bwt.cpp:72-282): byte histogram + C table
// (bwt.cpp:141-169), the LF/rank table (the inverse permutation of the reference's Map, bwt.cpp:171-174)
// and the multi-index walk (bwt.cpp:176-183, :261-275) -- restructured for a machine that needs
// >= 10^4 loads in flight rather than 120 (SURVEY.md finding 5, Appendix D):
//
//   rows        the nlen+1 rows of the sentinel-inclusive BWT matrix; row 0 is the empty suffix, row idx
//               (= stored index 0) is suffix 0, whose L symbol is the dropped '$'. Byte i of the stored
//               BWT belongs to row i + (i >= idx).
//   lf[i]       = 1 + C[bwt[i]] + #{j < i : bwt[j] == bwt[i]} : the row of the suffix one position to the
//               left. The symbol stepped over is NOT fetched from the BWT: it is the F-column symbol of
//               lf[i], found in the 257-entry C table (shared memory). One random 32 B sector per byte.
//   marks       bit 31 of lf[i] (row numbers need <= 30 bits, format.hpp:22) flags row(i) as the start of a
//               sub-chain: one hashed row in every window of m rows, plus the 120 rows the stored indices name
//               (text positions k*step) and row 0 (text position nlen) -- the anchors, which carry bit 30 as well.
//   two-pass    (blocks under 30 Mi)
//     pass 1    one walker per sub-chain start: walk left until the next mark; record (next start, length).
//     ranking   asynchronous pointer jumping over the (next, length) records; the 121 anchors absorb. Every
//               sub-chain learns which decode unit it belongs to and its offset from that unit's start, so all
//               120 decode units of the format run concurrently, each cut into ~step/m independent pieces.
//     pass 2    the same walk again, now writing bytes right-to-left at the known text offset, packed into
//               aligned 16-byte windows.
//   single-walk (blocks of 30 Mi and more) the walk is done once: bytes are parked in per-warp streams, and after
//               the ranking a replay of each warp's loop puts them in place (see "single-walk path" below).
//
// Device memory: caller's in (N+480) + out (N) + lf (4N) = 6N, plus o(N): 8 B per sub-chain (N/2 at m = 16, kept in
// the consumed input block) and 1 KiB per 64 KiB tile (N/64); the single-walk streams live in `out` and the free half
// of `in`.
#include "bwt_internal.cuh"
#include <algorithm>

namespace jp {

constexpr int  INV_THREADS    = 256;
constexpr int  INV_WARPS      = INV_THREADS / 32;
constexpr int  INV_TILE       = 65536;          // bytes per histogram / LF tile
constexpr int  INV_CHUNK      = INV_WARPS * 512; // bytes ranked per block iteration of the LF build
constexpr u32  LF_MARK        = 0x80000000u;     // row starts a sub-chain
constexpr u32  LF_ANCHOR      = 0x40000000u;     // ... and is one of the 121 anchor rows (set together with LF_MARK)
constexpr u32  LF_MASK        = 0x3fffffffu;     // row numbers need 30 bits: nlen <= 1000 * 2^20 < 2^30 (format.hpp:22)
static_assert((u64)JP_BWT_MAX_CALL_LEN + 1 <= (u64)LF_MASK, "row numbers of every accepted block must leave bits 30 and 31 free");
constexpr int  N_ANCHOR       = JP_BWT_UNITS + 1; // anchor k = row of text position k*step, k = 0..120
constexpr u32  REC_INVALID    = 0xffffffffu;
constexpr int  RANK_HOP_CAP   = 1 << 22;

struct InvMeta {
	i32 idx;                       // stored index 0 = row of suffix 0
	i32 anchor_row[N_ANCHOR];      // [0] = idx (terminal), [k] = stored index k, [120] = 0
	i32 sorted_row[N_ANCHOR];      // anchor rows ascending ...
	i32 sorted_id[N_ANCHOR];       // ... and which anchor each one is
	i32 ctable[257];               // C[c] = #{bytes < c}; C[256] = nlen
};

// The marked row of window w (rows [w << log2m, (w + 1) << log2m)): the top log2m bits of a multiplicative hash. It
// only has to be unrelated to the structure of the text; two instructions, evaluated once per byte by the LF build
// and once per sub-chain by the walkers (log2m >= 2).
__device__ __forceinline__ u32 mark_slot(u32 window, int log2m) { return (window * 0x9E3779B1u) >> (32 - log2m); }

// ---- prepare: read + validate the 120 stored indices, sort the anchors, copy the raw tail ---------
// bwt.cpp:80-89. The trailer sits at byte offset Len, unaligned, native-endian (little on every CUDA host).
__global__ void k_inv_prepare(const u8* __restrict__ in, i32 len, i32 nlen, u8* __restrict__ out,
                              InvMeta* __restrict__ meta, int* __restrict__ err)
{
	__shared__ i32 row[N_ANCHOR + 3];   // padded: the rank loop below is vectorised into 8-byte reads
	const int t = threadIdx.x;
	if (t < JP_BWT_UNITS) {
		const u8* p = in + len + 4 * t;
		u32 v = (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24);
		i32 r = (i32)v;
		if (r < 1 || r > nlen) { dev_fail(err, DE_BAD_INDEX); r = 1; }
		row[t] = r;
	}
	if (t == JP_BWT_UNITS) row[t] = 0;
	__syncthreads();
	if (t < N_ANCHOR) {
		const i32 mine = row[t];
		int rank = 0;
		for (int j = 0; j < N_ANCHOR; j++) {
			const i32 o = row[j];
			if (o < mine || (o == mine && j < t)) rank++;
			if (o == mine && j != t) dev_fail(err, DE_BAD_INDEX);   // anchors must be distinct rows
		}
		meta->anchor_row[t] = mine;
		meta->sorted_row[rank] = mine;
		meta->sorted_id[rank] = t;
		if (t == 0) meta->idx = mine;
	}
	for (int i = t; i < len - nlen; i += blockDim.x) out[nlen + i] = in[nlen + i];   // bwt.cpp:82-83
}

// ---- byte histogram per 64 KiB tile (bwt.cpp:141-167) --------------------------------------------
// BWT output is run-heavy, so same-address shared atomics would serialise; each thread folds the runs
// inside its own 16 bytes before touching the per-warp histogram.
__global__ void __launch_bounds__(INV_THREADS) k_inv_hist(const u8* __restrict__ bwt, i32 n, u32* __restrict__ tile_hist)
{
	__shared__ u32 h[INV_WARPS][256];
	const int t = threadIdx.x, w = t >> 5;
	for (int i = t; i < INV_WARPS * 256; i += INV_THREADS) (&h[0][0])[i] = 0;
	__syncthreads();
	const i64 base = (i64)blockIdx.x * INV_TILE;
	const i64 end = min((i64)n, base + INV_TILE);
	u32* hw = h[w];
	for (i64 p = base + (i64)t * 16; p < end; p += INV_THREADS * 16) {
		if (p + 16 <= end) {
			const uint4 q = __ldg(reinterpret_cast<const uint4*>(bwt + p));
			const u32 wd[4] = {q.x, q.y, q.z, q.w};
			u32 cur = wd[0] & 255, run = 0;
			#pragma unroll
			for (int k = 0; k < 16; k++) {
				const u32 c = (wd[k >> 2] >> ((k & 3) * 8)) & 255;
				if (c != cur) { atomicAdd(&hw[cur], run); cur = c; run = 0; }
				run++;
			}
			atomicAdd(&hw[cur], run);
		} else {
			for (i64 q = p; q < end; q++) atomicAdd(&hw[bwt[q]], 1u);
		}
	}
	__syncthreads();
	u32 s = 0;
	#pragma unroll
	for (int k = 0; k < INV_WARPS; k++) s += h[k][t];
	tile_hist[(size_t)blockIdx.x * 256 + t] = s;
}

// ---- per-symbol exclusive scan down the tiles; block b owns symbol b ------------------------------
__global__ void __launch_bounds__(256) k_inv_scan_tiles(u32* __restrict__ tile_hist, int tiles, u32* __restrict__ bin_total)
{
	__shared__ u32 ws[32];
	const int b = blockIdx.x, t = threadIdx.x;
	const int per = (tiles + 255) / 256;
	const int lo = min(tiles, t * per), hi = min(tiles, lo + per);
	u32 s = 0;
	for (int k = lo; k < hi; k++) s += tile_hist[(size_t)k * 256 + b];
	u32 total;
	u32 inc = block_incl_sum(s, ws, &total);
	u32 run = inc - s;
	for (int k = lo; k < hi; k++) {
		const size_t a = (size_t)k * 256 + b;
		const u32 v = tile_hist[a];
		tile_hist[a] = run;
		run += v;
	}
	if (t == 0) bin_total[b] = total;
}

// ---- C table: exclusive prefix sum of the 256 totals (bwt.cpp:168-169) ----------------------------
__global__ void __launch_bounds__(256) k_inv_ctable(const u32* __restrict__ bin_total, InvMeta* __restrict__ meta, i32 n)
{
	__shared__ u32 ws[32];
	const int t = threadIdx.x;
	const u32 v = bin_total[t];
	u32 total;
	const u32 inc = block_incl_sum(v, ws, &total);
	meta->ctable[t] = (i32)(inc - v);
	if (t == 255) meta->ctable[256] = (i32)inc;   // == n
	(void)n;
}

// ---- LF/rank table build --------------------------------------------------------------------------
// One block per 64 KiB tile, 4 KiB per iteration. Each warp owns 512 consecutive bytes, read as four
// coalesced 128 B rows of 32-bit words and re-distributed by shuffle so that lane l handles byte
// 32*k + l of the row (text order == lane order). Equal bytes inside a warp step are ranked with
// match.any; running per-warp counters live in shared memory; LF values leave as 128 B coalesced rows.
// 4 resident blocks per SM asked of the compiler: that caps the kernel at 64 registers with 16 bytes of spill and was
// measured at 0.185 ms per 64 MiB block against 0.222 ms for the 80-register / 3-block build (profiles/ab_r02.md).
__global__ void __launch_bounds__(INV_THREADS, 4) k_inv_lf(const u8* __restrict__ bwt, i32 n,
                                                        const u32* __restrict__ tile_excl, const InvMeta* __restrict__ meta,
                                                        u32* __restrict__ lf, int log2m)
{
	__shared__ u32 wcnt[INV_WARPS][256];
	__shared__ u32 basec[256];
	const int t = threadIdx.x, w = t >> 5, lane = t & 31;
	const u32 lt = lanemask_lt();
	const u32 mmask = (1u << log2m) - 1;
	basec[t] = 1u + (u32)meta->ctable[t] + tile_excl[(size_t)blockIdx.x * 256 + t];
	const i64 tile_base = (i64)blockIdx.x * INV_TILE;
	const i64 tile_end = min((i64)n, tile_base + INV_TILE);

	for (i64 cb = tile_base; cb < tile_end; cb += INV_CHUNK) {
		for (int i = t; i < INV_WARPS * 256; i += INV_THREADS) (&wcnt[0][0])[i] = 0;
		__syncthreads();

		const i64 seg = cb + (i64)w * 512;
		u32 word[4];
		#pragma unroll
		for (int r = 0; r < 4; r++) {
			const i64 p = seg + r * 128 + lane * 4;
			u32 v = 0;
			if (p + 4 <= n) v = __ldg(reinterpret_cast<const u32*>(bwt + p));
			else { for (int b = 0; b < 4; b++) if (p + b < n) v |= (u32)bwt[p + b] << (8 * b); }
			word[r] = v;
		}
		u32 lrank[16];
		u32* mycnt = wcnt[w];
		#pragma unroll
		for (int it = 0; it < 16; it++) {
			const int r = it >> 2, k = it & 3;
			const u32 src = __shfl_sync(0xffffffffu, word[r], 8 * k + (lane >> 2));
			const i64 gp = seg + r * 128 + k * 32 + lane;
			const bool valid = gp < n;
			const u32 c = valid ? ((src >> (8 * (lane & 3))) & 255u) : 0u;
			const u32 vmask = __ballot_sync(0xffffffffu, valid);
			const u32 same = match_any8_adaptive<28>(c);                // every lane takes part
			const u32 peers = valid ? (same & vmask) : ~vmask;       // bytes past the end only match each other
			const u32 below = __popc(peers & lt);
			u32 before = 0;
			if (below == 0 && valid) { before = mycnt[c]; mycnt[c] = before + __popc(peers); }
			before = __shfl_sync(0xffffffffu, before, __ffs(peers) - 1);
			lrank[it] = before + below;
			__syncwarp();
		}
		__syncthreads();
		{   // exclusive scan across the warps of this chunk, folded into the running per-symbol base
			u32 run = basec[t];
			#pragma unroll
			for (int k = 0; k < INV_WARPS; k++) { const u32 v = wcnt[k][t]; wcnt[k][t] = run; run += v; }
			basec[t] = run;
		}
		__syncthreads();
		#pragma unroll
		for (int it = 0; it < 16; it++) {
			const int r = it >> 2, k = it & 3;
			const u32 src = __shfl_sync(0xffffffffu, word[r], 8 * k + (lane >> 2));
			const i64 gp = seg + r * 128 + k * 32 + lane;
			if (gp < n) {
				const u32 c = (src >> (8 * (lane & 3))) & 255u;
				u32 v = mycnt[c] + lrank[it];
				const u32 gi = (u32)gp;
				if ((gi & mmask) == mark_slot(gi >> log2m, log2m)) v |= LF_MARK;
				lf[gp] = v;
			}
		}
		__syncthreads();
	}
}

__device__ __forceinline__ i32 row_to_byte(i32 row, i32 idx) { return row - (row > idx ? 1 : 0); }

// ---- mark the anchors (stored indices 1..119 and row 0) -------------------------------------------
__global__ void k_inv_mark_anchors(const InvMeta* __restrict__ meta, u32* __restrict__ lf, i32 n)
{
	const int k = threadIdx.x;
	if (k >= 1 && k < N_ANCHOR) {
		const i32 idx = meta->idx;
		i32 i = row_to_byte(meta->anchor_row[k], idx);
		if (i < 0) i = 0;
		if (i >= n) i = n - 1;          // only reachable with indices already flagged as bad
		atomicOr(&lf[i], LF_MARK | LF_ANCHOR);
	}
}

// ---- walkers ---------------------------------------------------------------------------------------
struct AnchorTable {
	i32 row[N_ANCHOR];             // anchor rows ascending ...
	i32 id[N_ANCHOR];              // ... and which anchor each one is
};

__device__ __forceinline__ void load_anchor_table(AnchorTable& a, const InvMeta* meta)
{
	for (int i = threadIdx.x; i < N_ANCHOR; i += blockDim.x) { a.row[i] = meta->sorted_row[i]; a.id[i] = meta->sorted_id[i]; }
}

// node id of the marked row `row` (byte index bi) whose LF entry is v: an anchor when the entry says so (121 rows in
// the whole block: the search below is off the common path), else the window's own node
__device__ __forceinline__ u32 node_of(const AnchorTable& a, u32 v, i32 row, i32 bi, int log2m, u32 S)
{
	if ((v & LF_ANCHOR) == 0) return (u32)bi >> log2m;
	int lo = 0, hi = N_ANCHOR;          // first position with a.row[pos] >= row
	while (lo < hi) { const int mid = (lo + hi) >> 1; if (a.row[mid] < row) lo = mid + 1; else hi = mid; }
	if (lo < N_ANCHOR && a.row[lo] == row) return S + (u32)a.id[lo];
	return (u32)bi >> log2m;
}

// start row (as a byte index) of node `id`, or -1 when the node does not exist
__device__ __forceinline__ i32 node_start(u32 id, u32 S, i32 n, int log2m, const InvMeta* meta, i32 idx)
{
	if (id < S) {
		const u32 bi = (id << log2m) + mark_slot(id, log2m);
		return bi < (u32)n ? (i32)bi : -1;
	}
	const u32 k = id - S;
	if (k == 0) return -1;              // anchor 0 is the terminal (text position 0): nothing lies to its left
	return row_to_byte(meta->anchor_row[k], idx);
}

__device__ __forceinline__ u64 pack_rec(u32 next, u32 dist) { return ((u64)next << 32) | dist; }

// ---- work distribution ------------------------------------------------------------------------------
// Sub-chain lengths are geometric, so lanes finish at different times. A lane that is out of work takes the
// next node id from its warp's private range; only when the range runs dry does the warp touch the global
// ticket counter (one atomic per WALK_BATCH sub-chains instead of one per finished sub-chain -- the round trip
// of that atomic stalls the whole warp).
constexpr u32 WALK_BATCH = 128;
struct WarpTickets { u32 next, end; };

__device__ __forceinline__ u32 take_ticket(WarpTickets& wt, bool need, u32* __restrict__ ticket, u32 nodes, bool& done)
{
	const u32 nm = __ballot_sync(0xffffffffu, need);
	if (nm == 0) return REC_INVALID;
	const u32 cnt = __popc(nm), r = __popc(nm & lanemask_lt()), avail = wt.end - wt.next;
	u32 my = REC_INVALID;
	if (need && r < avail) my = wt.next + r;
	if (cnt > avail) {
		u32 base = 0;
		if (lane_id() == 0) base = atomicAdd(ticket, WALK_BATCH);
		base = __shfl_sync(0xffffffffu, base, 0);
		if (need && r >= avail) my = base + (r - avail);
		wt.next = base + (cnt - avail);
		wt.end = base + WALK_BATCH;
	} else wt.next += cnt;
	if (need && my >= nodes) { done = true; my = REC_INVALID; }
	return my;
}

// ---- L2 policies ---------------------------------------------------------------------------------------
// LF gathers have no reuse (11 % L2 hit rate) and flow through L2 at ~3 TB/s; the output block is written
// once, in pieces, over the whole launch. Marking the gathers evict-first and the stores evict-last lets a
// 64 MiB output sit in the 126 MB L2 until its sectors are complete instead of being evicted half-written.
__device__ __forceinline__ u64 policy_evict_first() { u64 p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ u64 policy_evict_last()  { u64 p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ u32 ld_lf(const u32* p, u64 pol, bool hinted)
{
	u32 v;
	if (hinted) asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
	else v = __ldg(p);
	return v;
}
__device__ __forceinline__ void st_out16(u8* p, u32 a0, u32 a1, u32 a2, u32 a3, u64 pol, bool hinted)
{
	if (hinted) asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;" :: "l"(p), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "l"(pol) : "memory");
	else *reinterpret_cast<uint4*>(p) = make_uint4(a0, a1, a2, a3);
}
__device__ __forceinline__ void st_out4(u8* p, u32 v, u64 pol, bool hinted)
{
	if (hinted) asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" :: "l"(p), "r"(v), "l"(pol) : "memory");
	else *reinterpret_cast<u32*>(p) = v;
}
__device__ __forceinline__ void st_out1(u8* p, u32 v, u64 pol, bool hinted)
{
	if (hinted) asm volatile("st.global.L2::cache_hint.u8 [%0], %1, %2;" :: "l"(p), "r"(v), "l"(pol) : "memory");
	else *p = (u8)v;
}
constexpr int WF_LOAD_EVICT_FIRST = 1, WF_STORE_EVICT_LAST = 2;

// Pass 1: per sub-chain (next node, length).
__global__ void __launch_bounds__(INV_THREADS) k_inv_walk_len(const u32* __restrict__ lf, const InvMeta* __restrict__ meta,
                                                              i32 n, int log2m, u32 S, u64* __restrict__ rec,
                                                              u32* __restrict__ ticket, int flags)
{
	__shared__ AnchorTable anchors;
	load_anchor_table(anchors, meta);
	__syncthreads();
	const i32 idx = meta->idx;
	const u32 nodes = S + N_ANCHOR;
	const bool lh = (flags & WF_LOAD_EVICT_FIRST) != 0;
	const u64 pol_ld = policy_evict_first();
	WarpTickets wt = {0, 0};
	u32 id = REC_INVALID, v = 0, len = 0;
	bool done = false;
	for (;;) {
		const u32 my = take_ticket(wt, !done && id == REC_INVALID, ticket, nodes, done);
		if (my != REC_INVALID) {
			const i32 bi = node_start(my, S, n, log2m, meta, idx);
			bool live = bi >= 0;
			if (live) {
				v = ld_lf(lf + bi, pol_ld, lh);
				// a window's own mark that happens to sit on an anchor row would duplicate the anchor's sub-chain: drop it
				live = my >= S || (v & LF_ANCHOR) == 0;
			}
			if (!live) rec[my] = pack_rec(REC_INVALID, 0);
			else { id = my; len = 0; }
		}
		if (__ballot_sync(0xffffffffu, !done) == 0) break;
		if (id != REC_INVALID) {
			const i32 r = (i32)(v & LF_MASK);
			len++;
			if (r == idx) { rec[id] = pack_rec(S, len); id = REC_INVALID; }
			else {
				const i32 bi = row_to_byte(r, idx);
				v = ld_lf(lf + bi, pol_ld, lh);
				if (v & LF_MARK) { rec[id] = pack_rec(node_of(anchors, v, r, bi, log2m, S), len); id = REC_INVALID; }
			}
		}
	}
}

// Ranking: asynchronous pointer jumping. A record (next, dist) always means "dist bytes to the left of my
// start lies the start of `next`"; replacing it by (next.next, dist + next.dist) keeps that true whichever
// version of next's record was read, so no double buffering or rounds are needed: 8-byte loads/stores are
// single transactions. Anchors (ids >= S) absorb.
// (spread_bits, budget: the launch plan described at k_inv_rank_packed)
__global__ void __launch_bounds__(256) k_inv_rank(u64* __restrict__ rec, u32 S, i32 step, int* __restrict__ err, int check_units, int hop_cap,
                                                  int spread_bits, int budget)
{
	u32 id = blockIdx.x * blockDim.x + threadIdx.x;
	if (spread_bits) id = (((blockIdx.x * 0x9E3779B1u) & ((1u << spread_bits) - 1u)) << 8) | threadIdx.x;
	const u32 nodes = S + N_ANCHOR;
	if (id >= nodes) return;
	volatile u64* vrec = rec;
	u64 r = vrec[id];
	u32 nxt = (u32)(r >> 32), dist = (u32)r;
	if (nxt == REC_INVALID) return;
	int hops = 0;
	while (nxt < S) {
		if (budget && hops >= budget) { vrec[id] = pack_rec(nxt, dist); return; }
		const u64 o = vrec[nxt];
		nxt = (u32)(o >> 32);
		dist += (u32)o;
		if (nxt == REC_INVALID || ++hops > hop_cap) { dev_fail(err, DE_RANK_LOOP); nxt = S; dist = 0; break; }
		vrec[id] = pack_rec(nxt, dist);
	}
	vrec[id] = pack_rec(nxt, dist);
	if (check_units && id > S) {   // decode unit k-1 runs from anchor k down to anchor k-1 and is exactly `step` long
		if (nxt != id - 1 || dist != (u32)step) dev_fail(err, DE_CHAIN_LEN);
	}
}

// F-column symbol of a row: F[row] = c  <=>  C[c] <= row-1 < C[c+1]. A 1024-cell table over the top bits of row-1
// gives the last symbol starting at or before the cell; a short scan finishes the job (a cell holds parts of two
// symbols at most unless the symbols in it are rare, and rare symbols are rarely looked up): ~10 instructions against
// 32 for the 8-level binary search this replaces.
constexpr int SYM_CELLS = 1024;
struct SymTable { i32 C[257]; u8 cell[SYM_CELLS]; };
__device__ __forceinline__ int sym_shift(i32 n) { int sh = 0; while (((u32)n >> sh) >= (u32)SYM_CELLS) sh++; return sh; }
__device__ __forceinline__ void load_sym_table(SymTable& t, const InvMeta* __restrict__ meta, int sh)   // whole block; ends with a barrier
{
	for (int i = threadIdx.x; i < 257; i += blockDim.x) t.C[i] = meta->ctable[i];
	__syncthreads();
	for (int k = threadIdx.x; k < SYM_CELLS; k += blockDim.x) {
		const i32 x = (i32)((u32)k << sh);
		u32 lo = 0;
		#pragma unroll
		for (u32 st = 128; st > 0; st >>= 1) if (t.C[lo + st] <= x) lo += st;
		t.cell[k] = (u8)lo;
	}
	__syncthreads();
}
__device__ __forceinline__ u32 symbol_of_row(const SymTable& t, int sh, i32 row)
{
	const i32 x = row - 1;
	u32 c = t.cell[(u32)x >> sh];
	while (t.C[c + 1] <= x) c++;        // C[256] = nlen > x: stops at 255 at the latest
	return c;
}

// One aligned 16-byte window of the output, bytes [lo, hi) valid in (a0..a3), the rest zero. Whole words go
// out as stores, a word shared with a neighbouring sub-chain as a RED.OR into the zeroed output -- the two
// sub-chains own different bytes of it, so OR-ing is exact and needs no ordering. No loops, no byte stores.
__device__ __forceinline__ void flush_window(u8* __restrict__ win, u32 lo, u32 hi, u32 a0, u32 a1, u32 a2, u32 a3)
{
	if (lo == 0 && hi == 16) { *reinterpret_cast<uint4*>(win) = make_uint4(a0, a1, a2, a3); return; }
	const u64 h[2] = {((u64)a1 << 32) | a0, ((u64)a3 << 32) | a2};
	#pragma unroll
	for (u32 i = 0; i < 2; i++) {                       // 8-byte halves: at most two memory operations per window
		const u32 b0 = 8 * i, wlo = max(lo, b0), whi = min(hi, b0 + 8);
		if (whi > wlo) {
			unsigned long long* p = reinterpret_cast<unsigned long long*>(win + b0);
			if (whi - wlo == 8) *p = h[i];
			else atomicOr(p, (unsigned long long)h[i]);  // result unused: compiles to RED.OR.64
		}
	}
}

// Pass 2: the same walk, now emitting. A byte produced for text position q is OR-ed into its place in a
// 128-bit image of the aligned 16-byte window around q; the window is flushed when the walk leaves it or the
// sub-chain ends. The output block is zeroed beforehand (RED.OR merging of shared words).
__global__ void __launch_bounds__(INV_THREADS) k_inv_walk_emit(const u32* __restrict__ lf, const InvMeta* __restrict__ meta,
                                                               i32 n, i32 step, int log2m, u32 S, const u64* __restrict__ rec,
                                                               u32* __restrict__ ticket, u8* __restrict__ out,
                                                               int* __restrict__ err, int flags)
{
	__shared__ SymTable sym;
	const int sh = sym_shift(n);
	load_sym_table(sym, meta, sh);
	const i32 idx = meta->idx;
	const u32 nodes = S + N_ANCHOR;
	const bool lh = (flags & WF_LOAD_EVICT_FIRST) != 0;
	const u64 pol_ld = policy_evict_first();
	WarpTickets wt = {0, 0};
	u32 id = REC_INVALID, v = 0, a0 = 0, a1 = 0, a2 = 0, a3 = 0;
	i32 pos = 0, pos_end = 0;
	bool done = false;
	for (;;) {
		const u32 my = take_ticket(wt, !done && id == REC_INVALID, ticket, nodes, done);
		if (my != REC_INVALID) {
			const u64 r = rec[my];
			const u32 nxt = (u32)(r >> 32);
			const i32 bi = (nxt == REC_INVALID) ? -1 : node_start(my, S, n, log2m, meta, idx);
			if (bi >= 0) {
				const i64 pe = (i64)(nxt - S) * step + (u32)r;
				if (nxt < S || pe > n || pe <= 0) dev_fail(err, DE_CHAIN_RANGE);
				else { id = my; pos = pos_end = (i32)pe; a0 = a1 = a2 = a3 = 0; v = ld_lf(lf + bi, pol_ld, lh); }
			}
		}
		if (__ballot_sync(0xffffffffu, !done) == 0) break;
		if (id != REC_INVALID) {
			const i32 r = (i32)(v & LF_MASK);
			bool stop = (r == idx);
			if (!stop) {
				v = ld_lf(lf + row_to_byte(r, idx), pol_ld, lh);   // next gather goes out before the symbol search below
				stop = (v & LF_MARK) != 0;
			}
			const u32 c = symbol_of_row(sym, sh, r);
			if (pos <= 0) { dev_fail(err, DE_CHAIN_RANGE); id = REC_INVALID; }
			else {
				pos--;
				const u32 k = ((u32)pos >> 2) & 3u, bits = c << (((u32)pos & 3u) * 8);
				a0 |= (k == 0) ? bits : 0u; a1 |= (k == 1) ? bits : 0u; a2 |= (k == 2) ? bits : 0u; a3 |= (k == 3) ? bits : 0u;
				if (((u32)pos & 15u) == 0 || stop) {
					const i32 wbase = pos & ~15;
					flush_window(out + wbase, (u32)(pos - wbase), (u32)min(16, pos_end - wbase), a0, a1, a2, a3);
					a0 = a1 = a2 = a3 = 0;
				}
				if (stop) id = REC_INVALID;
			}
		}
	}
}

// ---- single-walk path ----------------------------------------------------------------------------------
// The two passes above gather every LF entry twice (once for the lengths, once for the bytes): 2 * nlen random
// sectors. For large blocks the walk is done ONCE: a walker decodes its sub-chain while it measures it and parks
// the bytes in a per-warp stream (row r of a warp's stream = the 32 bytes its lanes produced in loop iteration r,
// one coalesced 32 B store). After the ranking has told every sub-chain where it ends, k_inv_place REPLAYS each
// warp's loop -- same tickets, same lane assignment, no gathers -- reading the rows back in order and writing the
// bytes to their text positions with the same 16-byte window logic as k_inv_walk_emit.
//   * The replay is deterministic given (a) the ticket batches the warp drew (logged: batch_head/batch_next) and
//     (b) each sub-chain's length, which rides in the top bits of its record through the ranking.
//   * Streams are carved in 1 KiB chunks (32 rows) from a bump counter; chunks live in the OUTPUT block and in the
//     part of the consumed input block the records leave free, so the call still fits 6N: the placed text is
//     assembled in the (by then dead) LF table and copied to the output at the end.
//   * Tail rows (lanes idle while the warp's last sub-chains finish) are the overhead: ~25 % of nlen for 64 MiB.
//     If the chunk space or a length field overflows, DE_STREAM_OVERFLOW is raised and the host reruns the
//     two-pass path on the intact LF table.
#include "inv_stream.cuh"

// take_ticket, with the batch bases logged per warp (prev_batch is lane 0's)
__device__ __forceinline__ u32 take_ticket_log(WarpTickets& wt, bool need, u32* __restrict__ ticket, u32 nodes, bool& done,
                                               u32& prev_batch, u32 wgid, const StreamSpace& sp, int* __restrict__ err)
{
	const u32 nm = __ballot_sync(0xffffffffu, need);
	if (nm == 0) return REC_INVALID;
	const u32 cnt = __popc(nm), r = __popc(nm & lanemask_lt()), avail = wt.end - wt.next;
	u32 my = REC_INVALID;
	if (need && r < avail) my = wt.next + r;
	if (cnt > avail) {
		u32 base = 0;
		if (lane_id() == 0) {
			base = atomicAdd(ticket, WALK_BATCH);
			const u32 kb = base / WALK_BATCH;
			if (kb < sp.batch_cap) { if (prev_batch == ST_NONE) sp.batch_head[wgid] = kb; else sp.batch_next[prev_batch] = kb; }
			else dev_fail(err, DE_STREAM_OVERFLOW);
			prev_batch = kb;
		}
		base = __shfl_sync(0xffffffffu, base, 0);
		if (need && r >= avail) my = base + (r - avail);
		wt.next = base + (cnt - avail);
		wt.end = base + WALK_BATCH;
	} else wt.next += cnt;
	if (need && my >= nodes) { done = true; my = REC_INVALID; }
	return my;
}

__global__ void __launch_bounds__(INV_THREADS) k_inv_walk_stream(const u32* __restrict__ lf, const InvMeta* __restrict__ meta,
                                                                 i32 n, int log2m, u32 S, u64* __restrict__ rec,
                                                                 u32* __restrict__ ticket, u32* __restrict__ chunk_ctr,
                                                                 StreamSpace sp, int* __restrict__ err)
{
	__shared__ AnchorTable anchors;
	__shared__ SymTable sym;
	const int sh = sym_shift(n);
	load_anchor_table(anchors, meta);
	load_sym_table(sym, meta, sh);
	const i32 idx = meta->idx;
	const u32 nodes = S + N_ANCHOR;
	const u32 lane = lane_id();
	const u32 wgid = blockIdx.x * INV_WARPS + (threadIdx.x >> 5);
	const u32 cap = sp.cap0 + sp.cap1;
	const u64 pol_ld = policy_evict_first();
	WarpTickets wt = {0, 0};
	u32 id = REC_INVALID, v = 0, len = 0;
	u32 prev_batch = ST_NONE, chunk = ST_NONE, nextc = 0, row = ST_ROWS;
	u8* cp = nullptr;
	bool done = false;
	for (;;) {
		const u32 my = take_ticket_log(wt, !done && id == REC_INVALID, ticket, nodes, done, prev_batch, wgid, sp, err);
		if (my != REC_INVALID) {
			const i32 bi = node_start(my, S, n, log2m, meta, idx);
			bool live = bi >= 0;
			if (live) {
				v = ld_lf(lf + bi, pol_ld, false);
				live = my >= S || (v & LF_ANCHOR) == 0;             // (duplicate of an anchor's sub-chain, as in k_inv_walk_len)
			}
			if (!live) rec[my] = pack3(0, PR_NXT_INVALID, 0);
			else { id = my; len = 0; }
		}
		if (__ballot_sync(0xffffffffu, !done) == 0) break;
		if (row == ST_ROWS) {                                   // this iteration opens a new chunk (warp-uniform)
			if (chunk == ST_NONE && lane == 0) nextc = atomicAdd(chunk_ctr, 1u);
			const u32 c = __shfl_sync(0xffffffffu, nextc, 0);
			if (c >= cap) { dev_fail(err, DE_STREAM_OVERFLOW); break; }
			if (lane == 0) { if (chunk == ST_NONE) sp.chunk_head[wgid] = c; else sp.chunk_next[chunk] = c; }
			chunk = c; cp = chunk_ptr(sp, c); row = 0;
		}
		if (row == ST_ROWS - ST_AHEAD && lane == 0) nextc = atomicAdd(chunk_ctr, 1u);   // consumed ST_AHEAD iterations from now
		if (id != REC_INVALID) {
			const i32 r = (i32)(v & LF_MASK);
			len++;
			bool stop = (r == idx);
			u32 nxt = S;
			if (!stop) {
				const i32 bi = row_to_byte(r, idx);
				v = ld_lf(lf + bi, pol_ld, false);               // next gather goes out before the symbol search below
				if (v & LF_MARK) { stop = true; nxt = node_of(anchors, v, r, bi, log2m, S); }
			}
			cp[row * 32 + lane] = (u8)symbol_of_row(sym, sh, r);
			if (stop) {
				if (len > PR_LEN_MAX) { dev_fail(err, DE_STREAM_OVERFLOW); len = PR_LEN_MAX; }
				rec[id] = pack3(len, nxt, len);
				id = REC_INVALID;
			}
		}
		row++;
	}
}

// Pointer jumping over the packed records; the length field of a record is its own and is carried along.
// The jumping is asynchronous -- a thread profits from what the threads of its successors have already published -- and
// how well that works depends on the order the blocks start in. Index order suits lists that run through neighbouring
// records (single-symbol runs: each block finds the next one at work) and random lists (Markov text: 10.5 record reads per
// node); on real text, where the LF map sends neighbouring rows to neighbouring rows and whole bundles of lists run in
// parallel through records nobody has touched yet, it needs 0.83-0.90 ms instead of 0.30. Blocks taken in scattered order
// (`spread_bits`: block x works on the 256 nodes of block (x * odd) mod 2^spread_bits) do real text in 0.32 ms and a block
// of one repeated byte in 3.5 ms. So the ranking is two launches: scattered blocks with a budget of four hops per node --
// after which every record has jumped some nodes ahead, whatever the text -- then index order to the end.
// Measured (64 MiB, `tools/rank_order_ab.py`): Markov 0.33 (one launch: 0.30), source text 0.34 (0.90), all-`a` 0.29 (0.32),
// Markov with 0.2-50 % of zero runs 0.31-0.35 (0.30-0.33; scattered alone up to 1.2).
__global__ void __launch_bounds__(256) k_inv_rank_packed(u64* __restrict__ rec, u32 S, i32 step, int* __restrict__ err, int hop_cap, int spread_bits,
                                                         int budget /* hops this launch may take; 0 = to the end, with the final checks */)
{
	u32 id = blockIdx.x * blockDim.x + threadIdx.x;
	if (spread_bits) id = (((blockIdx.x * 0x9E3779B1u) & ((1u << spread_bits) - 1u)) << 8) | threadIdx.x;   // (blocks of 256 consecutive nodes, in scattered order)
	const u32 nodes = S + N_ANCHOR;
	if (id >= nodes || *(volatile int*)err != 0) return;
	volatile u64* vrec = rec;
	const u64 r = vrec[id];
	const u32 len = pr_len(r);
	u32 nxt = pr_nxt(r), dist = pr_dist(r);
	if (nxt == PR_NXT_INVALID) return;
	int hops = 0;
	while (nxt < S) {
		if (budget && hops >= budget) { vrec[id] = pack3(len, nxt, dist); return; }   // (a later launch goes on from here)
		const u64 o = vrec[nxt];
		const u32 onxt = pr_nxt(o);
		dist += pr_dist(o);
		if (onxt == PR_NXT_INVALID || ++hops > hop_cap) { dev_fail(err, DE_RANK_LOOP); nxt = S; dist = 0; break; }
		if (dist > PR_DIST_MASK) { dev_fail(err, DE_CHAIN_LEN); nxt = S; dist = 0; break; }   // no honest chain is longer than a unit
		nxt = onxt;
		vrec[id] = pack3(len, nxt, dist);
	}
	vrec[id] = pack3(len, nxt, dist);
	if (id > S && (nxt != id - 1 || dist != (u32)step)) dev_fail(err, DE_CHAIN_LEN);
}

// Zeroes the area the text is assembled in -- the head of the LF table -- unless the walk failed: the two-pass rerun
// needs the table intact (in consume mode the BWT itself is gone by then).
__global__ void __launch_bounds__(256) k_inv_clear_text(uint4* __restrict__ text, u32 n16, const int* __restrict__ err)
{
	if (*(volatile const int*)err != 0) return;
	for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) text[i] = make_uint4(0, 0, 0, 0);
}

// Replay of k_inv_walk_stream's loop: iteration i of warp w reads row i of w's stream.
// What the kernel is short of is issue slots, not bandwidth (round 1: 125 warp instructions per 32-byte row, two or three
// lanes leaving the common path in nearly every row), so everything per-row is kept to a few instructions:
//   * records are decoded and range-checked when a ticket batch is staged in shared memory (all lanes, 4 records each),
//     so a lane that starts a sub-chain reads back (end position, length) and nothing else;
//   * the exit conditions (all tickets drawn, a failed check) can only change where tickets are drawn, and are voted on there;
//   * the 16-byte window image is a 128-bit shift register: text positions fall by one per step, so the new byte enters
//     at the bottom and a window that was filled from its top needs no alignment at all; a window left part-way is
//     shifted into place once, on the way out;
//   * a full window leaves as one 16-byte store, anything else as RED.OR.64 of its non-zero halves into the zeroed text
//     (bytes a sub-chain does not own are zero in its image, so no byte ranges are worked out).
__global__ void __launch_bounds__(INV_THREADS) k_inv_place(i32 n, i32 step, u32 S, const u64* __restrict__ rec,
                                                           StreamSpace sp, u8* __restrict__ out, int* __restrict__ err)
{
	__shared__ uint4 sdata[INV_WARPS][ST_CHUNK / 16];
	__shared__ uint2 srec[INV_WARPS][2][WALK_BATCH];             // decoded: .x end position, .y length (0 = nothing to place, bit 31 = failed a check)
	if (*(volatile int*)err != 0) return;                       // the stream is incomplete: the host falls back
	const u32 nodes = S + N_ANCHOR;
	const u32 lane = lane_id(), w = threadIdx.x >> 5;
	const u32 wgid = blockIdx.x * INV_WARPS + w;
	const u32 cap = sp.cap0 + sp.cap1;
	const u8* sbytes = reinterpret_cast<const u8*>(sdata[w]) + lane;
	u32 next_t = 0, end_t = 0, base_t = 0, par = 0;
	u32 prev_batch = ST_NONE, chunk = ST_NONE, row = ST_ROWS;
	u32 left = 0, a0 = 0, a1 = 0, a2 = 0, a3 = 0;                // left > 0: a sub-chain is being placed
	i32 pos = 0, pos_end = 0;
	bool done = false;
	// (a replay that has diverged from the walk -- only possible after a failed check -- must not chase stale log
	// entries: any failure ends the warp, and the iteration count is bounded by the rows that exist)
	for (u32 iter = 0; iter < cap * ST_ROWS + 2; iter++) {
		// ---- the ticket logic of take_ticket_log, batches read back from the log, records from shared memory
		const bool need = !done && left == 0;
		const u32 nm = __ballot_sync(0xffffffffu, need);
		if (nm != 0) {
			const u32 cnt = __popc(nm), r = __popc(nm & lanemask_lt()), avail = end_t - next_t;
			u32 my = REC_INVALID; uint2 rv = make_uint2(0, 0);
			if (need && r < avail) { my = next_t + r; rv = srec[w][par][my - base_t]; }
			if (cnt > avail) {
				u32 kb = 0;
				if (lane == 0) { kb = (prev_batch == ST_NONE) ? sp.batch_head[wgid] : sp.batch_next[prev_batch]; prev_batch = kb; }
				kb = __shfl_sync(0xffffffffu, kb, 0);
				if (kb >= sp.batch_cap) { dev_fail(err, DE_STREAM_OVERFLOW); break; }
				const u32 base = kb * WALK_BATCH;
				par ^= 1;
				__syncwarp();
				#pragma unroll
				for (int k = 0; k < (int)WALK_BATCH / 32; k++) {
					const u32 q = base + k * 32 + lane;
					uint2 dv = make_uint2(0, 0);
					if (q < nodes) {
						const u64 rr = rec[q];
						const u32 nxt = pr_nxt(rr), L = pr_len(rr);
						if (nxt != PR_NXT_INVALID) {
							const i64 pe = (i64)(nxt - S) * step + pr_dist(rr);
							const bool ok = nxt >= S && nxt < nodes && L != 0 && pe <= n && pe >= (i64)L;
							dv = ok ? make_uint2((u32)pe, L) : make_uint2(0, 0x80000000u);
						}
					}
					srec[w][par][k * 32 + lane] = dv;
				}
				__syncwarp();
				if (need && r >= avail) { my = base + (r - avail); rv = srec[w][par][my - base]; }
				next_t = base + (cnt - avail); end_t = base + WALK_BATCH; base_t = base;
			} else next_t += cnt;
			if (need && my >= nodes) { done = true; rv.y = 0; }
			if (rv.y & 0x80000000u) { dev_fail(err, DE_CHAIN_RANGE); done = true; }      // (the call fails; this lane stops, the others stay inside their checked ranges)
			else if (rv.y != 0) { left = rv.y; pos = pos_end = (i32)rv.x; a0 = a1 = a2 = a3 = 0; }
			if (__ballot_sync(0xffffffffu, !done) == 0) break;       // (the only place it can change)
		}
		if (row == ST_ROWS) {
			u32 c = 0;
			if (lane == 0) c = (chunk == ST_NONE) ? sp.chunk_head[wgid] : sp.chunk_next[chunk];
			c = __shfl_sync(0xffffffffu, c, 0);
			if (c >= cap) { dev_fail(err, DE_STREAM_OVERFLOW); break; }
			const uint4* src = reinterpret_cast<const uint4*>(chunk_ptr(sp, c));
			__syncwarp();
			sdata[w][lane] = __ldcs(src + lane);
			sdata[w][32 + lane] = __ldcs(src + 32 + lane);
			__syncwarp();
			chunk = c; row = 0;
		}
		if (left != 0) {
			const u32 c = sbytes[row * 32];
			pos--; left--;
			a3 = (a3 << 8) | (a2 >> 24); a2 = (a2 << 8) | (a1 >> 24); a1 = (a1 << 8) | (a0 >> 24); a0 = (a0 << 8) | c;
			const u32 sh = (u32)pos & 15u;
			if (sh == 0 || left == 0) {
				u8* win = out + (pos & ~15);
				if (sh == 0 && pos_end - pos >= 16) *reinterpret_cast<uint4*>(win) = make_uint4(a0, a1, a2, a3);   // the whole window is this sub-chain's
				else {
					// left part-way (or entered part-way): the newest byte belongs at offset sh -- whole words by selects, the
					// rest by funnel shifts, no branches (the lanes that get here differ in sh)
					const u32 q = sh >> 2, rb = (sh & 3u) * 8;
					const u32 w0 = q == 0 ? a0 : 0u;
					const u32 w1 = q == 0 ? a1 : q == 1 ? a0 : 0u;
					const u32 w2 = q == 0 ? a2 : q == 1 ? a1 : q == 2 ? a0 : 0u;
					const u32 w3 = q == 0 ? a3 : q == 1 ? a2 : q == 2 ? a1 : a0;
					const u64 lo = ((u64)__funnelshift_l(w0, w1, rb) << 32) | (u64)(w0 << rb);
					const u64 hi = ((u64)__funnelshift_l(w2, w3, rb) << 32) | (u64)__funnelshift_l(w1, w2, rb);
					if (lo) atomicOr(reinterpret_cast<unsigned long long*>(win), (unsigned long long)lo);       // result unused: RED.OR.64
					if (hi) atomicOr(reinterpret_cast<unsigned long long*>(win) + 1, (unsigned long long)hi);
				}
				a0 = a1 = a2 = a3 = 0;
			}
		}
		row++;
	}
}

// ---- host driver -----------------------------------------------------------------------------------
// Marker spacing m. A pass costs about nlen / (gather rate) + (longest sub-chain) * (unloaded DRAM latency), and
// the longest of nlen/m geometric sub-chains is ~ m * ln(nlen/m): measured on 64 MiB, m = 64 / 32 / 16 give
// 1.46 / 1.29 / 1.19 ms for pass 1. Smaller m means more 8-byte records (8 * nlen / m bytes) and more ranking work.
static int pick_log2m(i32 nlen)
{
	if (const char* e = getenv("JP_BWT_INV_LOG2M")) { int v = atoi(e); if (v >= 2 && v <= 12) return v; }
	return nlen >= (1 << 22) ? 4 : 3;
}

static int walk_flags()
{
	if (const char* e = getenv("JP_BWT_INV_FLAGS")) return atoi(e);
	return 0;   // measured on B200: the L2 policies change nothing (3.73 ms without, 3.82-3.91 ms with)
}

// Launch plan of the ranking kernels. JP_BWT_INV_RANK_PLAN: launches as "<order><hops>" separated by commas, order i
// (index) or s (scattered blocks), hops 0 = unbounded; the last launch is always unbounded (it carries the checks).
template <typename Launch>
static int rank_plan(Ctx& c, u32 nodes, Launch launch)
{
	int spread = bit_length(((u64)nodes + 255) / 256 - 1);
	if (spread < 1 || spread > 23) spread = 0;
	const char* plan = getenv("JP_BWT_INV_RANK_PLAN") ? getenv("JP_BWT_INV_RANK_PLAN") : "s4,i0";
	bool complete = false;
	for (const char* q = plan; *q && !complete;) {
		const bool scattered = *q == 's' && spread != 0;
		const char* comma = strchr(q, ',');
		const int hops = (comma && comma[1]) ? std::max(0, atoi(q + 1)) : 0;
		launch(scattered ? (1u << spread) : (nodes + 255) / 256, scattered ? spread : 0, hops);
		JP_LAUNCH(c);
		complete = hops == 0;
		q = comma ? comma + 1 : q + strlen(q);
	}
	if (!complete) { launch((nodes + 255) / 256, 0, 0); JP_LAUNCH(c); }      // (whatever the plan said: the ranking ends with an unbounded launch)
	return JP_OK;
}

static int walker_blocks(Ctx& c, const void* kernel)
{
	int per_sm = 0;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, INV_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
	return c.sm_count * per_sm;
}

struct InvBuffers {
	u32* lf; u32* tile_hist; u32* bin_total; InvMeta* meta; u64* rec; u32* ticket; int* err;
	int tiles; int log2m; u32 S;
	bool single;                  // single-walk path (k_inv_walk_stream / k_inv_place)
	int  wblocks;                 // walker blocks of the single-walk kernels (the replay needs the same grid)
	StreamSpace sp;
};

// Single-walk path: on by default for blocks of 30 Mi and more, where the stream overhead (tail rows, ~76 per
// walker warp) fits beside the records in the consumed input block. JP_BWT_INV_SINGLE=0 turns it off, =1 forces it
// for every block of 64 Ki and more with the stream space topped up from the workspace (tests).
constexpr i32 SINGLE_WALK_MIN = 30 << 20;   // (a 32 MiB block has nlen = 2^25 - 32)
static int single_mode()
{
	if (const char* e = getenv("JP_BWT_INV_SINGLE")) return atoi(e) != 0 ? 1 : 0;
	return -1;
}

// `scratch_in`: when the caller's input block may be overwritten once the LF table is built, the sub-chain
// records live there (8 * nodes <= nlen bytes) and the workspace is lf + per-tile histograms: 6N + N/64 in all.
static int inv_alloc(Ctx& c, i32 nlen, InvBuffers& b, u8* scratch_in = nullptr, u8* d_out = nullptr)
{
	b.tiles = (int)(((i64)nlen + INV_TILE - 1) / INV_TILE);
	b.log2m = pick_log2m(nlen);
	b.S = (u32)(((i64)nlen + (1 << b.log2m) - 1) >> b.log2m);
	const size_t nodes = (size_t)b.S + N_ANCHOR;
	const bool rec_in_input = scratch_in != nullptr && nodes * 8 <= (size_t)nlen && ((uintptr_t)scratch_in & 7) == 0;

	// stream space of the single-walk path
	const int mode = single_mode();
	b.single = d_out != nullptr && ((uintptr_t)d_out & 15) == 0 && nodes < PR_NXT_INVALID && (u32)(nlen / JP_BWT_UNITS) <= PR_DIST_MASK &&
	           (mode == 1 ? nlen >= (1 << 16) : (mode == -1 && nlen >= SINGLE_WALK_MIN));
	size_t extra_stream = 0, in_free_off = 0, in_free = 0;
	b.wblocks = 0;
	b.sp = StreamSpace{};
	const size_t chunk_bytes = (size_t)ST_CHUNK;
	const int wthreads = INV_THREADS;
	if (b.single) {
		// every lane should see a handful of sub-chains, or the tail rows dominate the stream
		// (measured on 64 MiB: 8 / 6 / 5 / 4 / 3 blocks per SM walk in 1.154 / 1.142 / 1.146 / 1.173 / 1.388 ms and leave
		// streams of 1.32 / 1.24 / 1.20 / 1.16 / 1.12 nlen: the gathers saturate DRAM from 4 blocks per SM on, and every
		// walker warp adds ~76 tail rows)
		// Under 47 Mi one block per SM less: ~2.5 MB less of tail rows, which is what still fits beside the records.
		int resident = walker_blocks(c, (const void*)k_inv_walk_stream), per_sm = nlen >= (47 << 20) ? 5 : 4;
		if (const char* e = getenv("JP_BWT_INV_WBLOCKS_PER_SM")) { const int v = atoi(e); if (v >= 1) per_sm = v; }
		resident = std::min(resident, per_sm * c.sm_count);
		b.wblocks = (int)std::max<size_t>(1, std::min<size_t>((size_t)resident, nodes / (INV_THREADS * 8)));
		if (rec_in_input) {
			in_free_off = (nodes * 8 + 15) & ~(size_t)15;
			in_free = (size_t)nlen > in_free_off ? (size_t)nlen - in_free_off : 0;
		} else extra_stream = (size_t)nlen / 2;
		if (mode == 1) extra_stream += (size_t)nlen / 2 + (size_t)b.wblocks * INV_WARPS * 160 * 32;
		b.sp.cap0 = (u32)((size_t)nlen / ST_CHUNK);
		b.sp.batch_cap = (u32)(nodes / WALK_BATCH + (size_t)b.wblocks * INV_WARPS * 4 + 16);
	}
	const size_t walker_warps = (size_t)b.wblocks * (wthreads / 32);
	if (const char* e = getenv("JP_BWT_INV_STREAM_CAP")) {          // tests: shrink the stream space to provoke the two-pass rerun
		const long v = atol(e);
		if (b.single && v >= 0 && (u32)v < b.sp.cap0) { b.sp.cap0 = (u32)v; in_free = 0; extra_stream = 0; }
	}
	const u32 cap_in = (u32)(in_free / chunk_bytes), cap_extra = (u32)(extra_stream / chunk_bytes);
	if (b.single) {
		// region 1 is ONE address range: the free part of the input when there is no top-up, else the workspace piece
		// alone (the input's free part is then left unused -- only tests and the non-consuming entry point get here)
		b.sp.cap1 = cap_extra > 0 ? cap_extra : cap_in;
	}
	size_t total = Arena::align((size_t)nlen * 4) + Arena::align((size_t)b.tiles * 256 * 4) + Arena::align(256 * 4) +
	               Arena::align(sizeof(InvMeta)) + (rec_in_input ? 0 : Arena::align(nodes * 8)) + Arena::align(64) + Arena::align(64);
	if (b.single) total += Arena::align((size_t)cap_extra * chunk_bytes) + 2 * Arena::align(walker_warps * 4) +
	                       Arena::align(((size_t)b.sp.cap0 + b.sp.cap1) * 4) + Arena::align((size_t)b.sp.batch_cap * 4);
	JP_TRY(arena_reserve(c, total));
	b.lf = arena_take<u32>(c, (size_t)nlen);
	b.tile_hist = arena_take<u32>(c, (size_t)b.tiles * 256);
	b.bin_total = arena_take<u32>(c, 256);
	b.meta = arena_take<InvMeta>(c, 1);
	b.rec = rec_in_input ? reinterpret_cast<u64*>(scratch_in) : arena_take<u64>(c, nodes);
	b.ticket = arena_take<u32>(c, 16);
	b.err = arena_take<int>(c, 16);
	if (b.single) {
		b.sp.base0 = d_out;
		b.sp.base1 = cap_extra > 0 ? arena_take<u8>(c, (size_t)cap_extra * chunk_bytes) : scratch_in + in_free_off;
		b.sp.chunk_head = arena_take<u32>(c, walker_warps);
		b.sp.batch_head = arena_take<u32>(c, walker_warps);
		b.sp.chunk_next = arena_take<u32>(c, (size_t)b.sp.cap0 + b.sp.cap1);
		b.sp.batch_next = arena_take<u32>(c, b.sp.batch_cap);
	}
	return JP_OK;
}

static int inv_build_table(Ctx& c, const u8* d_in, i32 len, i32 nlen, u8* d_out, InvBuffers& b, cudaStream_t s)
{
	JP_CUDA(cudaMemsetAsync(b.ticket, 0, 64, s));
	JP_CUDA(cudaMemsetAsync(b.err, 0, 64, s));
	k_inv_prepare<<<1, 128, 0, s>>>(d_in, len, nlen, d_out, b.meta, b.err); JP_LAUNCH(c);
	k_inv_hist<<<b.tiles, INV_THREADS, 0, s>>>(d_in, nlen, b.tile_hist); JP_LAUNCH(c);
	k_inv_scan_tiles<<<256, 256, 0, s>>>(b.tile_hist, b.tiles, b.bin_total); JP_LAUNCH(c);
	k_inv_ctable<<<1, 256, 0, s>>>(b.bin_total, b.meta, nlen); JP_LAUNCH(c);
	JP_KCHECK();
	return JP_OK;
}

int inverse_device(Ctx& c, const u8* d_in, i32 len_with_trailer, u8* d_out, cudaStream_t s, jp_bwt_stats* st, u8* scratch_in)
{
	const i32 len = len_with_trailer - JP_BWT_TRAILER_BYTES;            // bwt.cpp:77
	if (len < 0) return JP_ERR_ARG;
	const i32 nlen = len - len % JP_BWT_UNITS;                          // bwt.cpp:80-81
	st->direction = 1; st->len = len; st->nlen = nlen; st->device = c.device;
	JP_CUDA(cudaEventRecord(c.ev[0], s));
	if (nlen == 0) {                                                    // bwt.cpp:85: nothing but the raw tail
		if (len > 0) JP_CUDA(cudaMemcpyAsync(d_out, d_in, (size_t)len, cudaMemcpyDeviceToDevice, s));
		JP_CUDA(cudaEventRecord(c.ev[1], s));
		JP_CUDA(cudaStreamSynchronize(s));
		JP_CUDA(cudaEventElapsedTime(&st->ms_total, c.ev[0], c.ev[1]));
		return JP_OK;
	}
	const i32 step = nlen / JP_BWT_UNITS;                               // bwt.cpp:176 with N_Units = 120
	InvBuffers b;
	JP_TRY(inv_alloc(c, nlen, b, scratch_in, d_out));
	JP_TRY(inv_build_table(c, d_in, len, nlen, d_out, b, s));
	JP_CUDA(cudaEventRecord(c.ev[1], s));
	k_inv_lf<<<b.tiles, INV_THREADS, 0, s>>>(d_in, nlen, b.tile_hist, b.meta, b.lf, b.log2m);
	JP_LAUNCH(c);
	k_inv_mark_anchors<<<1, 128, 0, s>>>(b.meta, b.lf, nlen); JP_LAUNCH(c);
	JP_KCHECK();
	JP_CUDA(cudaEventRecord(c.ev[2], s));
	const u32 nodes = b.S + N_ANCHOR;
	// a decode unit holds about nodes/120 sub-chains, so no honest list is longer than a few times that; the cap
	// only bounds the time spent on a corrupt block whose records form cycles
	const int hop_cap = (int)std::min<u64>((u64)RANK_HOP_CAP, (u64)nodes / 16 + 4096);
	st->stream_chunks = 0;
	bool two_pass = !b.single;
	if (b.single) {
		u8* text = reinterpret_cast<u8*>(b.lf);                             // the LF table is dead once the walk is over
		// (serialising the walks of the blocks in flight on one GPU behind a host-side gate, so that the other blocks'
		// table builds, rankings and placements could fill in beside a single DRAM-bound walk, was measured: 35.3-35.6
		// against 35.7-36.0 GB/s with four blocks in flight -- the hardware's own interleaving is as good; the same chain
		// built from event dependencies between the callers' streams, no host thread blocking: 37.3 against 36.4.
		// tools/inflight_phases.py shows why neither helps: with four blocks in flight every kernel of a call stretches,
		// the walk 1.9x, the others 3-9x -- the kernels take turns on the SMs rather than overlap)
		k_inv_walk_stream<<<b.wblocks, INV_THREADS, 0, s>>>(b.lf, b.meta, nlen, b.log2m, b.S, b.rec, b.ticket, b.ticket + 2, b.sp, b.err);
		JP_LAUNCH(c);
		JP_KCHECK();
		JP_CUDA(cudaEventRecord(c.ev[3], s));
		JP_TRY(rank_plan(c, nodes, [&](u32 blocks, int spread, int budget) {
			k_inv_rank_packed<<<blocks, 256, 0, s>>>(b.rec, b.S, step, b.err, hop_cap, spread, budget);
		}));
		JP_KCHECK();
		JP_CUDA(cudaEventRecord(c.ev[4], s));
		k_inv_clear_text<<<c.sm_count * 8, 256, 0, s>>>(reinterpret_cast<uint4*>(text), (u32)(((size_t)nlen + 15) / 16), b.err); JP_LAUNCH(c);
		k_inv_place<<<b.wblocks, INV_THREADS, 0, s>>>(nlen, step, b.S, b.rec, b.sp, text, b.err);
		JP_LAUNCH(c);
		JP_KCHECK();
		JP_CUDA(cudaMemcpyAsync(d_out, text, (size_t)nlen, cudaMemcpyDeviceToDevice, s));   // (meaningless but harmless after a failure)
		JP_CUDA(cudaEventRecord(c.ev[5], s));
		JP_CUDA(cudaMemcpyAsync(c.h_small, b.err, sizeof(int), cudaMemcpyDeviceToHost, s));
		JP_CUDA(cudaMemcpyAsync(c.h_small + 4, b.ticket + 2, sizeof(u32), cudaMemcpyDeviceToHost, s));
		JP_CUDA(cudaStreamSynchronize(s));
		st->stream_chunks = (i32)std::min<u64>((u64)(u32)c.h_small[4], 0x7fffffffull);   // 1 KiB chunks
		st->random_sectors = (u64)nlen;
		if (c.h_small[0] == DE_STREAM_OVERFLOW) {                           // rare: rerun on the intact LF table
			two_pass = true;
			st->stream_chunks = -st->stream_chunks;
			JP_CUDA(cudaMemsetAsync(b.ticket, 0, 64, s));
			JP_CUDA(cudaMemsetAsync(b.err, 0, 64, s));
			JP_CUDA(cudaEventRecord(c.ev[2], s));
		}
	}
	if (two_pass) {
		const int wb1 = walker_blocks(c, (const void*)k_inv_walk_len);
		k_inv_walk_len<<<wb1, INV_THREADS, 0, s>>>(b.lf, b.meta, nlen, b.log2m, b.S, b.rec, b.ticket, walk_flags()); JP_LAUNCH(c);
		JP_KCHECK();
		JP_CUDA(cudaEventRecord(c.ev[3], s));
		// (a two-level scheme -- majors walk to majors, jump, hand down -- was tried: 0.86 ms against 0.28 ms; the
		// dependent record-to-record walks are latency-bound, the flat jumping is not)
		JP_TRY(rank_plan(c, nodes, [&](u32 blocks, int spread, int budget) {
			k_inv_rank<<<blocks, 256, 0, s>>>(b.rec, b.S, step, b.err, 1, hop_cap, spread, budget);
		}));
		JP_KCHECK();
		JP_CUDA(cudaEventRecord(c.ev[4], s));
		JP_CUDA(cudaMemsetAsync(d_out, 0, (size_t)nlen, s));               // shared words of neighbouring sub-chains are merged by RED.OR
		const int wb2 = walker_blocks(c, (const void*)k_inv_walk_emit);
		k_inv_walk_emit<<<wb2, INV_THREADS, 0, s>>>(b.lf, b.meta, nlen, step, b.log2m, b.S, b.rec, b.ticket + 1, d_out, b.err, walk_flags()); JP_LAUNCH(c);
		JP_KCHECK();
		JP_CUDA(cudaEventRecord(c.ev[5], s));
		JP_CUDA(cudaMemcpyAsync(c.h_small, b.err, sizeof(int), cudaMemcpyDeviceToHost, s));
		JP_CUDA(cudaStreamSynchronize(s));
		st->random_sectors = 2ull * (u64)nlen;
	}
	for (int i = 0; i < 5; i++) JP_CUDA(cudaEventElapsedTime(&st->ms_phase[i], c.ev[i], c.ev[i + 1]));
	JP_CUDA(cudaEventElapsedTime(&st->ms_total, c.ev[0], c.ev[5]));
	st->subchains = (i32)nodes;
	st->subchain_spacing = 1 << b.log2m;
	st->device_bytes = c.arena.high;
	return map_dev_err(c.h_small[0]);
}

// ---- test hook: the LF table itself -----------------------------------------------------------------
__global__ void k_strip_marks(u32* lf, i32 n)
{
	const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) lf[i] &= LF_MASK;
}

int debug_lf(Ctx& c, const u8* h_in, i32 nlen, i32* h_lf, i32* h_ctable)
{
	if (nlen <= 0) return JP_ERR_ARG;
	cudaStream_t s = c.own_stream;
	InvBuffers b;
	c.arena.reset();
	const size_t extra = Arena::align((size_t)nlen + JP_BWT_TRAILER_BYTES + 16) + Arena::align((size_t)nlen + 16);
	JP_TRY(arena_reserve(c, extra + (size_t)nlen * 6 + (4 << 20)));
	u8* d_in = arena_take<u8>(c, (size_t)nlen + JP_BWT_TRAILER_BYTES + 16);
	u8* d_out = arena_take<u8>(c, (size_t)nlen + 16);
	JP_TRY(inv_alloc(c, nlen, b));
	JP_CUDA(cudaMemsetAsync(d_in, 0, (size_t)nlen + JP_BWT_TRAILER_BYTES, s));
	JP_CUDA(cudaMemcpyAsync(d_in, h_in, (size_t)nlen, cudaMemcpyHostToDevice, s));
	// a syntactically valid trailer (all distinct, in range) so prepare does not flag it
	{
		i32 fake[JP_BWT_UNITS];
		for (int k = 0; k < JP_BWT_UNITS; k++) fake[k] = 1 + (i32)(((i64)k * nlen) / JP_BWT_UNITS);
		JP_CUDA(cudaMemcpyAsync(d_in + nlen, fake, sizeof(fake), cudaMemcpyHostToDevice, s));
		JP_CUDA(cudaStreamSynchronize(s));
	}
	JP_TRY(inv_build_table(c, d_in, nlen, nlen, d_out, b, s));
	k_inv_lf<<<b.tiles, INV_THREADS, 0, s>>>(d_in, nlen, b.tile_hist, b.meta, b.lf, b.log2m); JP_LAUNCH(c);
	k_strip_marks<<<(nlen + 255) / 256, 256, 0, s>>>(b.lf, nlen); JP_LAUNCH(c);
	JP_KCHECK();
	JP_CUDA(cudaMemcpyAsync(h_lf, b.lf, (size_t)nlen * 4, cudaMemcpyDeviceToHost, s));
	InvMeta hm;
	JP_CUDA(cudaMemcpyAsync(&hm, b.meta, sizeof(InvMeta), cudaMemcpyDeviceToHost, s));
	JP_CUDA(cudaStreamSynchronize(s));
	for (int i = 0; i < 257; i++) h_ctable[i] = hm.ctable[i];
	return JP_OK;
}

// ---- micro-benchmark: random 4-byte gathers (the roofline denominator) -----------------
__global__ void k_fill_perm(u32* tab, u32 n_words)
{
	// successor table = the bijection x -> (a*x + c) mod 2^k (a odd): walkers never merge, so no step is
	// served from a line another walker just pulled in. n_words is a power of two.
	const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n_words) tab[i] = ((u32)i * 0x9E3779B1u + 0x7F4A7C15u) & (n_words - 1);
}
__global__ void __launch_bounds__(256) k_gather_dep(const u32* __restrict__ tab, u32 n_words, int steps, u32* __restrict__ sink)
{
	const u32 gid = blockIdx.x * blockDim.x + threadIdx.x;
	u32 p = (u32)(((u64)mix32(gid + 1) * n_words) >> 32);
	for (int i = 0; i < steps; i++) p = tab[p];
	if (p == 0xffffffffu) sink[0] = p;
}
__global__ void __launch_bounds__(256) k_gather_indep(const u32* __restrict__ tab, u32 n_words, int steps, u32* __restrict__ sink)
{
	const u32 gid = blockIdx.x * blockDim.x + threadIdx.x;
	u32 acc = 0, x = gid * 0x9E3779B1u + 7u;
	#pragma unroll 4
	for (int i = 0; i < steps; i++) { x = x * 1664525u + 1013904223u; acc += tab[(u32)(((u64)mix32(x) * n_words) >> 32)]; }
	if (acc == 0x12345678u) sink[0] = acc;
}

double debug_gather_rate(Ctx& c, u64 table_bytes, i32 chains, i32 steps, int dependent)
{
	cudaStream_t s = c.own_stream;
	if (table_bytes < 1024 || table_bytes > (8ull << 30) || chains <= 0 || steps <= 0) return JP_ERR_ARG;
	u32 n_words = 1;
	while ((u64)n_words * 8 <= table_bytes) n_words <<= 1;   // largest power of two with 4*n_words <= table_bytes
	c.arena.reset();
	if (arena_reserve(c, Arena::align((size_t)n_words * 4) + 4096) != JP_OK) return JP_ERR_OOM;
	u32* tab = arena_take<u32>(c, n_words);
	u32* sink = arena_take<u32>(c, 16);
	k_fill_perm<<<(n_words + 255) / 256, 256, 0, s>>>(tab, n_words);
	const int blocks = (chains + 255) / 256;
	float best = 1e30f;
	for (int rep = 0; rep < 4; rep++) {
		cudaEventRecord(c.ev[0], s);
		if (dependent) k_gather_dep<<<blocks, 256, 0, s>>>(tab, n_words, steps, sink);
		else k_gather_indep<<<blocks, 256, 0, s>>>(tab, n_words, steps, sink);
		cudaEventRecord(c.ev[1], s);
		if (cudaStreamSynchronize(s) != cudaSuccess) return JP_ERR_CUDA;
		float ms = 0; cudaEventElapsedTime(&ms, c.ev[0], c.ev[1]);
		if (rep > 0 && ms < best) best = ms;
	}
	return (double)blocks * 256.0 * (double)steps / ((double)best * 1e-3);
}

} // namespace jp
