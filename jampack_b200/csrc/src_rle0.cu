// src_rle0.cu -- first half of Jampack's second stage on the device: sorted rank coding and
// RLE0 of a BWT block that is still resident in HBM, chunk by chunk exactly like Ans::Encode (reference ans.cpp:134-160:
// StackSize = 1 MiB chunks, everything reset per chunk). The adaptive rANS that follows (ans.cpp:162-221) stays on
// the host for now: it consumes the 16-bit symbols and the 256 frequencies this file produces.
//
// Sorted rank coding (Postcoder::Encode, rank.cpp:45-90) = move-to-front ranks, stored not in text order but bucketed
// by the symbol they belong to, buckets ordered by descending symbol frequency (GenerateSortedMap, rank.cpp:15-38:
// ties go to the smaller byte value). The reference runs a sequential MTF list that starts in order of first
// appearance. Both are restated without a list:
//     rank(i) = #{ c : last_c(i) > last_{T[i]}(i) },    last_c(i) = last position of byte c before i in the chunk, -1 if none
// -- the bytes used more recently than T[i] are exactly those in front of it in the list; for a first occurrence the
// right-hand side is -1 and the count is the number of distinct bytes seen so far, which is where the reference's
// first-appearance list puts it (rank.cpp:55-63). That form is data-independent and parallel over SEGMENTS of a chunk:
//   k_src_segments   per segment: byte histogram and last position of every byte inside it
//   k_src_tables     per chunk: exclusive scan of both over the segments (counts before / last position before each
//                    segment), the chunk's frequencies, the bucket starts in sorted-map order
//   k_src_rank       a warp per segment walks it in text order with last_c and the bucket cursors in shared memory: one
//                    position costs every lane 8 compares (its share of the 256 bytes) and the warp a reduction
//   k_rle0           per chunk: runs of rank 0 become the bits of (run + 1) below its leading one, every other rank r
//                    becomes r + 1 (RLE::encode, rle.cpp:22-47); two sweeps of the chunk by one block (count, emit)
#include "bwt_internal.cuh"

namespace jp {

constexpr int SRC_CHUNK_LOG2 = 20;                       // ans.hpp:33 StackSize
constexpr int SRC_CHUNK = 1 << SRC_CHUNK_LOG2;
constexpr int SRC_SEG = 4096;                            // positions per segment (a warp's share)
constexpr int SRC_SEGS_PER_CHUNK = SRC_CHUNK / SRC_SEG;
static_assert(JP_ANS_CHUNK == SRC_CHUNK, "header and kernels agree on the chunk size");

// ---- per segment: histogram and last positions (chunk-relative, +1; 0 = byte absent) ------------------------
__global__ void __launch_bounds__(256) k_src_segments(const u8* __restrict__ T, u32 len, u32 nseg, u32* __restrict__ seg_cnt, u32* __restrict__ seg_last)
{
	__shared__ u32 sh[8][256], sl[8][256];
	const u32 w = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const u32 seg = blockIdx.x * 8 + w;
	for (u32 i = lane; i < 256; i += 32) { sh[w][i] = 0; sl[w][i] = 0; }
	__syncwarp();
	if (seg < nseg) {
		const u32 base = seg * SRC_SEG, end = min(len, base + (u32)SRC_SEG);
		const u32 rel0 = base & (SRC_CHUNK - 1);
		for (u32 p = base + lane * 16; p < end; p += 512) {
			if (p + 16 <= end) {
				const uint4 q = __ldg(reinterpret_cast<const uint4*>(T + p));
				const u32 wd[4] = {q.x, q.y, q.z, q.w};
				#pragma unroll
				for (int k = 0; k < 16; k++) {
					const u32 c = (wd[k >> 2] >> ((k & 3) * 8)) & 255u;
					atomicAdd(&sh[w][c], 1u);
					atomicMax(&sl[w][c], rel0 + (p - base) + k + 1);
				}
			} else for (u32 q = p; q < end; q++) { atomicAdd(&sh[w][T[q]], 1u); atomicMax(&sl[w][T[q]], rel0 + (q - base) + 1); }
		}
	}
	__syncwarp();
	if (seg < nseg) for (u32 i = lane; i < 256; i += 32) { seg_cnt[(size_t)seg * 256 + i] = sh[w][i]; seg_last[(size_t)seg * 256 + i] = sl[w][i]; }
}

// ---- per chunk: scans over its segments, frequencies, bucket starts -------------------------------------------
// thread c = byte value c. seg_cnt[s][c] <- bucket start of c + occurrences of c before segment s; seg_last[s][c] <-
// last position (+1) of c before segment s.
__global__ void __launch_bounds__(256) k_src_tables(u32* __restrict__ seg_cnt, u32* __restrict__ seg_last, u32 nseg, i32* __restrict__ freq)
{
	__shared__ u32 F[256], sorted_f[256];
	__shared__ u32 ws[32];
	const u32 c = threadIdx.x, chunk = blockIdx.x;
	const u32 s0 = chunk * SRC_SEGS_PER_CHUNK, s1 = min(nseg, s0 + (u32)SRC_SEGS_PER_CHUNK);
	u32 cnt = 0, last = 0;
	for (u32 s = s0; s < s1; s++) {
		const size_t a = (size_t)s * 256 + c;
		const u32 h = seg_cnt[a], l = seg_last[a];
		seg_cnt[a] = cnt; seg_last[a] = last;
		cnt += h;
		if (l) last = l;
	}
	F[c] = cnt;
	freq[(size_t)chunk * 256 + c] = (i32)cnt;
	__syncthreads();
	// place in the sorted map: bytes with a larger count first, ties by byte value (rank.cpp:21-31 scans upwards with '>')
	u32 ord = 0;
	for (u32 o = 0; o < 256; o++) { const u32 f = F[o]; ord += (f > cnt || (f == cnt && o < c)) ? 1u : 0u; }
	sorted_f[ord] = cnt;
	__syncthreads();
	u32 total;
	const u32 inc = block_incl_sum(sorted_f[c], ws, &total);
	__syncthreads();
	sorted_f[c] = inc - sorted_f[c];                       // exclusive: start of the bucket at sorted place c
	__syncthreads();
	const u32 start = sorted_f[ord];
	for (u32 s = s0; s < s1; s++) seg_cnt[(size_t)s * 256 + c] += start;
}

// ---- ranks, straight into their buckets ---------------------------------------------------------------------
// A warp per segment, in text order. The warp keeps last_c (+1) and the bucket cursor of all 256 bytes in shared memory;
// for one position every lane compares the eight last_c it is responsible for (two 16-byte loads) with last_{T[i]} and
// the warp adds the counts up. A byte equal to its predecessor has rank 0 and the next slot of the same bucket: runs --
// most of a BWT -- skip the ranking and touch the tables once, at their end. (A first version kept the tables in
// registers, eight bytes per lane, and spent its time on the selects that emulate indexing them: 8.0 ms per 64 MiB.)
__global__ void __launch_bounds__(256) k_src_rank(const u8* __restrict__ T, u32 len, u32 nseg, const u32* __restrict__ seg_cnt, const u32* __restrict__ seg_last,
                                                  u8* __restrict__ ranks)
{
	__shared__ uint4 s_last4[8][64];                         // (u32 [8][256], declared as vectors for the 16-byte loads)
	__shared__ u32 s_cur[8][256];
	const u32 w = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const u32 seg = blockIdx.x * 8 + w;
	if (seg >= nseg) return;
	const u32 base = seg * SRC_SEG, end = min(len, base + (u32)SRC_SEG);
	const u32 chunk_base = base & ~(u32)(SRC_CHUNK - 1), rel0 = base - chunk_base;
	u32* last = reinterpret_cast<u32*>(s_last4[w]); u32* cur = s_cur[w];
	for (u32 i = lane; i < 256; i += 32) { last[i] = seg_last[(size_t)seg * 256 + i]; cur[i] = seg_cnt[(size_t)seg * 256 + i]; }
	__syncwarp();
	const uint4* mine4 = reinterpret_cast<const uint4*>(last + lane * 8);
	u8* out = ranks + chunk_base;
	u32 prev_sym = 0x100, dst = 0;                          // the byte of the previous position and where its rank went
	for (u32 p0 = base; p0 < end; p0 += 512) {
		// 512 positions per step: lane l holds the 16 bytes p0 + 16 l ..
		uint4 q = make_uint4(0, 0, 0, 0);
		const u32 mine = p0 + lane * 16;
		if (mine + 16 <= end) q = __ldg(reinterpret_cast<const uint4*>(T + mine));
		else { u32 wd[4] = {0, 0, 0, 0}; for (u32 b = 0; b < 16 && mine + b < end; b++) wd[b >> 2] |= (u32)T[mine + b] << ((b & 3) * 8); q = make_uint4(wd[0], wd[1], wd[2], wd[3]); }
		const u32 count = min(512u, end - p0);
		for (u32 j0 = 0; j0 < count; j0 += 4) {
			const u32 wsel = (j0 >> 2) & 3;
			const u32 word = __shfl_sync(0xffffffffu, wsel == 0 ? q.x : wsel == 1 ? q.y : wsel == 2 ? q.z : q.w, (int)(j0 >> 4));   // four positions, the same in every lane
			const u32 jn = min(4u, count - j0);
			for (u32 b = 0; b < jn; b++) {
				const u32 s = (word >> (b * 8)) & 255u;
				const u32 pos1 = rel0 + (p0 - base) + j0 + b + 1;   // chunk-relative position, +1
				if (s == prev_sym) {                              // (all branches here are warp-uniform)
					dst++;
					if (lane == 0) out[dst] = 0;
					continue;
				}
				if (prev_sym < 0x100) {                           // the run that just ended: its byte was last seen one position back
					__syncwarp();                                   // (every lane has read the tables for the previous position)
					if (lane == 0) { last[prev_sym] = pos1 - 1; cur[prev_sym] = dst + 1; }
					__syncwarp();
				}
				const u32 prev = last[s];
				const uint4 a = mine4[0], c = mine4[1];
				const u32 n = (a.x > prev) + (a.y > prev) + (a.z > prev) + (a.w > prev) + (c.x > prev) + (c.y > prev) + (c.z > prev) + (c.w > prev);
				const u32 rank = __reduce_add_sync(0xffffffffu, n);
				dst = cur[s];
				if (lane == 0) out[dst] = (u8)rank;
				prev_sym = s;
			}
		}
	}
}

// ---- RLE0 (rle.cpp:22-47) -------------------------------------------------------------------------------------
constexpr int RLE_THREADS = 1024;
constexpr int RLE_TILE = RLE_THREADS * 4;
constexpr u32 RLE_NONE = 0xffffffffu;

// For the four positions a thread owns in tile `tile` of its chunk: the number of 16-bit symbols each one emits.
// next_nz: first position >= the end of the tile that holds a non-zero rank (or the chunk end).
__device__ __forceinline__ void rle_counts(const u8* __restrict__ R, u32 cbase, u32 clen, u32 tile, u32 next_nz, u32* sm_first /*[32]*/,
                                           u32 (&sym)[4], u32 (&runlen)[4], u32 (&cnt)[4], u32* tile_first_nz)
{
	const u32 t = threadIdx.x, lane = t & 31, w = t >> 5;
	const u32 p0 = tile * RLE_TILE + t * 4;                // chunk-relative
	u32 v[4];
	#pragma unroll
	for (int k = 0; k < 4; k++) v[k] = p0 + k < clen ? R[cbase + p0 + k] : 1u;    // past the end counts as a non-zero (run stopper)
	// first non-zero at or after each of my positions: mine, else the first one in a later thread, else next_nz
	u32 mine = RLE_NONE;
	#pragma unroll
	for (int k = 3; k >= 0; k--) if (v[k] != 0) mine = p0 + k;
	u32 incl = mine;
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const u32 x = __shfl_down_sync(0xffffffffu, incl, o); if (lane + o < 32) incl = min(incl, x); }
	u32 excl = __shfl_down_sync(0xffffffffu, incl, 1);
	if (lane == 31) excl = RLE_NONE;
	__syncthreads();                                       // (sm_first is reused between calls)
	if (lane == 0) sm_first[w] = incl;
	__syncthreads();
	u32 after = min(excl, next_nz);
	for (u32 k = 31; k > w; k--) after = min(after, sm_first[k]);
	if (tile_first_nz) { u32 f = next_nz; for (int k = 31; k >= 0; k--) f = min(f, sm_first[k]); *tile_first_nz = f; }
	u32 nz = after;                                        // first non-zero at or after position p0 + k, walking k = 3..0
	const u32 before = (p0 == 0 || p0 > clen) ? 1u : R[cbase + p0 - 1];   // (the chunk's first position always starts a run)
	#pragma unroll
	for (int k = 3; k >= 0; k--) {
		if (v[k] != 0) nz = p0 + k;
		const u32 prevv = k == 0 ? before : v[k - 1];
		sym[k] = v[k]; runlen[k] = 0; cnt[k] = 0;
		if (p0 + k < clen) {
			if (v[k] != 0) cnt[k] = 1;
			else if (prevv != 0) {                           // a run of zeroes starts here
				const u32 run = min(nz, clen) - (p0 + k);
				runlen[k] = run;
				cnt[k] = 31u - (u32)__clz((int)(run + 1));    // bits of (run + 1) below its leading one
			}
		}
	}
}

__global__ void __launch_bounds__(RLE_THREADS) k_rle0(const u8* __restrict__ R, u32 len, u16* __restrict__ out, i32* __restrict__ rlen)
{
	__shared__ u32 sm_first[32];
	__shared__ u32 ws[32];
	__shared__ u32 tile_next[SRC_CHUNK / RLE_TILE + 1];      // first non-zero position at or after the end of each tile
	__shared__ u32 tile_off[SRC_CHUNK / RLE_TILE + 1];       // symbols emitted before each tile
	__shared__ u32 s_first;
	const u32 t = threadIdx.x;
	const u32 cbase = blockIdx.x * SRC_CHUNK;
	const u32 clen = min((u32)SRC_CHUNK, len - cbase);
	const u32 tiles = (clen + RLE_TILE - 1) / RLE_TILE;
	u32 sym[4], runlen[4], cnt[4];
	// sweep 1, backwards: where the next non-zero lies, how many symbols each tile emits
	u32 next_nz = clen;
	for (u32 tile = tiles; tile-- > 0;) {
		if (t == 0) tile_next[tile] = next_nz;
		u32 first;
		rle_counts(R, cbase, clen, tile, next_nz, sm_first, sym, runlen, cnt, &first);
		u32 total;
		block_incl_sum(cnt[0] + cnt[1] + cnt[2] + cnt[3], ws, &total);
		if (t == 0) { tile_off[tile] = total; s_first = first; }
		__syncthreads();
		next_nz = s_first;
	}
	__syncthreads();
	if (t == 0) { u32 run = 0; for (u32 k = 0; k < tiles; k++) { const u32 c = tile_off[k]; tile_off[k] = run; run += c; } tile_off[tiles] = run; rlen[blockIdx.x] = (i32)run; }
	__syncthreads();
	// sweep 2, forwards: emit
	u16* o = out + (size_t)cbase;
	for (u32 tile = 0; tile < tiles; tile++) {
		rle_counts(R, cbase, clen, tile, tile_next[tile], sm_first, sym, runlen, cnt, nullptr);
		const u32 mine = cnt[0] + cnt[1] + cnt[2] + cnt[3];
		u32 total;
		u32 at = tile_off[tile] + block_incl_sum(mine, ws, &total) - mine;
		#pragma unroll
		for (int k = 0; k < 4; k++) {
			if (cnt[k] == 0) continue;
			if (sym[k] != 0) o[at++] = (u16)(sym[k] + 1);
			else { const u32 L = runlen[k] + 1; for (u32 m = cnt[k]; m-- > 0;) o[at++] = (u16)((L >> m) & 1u); }
		}
	}
}

// ---- host driver ------------------------------------------------------------------------------------------
int src_rle0_device(Ctx& c, const u8* d_in, i32 len, i32* d_freq, u16* d_rle, i32* d_rlen, cudaStream_t s, jp_bwt_stats* st)
{
	st->direction = 2; st->len = len; st->nlen = len; st->device = c.device;
	if (len == 0) return JP_OK;
	const u32 n = (u32)len;
	const u32 nseg = (n + SRC_SEG - 1) / SRC_SEG, nchunk = (n + SRC_CHUNK - 1) / SRC_CHUNK;
	JP_TRY(arena_reserve(c, 2 * Arena::align((size_t)nseg * 256 * 4) + Arena::align((size_t)n + 64)));
	u32* seg_cnt = arena_take<u32>(c, (size_t)nseg * 256);
	u32* seg_last = arena_take<u32>(c, (size_t)nseg * 256);
	u8* ranks = arena_take<u8>(c, (size_t)n + 64);
	JP_CUDA(cudaEventRecord(c.ev[0], s));
	k_src_segments<<<(nseg + 7) / 8, 256, 0, s>>>(d_in, n, nseg, seg_cnt, seg_last); JP_LAUNCH(c);
	k_src_tables<<<nchunk, 256, 0, s>>>(seg_cnt, seg_last, nseg, d_freq); JP_LAUNCH(c);
	JP_CUDA(cudaEventRecord(c.ev[1], s));
	k_src_rank<<<(nseg + 7) / 8, 256, 0, s>>>(d_in, n, nseg, seg_cnt, seg_last, ranks); JP_LAUNCH(c);
	JP_CUDA(cudaEventRecord(c.ev[2], s));
	k_rle0<<<nchunk, RLE_THREADS, 0, s>>>(ranks, n, d_rle, d_rlen); JP_LAUNCH(c);
	JP_KCHECK();
	JP_CUDA(cudaEventRecord(c.ev[3], s));
	JP_CUDA(cudaStreamSynchronize(s));
	for (int i = 0; i < 3; i++) JP_CUDA(cudaEventElapsedTime(&st->ms_phase[i], c.ev[i], c.ev[i + 1]));
	JP_CUDA(cudaEventElapsedTime(&st->ms_total, c.ev[0], c.ev[3]));
	st->device_bytes = c.arena.high;
	return JP_OK;
}

} // namespace jp
