// bwt_inverse_ilp.cu -- the replay kernel of the "several sub-chains per walker thread" variant of the single-walk inverse
// (JP_BWT_INV_ILP=4; off by default, to be measured). See bwt_inverse.cu for the algorithm; this file only exists so
// that the measured kernels of bwt_inverse.cu compile exactly as they did when they were measured.
#include "bwt_internal.cuh"

namespace jp {

#include "inv_stream.cuh"

constexpr int ILP_THREADS = 128;
constexpr int ILP_WARPS   = ILP_THREADS / 32;
constexpr int N_ANCHOR    = JP_BWT_UNITS + 1;
constexpr u32 REC_INVALID = INV_REC_INVALID;
constexpr u32 WALK_BATCH  = INV_WALK_BATCH;

// One aligned 16-byte window of the output (same as flush_window in bwt_inverse.cu).
__device__ __forceinline__ void flush_window_ilp(u8* __restrict__ win, u32 lo, u32 hi, u32 a0, u32 a1, u32 a2, u32 a3)
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

template <int NS>
__global__ void __launch_bounds__(ILP_THREADS) k_inv_place_ilp(i32 n, i32 step, u32 S, const u64* __restrict__ rec,
                                                               StreamSpace sp, u8* __restrict__ out, int* __restrict__ err)
{
	constexpr u32 ROW = 32u * NS, CHUNK = ST_ROWS * ROW;
	__shared__ uint4 sdata[ILP_WARPS][CHUNK / 16];
	__shared__ u64 srec[ILP_WARPS][2][WALK_BATCH];
	if (*(volatile int*)err != 0) return;
	const u32 nodes = S + N_ANCHOR;
	const u32 lane = lane_id(), w = threadIdx.x >> 5;
	const u32 wgid = blockIdx.x * ILP_WARPS + w;
	const u32 cap = sp.cap0 + sp.cap1;
	const u8* sbytes = reinterpret_cast<const u8*>(sdata[w]);
	u32 next_t = 0, end_t = 0, base_t = 0, par = 0;
	u32 prev_batch = ST_NONE, chunk = ST_NONE, row = ST_ROWS;
	u32 id[NS], left[NS], a0[NS], a1[NS], a2[NS], a3[NS];
	i32 pos[NS], pos_end[NS];
	#pragma unroll
	for (int k = 0; k < NS; k++) { id[k] = REC_INVALID; left[k] = 0; a0[k] = a1[k] = a2[k] = a3[k] = 0; pos[k] = pos_end[k] = 0; }
	bool done = false, bad = false;
	for (u32 iter = 0; iter < cap * ST_ROWS + 2; iter++) {
		bool any_live = false;
		#pragma unroll
		for (int k = 0; k < NS; k++) {
			// ---- the ticket logic of take_ticket_log, slot by slot as in the walk
			const bool need = !done && id[k] == REC_INVALID;
			const u32 nm = __ballot_sync(0xffffffffu, need);
			if (nm != 0) {
				const u32 cnt = __popc(nm), rk = __popc(nm & lanemask_lt()), avail = end_t - next_t;
				u32 my = REC_INVALID; u64 rv = 0;
				if (need && rk < avail) { my = next_t + rk; rv = srec[w][par][my - base_t]; }
				if (cnt > avail) {
					u32 kb = 0;
					if (lane == 0) { kb = (prev_batch == ST_NONE) ? sp.batch_head[wgid] : sp.batch_next[prev_batch]; prev_batch = kb; }
					kb = __shfl_sync(0xffffffffu, kb, 0);
					if (kb >= sp.batch_cap) { dev_fail(err, DE_STREAM_OVERFLOW); bad = true; }
					else {
						const u32 base = kb * WALK_BATCH;
						par ^= 1;
						__syncwarp();
						#pragma unroll
						for (int q4 = 0; q4 < (int)WALK_BATCH / 32; q4++) {
							const u32 q = base + q4 * 32 + lane;
							srec[w][par][q4 * 32 + lane] = q < nodes ? rec[q] : pack3(0, PR_NXT_INVALID, 0);
						}
						__syncwarp();
						if (need && rk >= avail) { my = base + (rk - avail); rv = srec[w][par][my - base]; }
						next_t = base + (cnt - avail); end_t = base + WALK_BATCH; base_t = base;
					}
				} else next_t += cnt;
				if (need && my >= nodes) { done = true; my = REC_INVALID; }
				if (my != REC_INVALID && pr_nxt(rv) != PR_NXT_INVALID) {
					const u32 nxt = pr_nxt(rv), L = pr_len(rv);
					const i64 pe = (i64)(nxt - S) * step + pr_dist(rv);
					if (nxt < S || nxt >= nodes || L == 0 || pe > n || pe < (i64)L) { dev_fail(err, DE_CHAIN_RANGE); bad = true; }
					else { id[k] = my; left[k] = L; pos[k] = pos_end[k] = (i32)pe; a0[k] = a1[k] = a2[k] = a3[k] = 0; }
				}
			}
			any_live = any_live || id[k] != REC_INVALID;
		}
		if (__any_sync(0xffffffffu, bad)) break;
		if (__ballot_sync(0xffffffffu, !done || any_live) == 0) break;
		if (row == ST_ROWS) {
			u32 c = 0;
			if (lane == 0) c = (chunk == ST_NONE) ? sp.chunk_head[wgid] : sp.chunk_next[chunk];
			c = __shfl_sync(0xffffffffu, c, 0);
			if (c >= cap) { dev_fail(err, DE_STREAM_OVERFLOW); break; }
			const uint4* src = reinterpret_cast<const uint4*>(c < sp.cap0 ? sp.base0 + (size_t)c * CHUNK : sp.base1 + (size_t)(c - sp.cap0) * CHUNK);
			__syncwarp();
			#pragma unroll
			for (int q = 0; q < (int)(CHUNK / 16 / 32); q++) sdata[w][q * 32 + lane] = __ldcs(src + q * 32 + lane);
			__syncwarp();
			chunk = c; row = 0;
		}
		#pragma unroll
		for (int k = 0; k < NS; k++) {
			if (id[k] != REC_INVALID) {
				const u32 c = sbytes[row * ROW + k * 32 + lane];
				pos[k]--; left[k]--;
				const bool stop = left[k] == 0;
				const u32 wd = ((u32)pos[k] >> 2) & 3u, bits = c << (((u32)pos[k] & 3u) * 8);
				a0[k] |= (wd == 0) ? bits : 0u; a1[k] |= (wd == 1) ? bits : 0u; a2[k] |= (wd == 2) ? bits : 0u; a3[k] |= (wd == 3) ? bits : 0u;
				if (((u32)pos[k] & 15u) == 0 || stop) {
					const i32 wbase = pos[k] & ~15;
					flush_window_ilp(out + wbase, (u32)(pos[k] - wbase), (u32)min(16, pos_end[k] - wbase), a0[k], a1[k], a2[k], a3[k]);
					a0[k] = a1[k] = a2[k] = a3[k] = 0;
				}
				if (stop) id[k] = REC_INVALID;
			}
		}
		row++;
	}
}

void launch_inv_place_ilp4(int blocks, cudaStream_t s, i32 n, i32 step, u32 S, const u64* rec, StreamSpace sp, u8* out, int* err)
{
	k_inv_place_ilp<4><<<blocks, ILP_THREADS, 0, s>>>(n, step, S, rec, sp, out, err);
}

} // namespace jp
