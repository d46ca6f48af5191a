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
#include <vector>

namespace jp {

constexpr u32 AP_HEAD = 0x80000000u;     // active-set word: this slot is the first of its group
constexpr u32 AP_POS  = 0x3fffffffu;     // ... and its position in SA (blocks stay below 2^30, bwt_internal.cuh)

struct FwdMeta {
	u32 code[256];
	i32 sigma, bits, depth, key_bits;   // bits: per symbol (reported); depth: symbols per key; key_bits: bit length of the largest key
	u32 eq4;                            // aligned 16-byte vectors of one repeated byte: a cheap screen for long single-symbol runs
	u32 min_depth;                      // context-coded keys: the fewest symbols any key covers
	u32 ck_pad;
	unsigned long long ck_bits, ck_syms; // ... and the sampled positions with the total length of their codewords
	u32 hist[256];
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
			eq4 += (q.x == (q.x & 255u) * 0x01010101u && q.y == q.x && q.z == q.x && q.w == q.x) ? 1u : 0u;   // sixteen equal bytes
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
		meta->min_depth = 0x7fffffffu;
	}
}

// ---- 2. initial keys -------------------------------------------------------------------------------
// One block builds the keys of one radix tile (RS_TILE positions) and, having them in hand, also counts the
// lowest digit: the first radix pass starts from this tile histogram instead of re-reading the keys.
// COMPACT (run bypass, below): positions whose bit is set in `skip` get no key; the others are written back to back, tile
// t starting at skip_base[t] -- the output tiles no longer coincide with the input tiles, so no histogram is produced.
constexpr int KEY_TILE = RS_TILE;
template <bool COMPACT>
__global__ void __launch_bounds__(256) k_fwd_keys(const u8* __restrict__ T, i32 n, const FwdMeta* __restrict__ meta,
                                                  u64* __restrict__ keys, u32* __restrict__ vals,
                                                  u32* __restrict__ tile_hist, u32 stride,
                                                  const u32* __restrict__ skip, const u32* __restrict__ skip_base)
{
	__shared__ u32 sword[KEY_TILE / 32], wpre[KEY_TILE / 32];
	__shared__ u32 ws[32];
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
	if (COMPACT) {                                     // a tile that lies inside a run has nothing to sort: leave before reading it
		const i64 tile_end = min((i64)n, base + KEY_TILE);
		const u32 next_base = (blockIdx.x + 1 < gridDim.x) ? skip_base[blockIdx.x + 1] : 0xffffffffu;
		if (next_base == skip_base[blockIdx.x] && blockIdx.x + 1 < gridDim.x && tile_end > base) return;
	}
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
	u32 out_base = 0;
	if (COMPACT) {
		u32 keep = 0;                                  // positions of word t that get a key
		if (t < KEY_TILE / 32) {
			const i64 p0 = base + (i64)t * 32;
			const u32 valid = p0 + 32 <= n ? 0xffffffffu : (p0 < n ? (1u << (u32)(n - p0)) - 1u : 0u);
			const u32 wd = p0 < n ? skip[p0 >> 5] : 0u;
			sword[t] = keep = ~wd & valid;
		}
		u32 total;
		const u32 inc = block_incl_sum((u32)__popc(keep), ws, &total);
		if (t < KEY_TILE / 32) wpre[t] = inc - (u32)__popc(keep);
		out_base = skip_base[blockIdx.x];
	}
	__syncthreads();
	#pragma unroll 4
	for (int j = 0; j < KEY_TILE / 256; j++) {
		const int li = j * 256 + t;
		const i64 p = base + li;
		if (p < n) {
			u64 tail = part[li + q];
			if (depth & 1) tail = tail * radix + sc[li + 2 * q];
			const u64 k = (u64)part[li] * hi_scale + tail;
			if (COMPACT) {
				const u32 wd = sword[li >> 5], bit = 1u << (li & 31);
				if (wd & bit) { const u32 o = out_base + wpre[li >> 5] + (u32)__popc(wd & (bit - 1u)); keys[o] = k; vals[o] = (u32)p; }
			} else {
				keys[p] = k;
				vals[p] = (u32)p;
				atomicAdd(&h[w][(u32)k & 255u], 1u);       // low digits of text-order keys are spread: lanes rarely collide
			}
		}
	}
	if (COMPACT) return;
	__syncthreads();
	u32 sum = 0;
	#pragma unroll
	for (int k = 0; k < 8; k++) sum += h[k][t];
	tile_hist[(size_t)t * stride + blockIdx.x] = sum;
}


// ---- 2a. context-coded keys --------------------------------------------------------------------------------
// The mixed-radix key above spends log2(sigma+1) bits on every symbol whatever the text looks like: 8 symbols of source
// text, 10 of the order-2 Markov benchmark -- and it is the number of SYMBOLS in the key that decides how much of the
// block the doubling rounds still have to touch. Here the key is a bit string instead: the first `order` symbols as
// fixed-width fields, then the symbols that follow, each as a codeword of an ALPHABETIC (order-preserving, prefix-free)
// code chosen by the `order` symbols before it:
//     key(u) = sym(u) [sym(u+1)] . cw(u+order) cw(u+order+1) ...   cut after key_bits bits,
//     cw(i)  = code of T[i] in the context T[i-order .. i-1]  (a property of the position, not of the key it lands in).
// Two suffixes use the same code table up to their first difference, so the numeric order of two different keys is the
// lexicographic order of the suffixes -- however many symbols either key happens to cover. Equal keys share at least
// `order` + (complete codewords inside the cut) symbols; the minimum of that count over the block is the h the doubling
// starts from. Groups come out FINER than "equal h-prefix" (a key usually covers far more than the minimum), which the
// rounds do not mind: a group only ever has to consist of suffixes that agree on h symbols and to sit where the true
// order puts it. The code tables come from one pass over the text: which (context, symbol) pairs occur at all (exact, a
// bitmap) and how often (counted on a sample); a pair that does not occur gets no codeword, so the frequent successors
// of a context keep codewords of about -log2 p bits. With a key bit worth about a bit of information, log2 n + 13 key
// bits (5 radix passes for a 64 MiB block, where the mixed-radix key needs 8) leave the first doubling round a few per
// cent of the block instead of half of it; and keys that short fit into one 64-bit word with the position (host side).
constexpr int CK_LMAX = 12;                       // longest codeword (length-limited: a key always covers a few symbols beyond the fixed ones)
constexpr int CK_LA = 72;                         // symbols read beyond the tile: a codeword has >= 1 bit, a key < 64
constexpr u32 CK_MAX_S2 = 80;                     // order 2 up to 79 symbols + end: 80^3 table entries
struct CtxTabs { u32* present; u32* counts; u16* table; u32 S, order, entries; };

__device__ __forceinline__ u32 ck_index(u32 order, u32 S, u32 c2, u32 c1, u32 c0) { return ((order == 2 ? c2 * S + c1 : c1) * S) + c0; }

// pass over the text: exact presence bitmap (privatised in shared memory, test before set) + sampled counts.
// A thread takes 16 consecutive positions (one 16-byte load); a block 4096 at a time, every `every`-th of them counted.
__global__ void __launch_bounds__(256) k_ctx_scan(const u8* __restrict__ T, u32 n, const FwdMeta* __restrict__ meta, CtxTabs ct, u32 every)
{
	extern __shared__ __align__(16) u8 ck_smem[];
	u32* sb = reinterpret_cast<u32*>(ck_smem);
	__shared__ u16 code[256];
	const u32 t = threadIdx.x, words = (ct.entries + 31) / 32;
	const u32 order = ct.order, S = ct.S;
	code[t] = (u16)meta->code[t];
	for (u32 i = t; i < words; i += 256) sb[i] = 0;
	__syncthreads();
	const u32 chunks = (n + 4095) / 4096;
	for (u32 q = blockIdx.x; q < chunks; q += gridDim.x) {
		const u32 p0 = (q * 256 + t) * 16;
		const bool sampled = q % every == 0;             // (uniform over the block)
		u32 wd[4] = {0, 0, 0, 0};
		if (p0 + 16 <= n) { const uint4 v = __ldg(reinterpret_cast<const uint4*>(T + p0)); wd[0] = v.x; wd[1] = v.y; wd[2] = v.z; wd[3] = v.w; }
		else for (u32 k = 0; k < 16 && p0 + k < n; k++) wd[k >> 2] |= (u32)T[p0 + k] << ((k & 3) * 8);
		u32 c2 = (p0 >= 2 && p0 - 2 < n) ? code[T[p0 - 2]] : 0u, c1 = (p0 >= 1 && p0 - 1 < n) ? code[T[p0 - 1]] : 0u;
		#pragma unroll
		for (int k = 0; k < 16; k++) {
			const u32 j = p0 + k;
			const u32 c0 = code[(wd[k >> 2] >> ((k & 3) * 8)) & 255u];
			const bool ok = j >= order && j < n;
			const u32 idx = ok ? ck_index(order, S, c2, c1, c0) : 0xffffffffu;
			if (ok) {
				const u32 bit = 1u << (idx & 31);
				if (!(sb[idx >> 5] & bit)) atomicOr(&sb[idx >> 5], bit);
			}
			if (sampled) {
				const u32 peers = __match_any_sync(0xffffffffu, idx);
				if (ok && (peers & lanemask_lt()) == 0) atomicAdd(&ct.counts[idx], (u32)__popc(peers));
			}
			c2 = c1; c1 = c0;
		}
	}
	if (blockIdx.x == 0 && t == 0) {                 // the end of the text, in its context
		const u32 idx = ck_index(order, S, order == 2 ? code[T[n - 2]] : 0u, code[T[n - 1]], 0u);
		atomicOr(&sb[idx >> 5], 1u << (idx & 31));
	}
	__syncthreads();
	for (u32 i = t; i < words; i += 256) { const u32 v = sb[i]; if (v) atomicOr(&ct.present[i], v); }
}

// block = one context: a weight-balanced alphabetic tree over the symbols that occur in it (every thread walks down to
// its own leaf; all threads of a node compute the same split), at most CK_LMAX levels deep
__global__ void __launch_bounds__(288) k_ctx_codes(CtxTabs ct, FwdMeta* __restrict__ meta)
{
	__shared__ u32 W[288];                            // inclusive weight sums over the leaves of this context
	__shared__ u32 ws[32];
	const u32 t = threadIdx.x, ctx = blockIdx.x;
	const u32 idx = ctx * ct.S + t;
	const bool here = t < ct.S && ((ct.present[idx >> 5] >> (idx & 31)) & 1u);
	u32 K, wtot;
	const u32 k = block_incl_sum(here ? 1u : 0u, ws, &K) - (here ? 1u : 0u);
	const u32 cnt = here ? min(ct.counts[idx], 1u << 23) : 0u;                    // (the sample is <= 2^23 positions: the sums fit 32 bits)
	const u32 wgt = here ? 16u * cnt + 1u : 0u;
	const u32 inc = block_incl_sum(wgt, ws, &wtot);
	if (K == 0) return;
	if (here) W[k] = inc;
	__syncthreads();
	u32 lo = 0, hi = here ? K - 1 : 0u, len = 0, cw = 0;
	while (lo < hi) {
		const u32 base = lo ? W[lo - 1] : 0u, tot = W[hi] - base;
		const u32 target = base + (tot + 1) / 2;
		u32 a = lo, b = hi - 1;                       // smallest m in [lo, hi-1] with W[m] >= target, else hi-1
		while (a < b) { const u32 mid = (a + b) >> 1; if (W[mid] >= target) b = mid; else a = mid + 1; }
		u32 m = a;
		if (m > lo) {                                 // the neighbour on the left may balance better
			const u64 l1 = (u64)(W[m] - base), l0 = (u64)(W[m - 1] - base);
			const u64 d1 = 2 * l1 > tot ? 2 * l1 - tot : tot - 2 * l1, d0 = 2 * l0 > tot ? 2 * l0 - tot : tot - 2 * l0;
			if (d0 < d1) m--;
		}
		const u32 cap = 1u << (CK_LMAX - len - 1);    // leaves either side may still hold
		if (m + 1 - lo > cap) m = lo + cap - 1;
		if (hi - m > cap) m = hi - cap;
		if (k <= m) { hi = m; cw <<= 1; } else { lo = m + 1; cw = (cw << 1) | 1u; }
		len++;
	}
	if (K == 1) { len = 1; cw = 0; }                  // (a codeword has at least one bit: a key never covers more than it has bits)
	if (here) ct.table[idx] = (u16)((len << 12) | cw);
	// what the code spends on the sample: the host sizes the keys by it
	u32 bits_total, cnt_total;
	block_incl_sum(cnt * len, ws, &bits_total);
	block_incl_sum(cnt, ws, &cnt_total);
	if (t == 0 && cnt_total) { atomicAdd(&meta->ck_bits, (unsigned long long)bits_total); atomicAdd(&meta->ck_syms, (unsigned long long)cnt_total); }
}

// keys of one radix tile + the histogram of their lowest digit + the smallest number of symbols a key covers
__global__ void __launch_bounds__(256) k_fwd_keys_ctx(const u8* __restrict__ T, i32 n, FwdMeta* __restrict__ meta, CtxTabs ct,
                                                      u32 sym_bits, u32 key_bits, u32 idx_bits, u64* __restrict__ keys, u32* __restrict__ vals,
                                                      u32* __restrict__ tile_hist, u32 stride, int* __restrict__ err)
{
	constexpr int SPAN = KEY_TILE + CK_LA;
	constexpr int PER = (SPAN + 255) / 256;
	__shared__ u16 sc[SPAN];
	__shared__ u16 P[SPAN + 1];                       // bit offset of each position's codeword; P[SPAN] = end of the last one
	__shared__ u32 buf[SPAN * CK_LMAX / 32 + 4];      // the codewords back to back, most significant bit first
	__shared__ u16 code[256];
	__shared__ u32 h[8][256];
	__shared__ u32 ws[32];
	__shared__ u32 s_min;
	const int t = threadIdx.x, w = t >> 5;
	code[t] = (u16)meta->code[t];
	for (int i = t; i < 8 * 256; i += 256) (&h[0][0])[i] = 0;
	for (int i = t; i < SPAN * CK_LMAX / 32 + 4; i += 256) buf[i] = 0;
	if (t == 0) s_min = 0xffffffffu;
	__syncthreads();
	const i64 base = (i64)blockIdx.x * KEY_TILE;
	for (int i = t; i < SPAN; i += 256) { const i64 p = base + i; sc[i] = p < n ? code[T[p]] : (u16)0; }
	__syncthreads();
	const u32 order = ct.order, S = ct.S;
	// codeword of every position of the span (thread t: PER consecutive positions), their bit offsets
	u32 e[PER]; u32 mine = 0;
	#pragma unroll
	for (int k = 0; k < PER; k++) {
		const int i = t * PER + k;
		e[k] = 0;
		if (i < SPAN && i >= (int)order && base + i <= n) {
			e[k] = ct.table[ck_index(order, S, order == 2 ? sc[i - 2] : 0u, sc[i - 1], sc[i])];
			if ((e[k] >> 12) == 0) dev_fail(err, DE_FWD_ROUNDS);          // a pair the scan did not see: cannot happen
		}
		mine += e[k] >> 12;
	}
	u32 total;
	u32 off = block_incl_sum(mine, ws, &total) - mine;
	{
		// the thread's codewords are consecutive bits of the stream: assembled in a register, left-aligned from the
		// word boundary below `off`; whole words are stored, the ragged first and last ones OR-ed in
		u32 wi = off >> 5, fill = off & 31u;
		bool shared_word = fill != 0;
		u64 acc = 0;
		#pragma unroll
		for (int k = 0; k < PER; k++) {
			const int i = t * PER + k;
			if (i < SPAN) {
				P[i] = (u16)off;
				const u32 len = e[k] >> 12;
				if (len) acc |= (u64)(e[k] & 0xfffu) << (64 - fill - len);  // (fill < 32, len <= 12)
				fill += len; off += len;
				if (fill >= 32) {
					if (shared_word) atomicOr(&buf[wi], (u32)(acc >> 32)); else buf[wi] = (u32)(acc >> 32);
					shared_word = false;
					acc <<= 32; fill -= 32; wi++;
				}
			}
		}
		if (fill && (u32)(acc >> 32)) atomicOr(&buf[wi], (u32)(acc >> 32));
	}
	if (t == 255) P[SPAN] = (u16)total;
	__syncthreads();
	const u32 B = key_bits - order * sym_bits;        // bits of the coded part
	#pragma unroll 4
	for (int j = 0; j < KEY_TILE / 256; j++) {
		const int li = j * 256 + t;
		const i64 p = base + li;
		if (p < n) {
			const u32 o = P[li + order], wi = o >> 5, sh = o & 31u;
			const u32 w0 = buf[wi], w1 = buf[wi + 1], w2 = buf[wi + 2];
			const u64 win = ((u64)__funnelshift_l(w1, w0, sh) << 32) | (u64)__funnelshift_l(w2, w1, sh);
			const u64 fixed = order == 2 ? ((u64)sc[li] << sym_bits) | sc[li + 1] : (u64)sc[li];
			const u64 k = (fixed << B) | (win >> (64 - B));
			if (idx_bits) keys[p] = (k << idx_bits) | (u64)p;             // packed records: the position rides below the key
			else { keys[p] = k; vals[p] = (u32)p; }
			atomicAdd(&h[w][(u32)k & 255u], 1u);
		}
	}
	// symbols covered: order + complete codewords inside the cut, i.e. the largest c <= 64 with P[i+order+c] - P[i+order] <= B.
	// The end of the window never moves back from one position to the next: a thread walks 16 consecutive positions.
	u32 dmin = 0xffffffffu;
	{
		const int l0 = t * (KEY_TILE / 256);
		u32 end = 0;
		for (int k = 0; k < KEY_TILE / 256; k++) {
			const int li = l0 + k;
			if (base + li >= n) break;
			const u32 i0 = (u32)li + order, o = P[i0];
			if (k == 0) {
				u32 a = 0, b = 64;
				while (a < b) { const u32 mid = (a + b + 1) >> 1; if ((u32)P[i0 + mid] - o <= B) a = mid; else b = mid - 1; }
				end = i0 + a;
			} else {
				if (end < i0) end = i0;
				while (end < i0 + 64 && (u32)P[end + 1] - o <= B) end++;
			}
			dmin = min(dmin, order + (end - i0));
		}
	}
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) dmin = min(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
	if ((t & 31) == 0) atomicMin(&s_min, dmin);
	__syncthreads();
	u32 sum = 0;
	#pragma unroll
	for (int k = 0; k < 8; k++) sum += h[k][t];
	tile_hist[(size_t)t * stride + blockIdx.x] = sum;
	if (t == 0 && s_min != 0xffffffffu) atomicMin(&meta->min_depth, s_min);
}


// ---- 2b. single-symbol runs: placed, not sorted ("run bypass") -------------------------------------------
// Prefix doubling pays one round per doubling of the longest repeat, and the cheapest way to make a long repeat is a
// run of one symbol (zero pages, padding, fill bytes): a block of n equal bytes costs log2(n / depth) rounds over the
// whole block. Runs have an exact shortcut. Let r(v) be the length of the run of T[v] that starts at v, and call v a
// RUN SUFFIX when r(v) >= depth, the number of symbols in the initial key: its key is c^depth, so the run suffixes of a
// symbol c are exactly the initial group of that key, and inside it suffix v = c^r(v) X(v), where X(v) starts with a
// symbol other than c (or is empty). Comparing c^r X with c^r' X', r < r', is decided at position r: X's first symbol
// against c. Hence the order inside the group is: first the suffixes whose run is followed by a SMALLER symbol (or by
// the end of the text), by ascending r; then those followed by a larger one, by descending r; ties -- equal class and
// r: they come from different runs -- by the order of the X's.
// So run suffixes never enter the initial sort. A maximal run of length L >= depth contributes the suffixes r = depth..L
// of its (symbol, class) BUCKET; level r of a bucket holds one suffix of every run of the bucket with L >= r. With the
// bucket's runs listed by descending L (a sort of the RUNS, of which there are few), a run's index in that list is its
// index inside every level it reaches, and the first position of level r is a prefix sum over the list:
//     offset(r) = sum over runs with L < r of (L - depth + 1)  +  #{L >= r} * (r - depth).
// Every run suffix therefore computes its own position in SA; the sorted non-run suffixes are laid around the buckets
// (a bucket sits where its key c^depth would sort); levels with one member are final at once, the others are ordinary
// groups that the doubling rounds refine by what follows the runs. A block of one repeated byte is finished without a
// single round.
constexpr int RUN_TILE = KEY_TILE;
constexpr u32 RUN_NONE = 0xffffffffu;
constexpr u32 RUN_MAXL = (1u << 30) - 1;

struct RunTabs {
	u32* bits;                  // bit v: suffix v is a run suffix                                    (n/32 + 2 words)
	u32* tile_first; u32* tile_next;   // per tile of RUN_TILE positions: run-end bookkeeping of the detection
	u32* tile_base;             // per tile: run suffixes in it; after the scan: non-run suffixes before it
	u32* run_s; u32* run_L;     // runs of length >= depth, in no particular order                        (cap entries)
	u64* rk[2]; u32* rv[2];     // (bucket, descending length) sort of the runs
	u32* PS;                    // exclusive prefix sums of (L - depth + 1) over the sorted runs; PS[R] = all run suffixes
	u32* bstart; u32* bend;     // sorted-run range of each bucket                                       (512 each)
	u32* lo;                    // first SA position of each bucket                                        (512)
	u64* shift_key; u32* shift_cum;   // run key of each symbol code, run suffixes of the codes below it   (256 / 257)
	u32* counters;              // [0] runs seen [1] run suffixes [2] run suffixes whose level holds more than one run
	u32 cap;
};

__device__ __forceinline__ u32 warp_min_u32(u32 v)
{
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}

// The same three kernels serve repeats of any period p (see "periodic repeats" further down): position j BREAKS the
// period when T[j] != T[j + p] or j + p >= n; for p = 1 that is the last position of a run.
__device__ __forceinline__ bool per_break(const u8* __restrict__ T, u32 n, u32 p, u32 j) { return j + p >= n || T[j] != T[j + p]; }

// first break inside the tile; RUN_NONE if there is none
__global__ void __launch_bounds__(256) k_run_first(const u8* __restrict__ T, u32 n, u32 p, u32* __restrict__ tile_first)
{
	__shared__ u32 wbest[8];
	const u32 base = blockIdx.x * RUN_TILE;
	u32 best = RUN_NONE;
	#pragma unroll 4
	for (int i = 0; i < RUN_TILE / 256; i++) {
		const u32 j = base + i * 256 + threadIdx.x;
		if (j < n && per_break(T, n, p, j)) best = min(best, j);
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

// r(j) = (first break at or after j) - j + p = length of the p-periodic stretch that starts at j; bits = (r(j) >= thr).
// LIST (run bypass, p = 1): the runs of length >= thr that START in the tile are appended to the run list; tile_base[tile] =
// run suffixes in the tile; counters[0] += runs, counters[1] += run suffixes. Otherwise (periodic repeats): RL[j] = r(j).
template <bool LIST>
__global__ void __launch_bounds__(256) k_run_fill(const u8* __restrict__ T, u32 n, u32 p, const u32* __restrict__ tile_next, u32 thr, RunTabs rt,
                                                  u32* __restrict__ RL)
{
	__shared__ u32 end_at[RUN_TILE];            // last position of the run each position of the tile lies in
	__shared__ u32 wfirst[8];
	__shared__ u32 s_cnt;
	constexpr int PER = RUN_TILE / 256;
	const u32 t = threadIdx.x, lane = t & 31, w = t >> 5;
	const u32 base = blockIdx.x * RUN_TILE;
	if (t == 0) s_cnt = 0;
	// each thread owns PER consecutive positions: its first run end, if any
	u32 mine = RUN_NONE;
	#pragma unroll 4
	for (int k = PER - 1; k >= 0; k--) {
		const u32 j = base + t * PER + k;
		if (j < n && per_break(T, n, p, j)) mine = j;
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
	#pragma unroll 4
	for (int k = PER - 1; k >= 0; k--) {
		const u32 j = base + t * PER + k;
		if (j < n && per_break(T, n, p, j)) carry = j;
		end_at[t * PER + k] = carry;
	}
	__syncthreads();
	u32 cnt = 0;
	#pragma unroll 4
	for (int i = 0; i < PER; i++) {
		const u32 j = base + i * 256 + t;
		u32 r = 0;
		if (j < n) r = end_at[i * 256 + t] - j + p;
		const bool rs = j < n && r >= thr;
		const u32 m = __ballot_sync(0xffffffffu, rs);
		if (lane == 0 && base + i * 256 + (t & ~31u) < n) { rt.bits[(base + i * 256 + t) >> 5] = m; cnt += __popc(m); }
		if (LIST) {
			if (rs && (j == 0 || T[j - 1] != T[j])) {         // a run of length >= thr starts here
				const u32 k = atomicAdd(&rt.counters[0], 1u);
				if (k < rt.cap) { rt.run_s[k] = j; rt.run_L[k] = r; }
			}
		} else if (j < n) RL[j] = r;
	}
	if (lane == 0 && cnt) atomicAdd(&s_cnt, cnt);
	__syncthreads();
	if (t == 0) { if (rt.tile_base) rt.tile_base[blockIdx.x] = s_cnt; if (s_cnt) atomicAdd(&rt.counters[1], s_cnt); }
}

// single block: tile_base[t] <- number of NON-run suffixes in the tiles before t
__global__ void __launch_bounds__(1024) k_run_tile_scan(u32* __restrict__ tile_base, u32 tiles, u32 n)
{
	__shared__ u32 ws[32];
	const u32 t = threadIdx.x;
	const u32 per = (tiles + 1023) / 1024;
	const u32 lo = min(tiles, t * per), hi = min(tiles, lo + per);
	u32 s = 0;
	for (u32 k = lo; k < hi; k++) s += min((u32)RUN_TILE, n - k * RUN_TILE) - tile_base[k];
	u32 total;
	u32 run = block_incl_sum(s, ws, &total) - s;
	for (u32 k = lo; k < hi; k++) { const u32 c = min((u32)RUN_TILE, n - k * RUN_TILE) - tile_base[k]; tile_base[k] = run; run += c; }
}

// sort key of a run: (bucket = 2 * (code - 1) + class, descending length)
__global__ void __launch_bounds__(256) k_run_keys(const u8* __restrict__ T, u32 n, const FwdMeta* __restrict__ meta, RunTabs rt, u32 R)
{
	const u32 k = blockIdx.x * 256 + threadIdx.x;
	if (k >= R) return;
	const u32 s = rt.run_s[k], L = rt.run_L[k], e = s + L;
	const u32 c = T[s];
	const u32 cls = (e >= n || T[e] < c) ? 0u : 1u;          // what follows the run: smaller (or the end of the text) / larger
	const u32 b = (meta->code[c] - 1u) * 2u + cls;
	rt.rk[0][k] = ((u64)b << 30) | (u64)(RUN_MAXL - L);
	rt.rv[0][k] = k;
}

// single block over the sorted runs: PS, bucket ranges
__global__ void __launch_bounds__(1024) k_run_tables(const u64* __restrict__ sk, u32 R, u32 depth, RunTabs rt)
{
	__shared__ u32 ws[32];
	const u32 t = threadIdx.x;
	if (t < 512) { rt.bstart[t] = R; rt.bend[t] = 0; }
	__syncthreads();
	const u32 per = (R + 1023) / 1024;
	const u32 lo = min(R, t * per), hi = min(R, lo + per);
	u32 s = 0;
	for (u32 i = lo; i < hi; i++) s += (RUN_MAXL - (u32)(sk[i] & RUN_MAXL)) - depth + 1;
	u32 total;
	u32 run = block_incl_sum(s, ws, &total) - s;
	for (u32 i = lo; i < hi; i++) {
		const u64 key = sk[i];
		const u32 b = (u32)(key >> 30);
		rt.PS[i] = run;
		run += (RUN_MAXL - (u32)(key & RUN_MAXL)) - depth + 1;
		if (i == 0 || (u32)(sk[i - 1] >> 30) != b) rt.bstart[b] = i;
		if (i + 1 == R || (u32)(sk[i + 1] >> 30) != b) rt.bend[b] = i + 1;
	}
	if (t == 1023) rt.PS[R] = total;
}

// one thread per symbol code: where its bucket pair sits in SA, and the shift table of the non-run suffixes
__global__ void __launch_bounds__(256) k_run_lo(const u64* __restrict__ K, u32 n_sorted, const FwdMeta* __restrict__ meta, RunTabs rt)
{
	__shared__ u32 ws[32];
	const u32 c = threadIdx.x;                       // code - 1
	const u32 sigma = (u32)meta->sigma, depth = (u32)meta->depth;
	const u64 radix = (u64)sigma + 1;
	u64 rep = 0;                                     // 1 + radix + ... + radix^(depth-1): the key of code 1 repeated
	for (u32 d = 0; d < depth; d++) rep = rep * radix + 1;
	u32 m0 = 0, m1 = 0, less = 0;
	u64 rkey = ~0ull;
	if (c < sigma) {
		rkey = rep * (u64)(c + 1);
		const u32 b0 = 2 * c, b1 = 2 * c + 1;
		if (rt.bend[b0] > rt.bstart[b0]) m0 = rt.PS[rt.bend[b0]] - rt.PS[rt.bstart[b0]];
		if (rt.bend[b1] > rt.bstart[b1]) m1 = rt.PS[rt.bend[b1]] - rt.PS[rt.bstart[b1]];
		u32 lo = 0, hi = n_sorted;                   // first sorted key >= rkey
		while (lo < hi) { const u32 mid = lo + ((hi - lo) >> 1); if (K[mid] < rkey) lo = mid + 1; else hi = mid; }
		less = lo;
	}
	u32 total;
	const u32 cum = block_incl_sum(m0 + m1, ws, &total) - (m0 + m1);
	rt.shift_key[c] = rkey;
	rt.shift_cum[c + 1] = cum + m0 + m1;
	if (c == 0) rt.shift_cum[0] = 0;
	if (c < sigma) { rt.lo[2 * c] = less + cum; rt.lo[2 * c + 1] = less + cum + m0; }
}

// the sorted non-run suffixes take their places around the buckets; head flags from the key changes
__global__ void __launch_bounds__(256) k_place_sorted(const u64* __restrict__ K, const u32* __restrict__ V, u32 n_sorted, RunTabs rt,
                                                      u32* __restrict__ SA, u8* __restrict__ F)
{
	__shared__ u64 skey[256];
	__shared__ u32 scum[257];
	const u32 t = threadIdx.x;
	skey[t] = rt.shift_key[t];
	scum[t] = rt.shift_cum[t];
	if (t == 0) scum[256] = rt.shift_cum[256];
	__syncthreads();
	const u32 j = blockIdx.x * 256 + t;
	if (j >= n_sorted) return;
	const u64 k = K[j];
	u32 lo = 0;                                       // number of run keys below k (none equals it)
	#pragma unroll
	for (u32 st = 128; st > 0; st >>= 1) if (skey[lo + st - 1] < k) lo += st;
	if (lo == 255 && skey[255] < k) lo = 256;
	const u32 pos = j + scum[lo];
	SA[pos] = V[j];
	F[pos] = (j == 0 || K[j - 1] != k) ? 1 : 0;
}

// every run suffix computes its own place
// A level that only one run reaches is a singleton for good: its rank is written here (consecutive suffixes, consecutive
// positions: coalesced) and its flag says so (F = 2), so that the grouping step neither reads nor ranks it again.
__global__ void __launch_bounds__(256) k_place_runs(const u64* __restrict__ sk, const u32* __restrict__ sv, u32 R, u32 M, u32 depth, RunTabs rt,
                                                    u32* __restrict__ SA, u8* __restrict__ F, u32* __restrict__ ISA)
{
	const u32 u = blockIdx.x * 256 + threadIdx.x;
	u32 q = 0;
	if (u < M) {
		u32 lo = 0, hi = R;                            // last run i with PS[i] <= u
		while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (rt.PS[mid] <= u) lo = mid; else hi = mid; }
		const u32 i = lo, toff = u - rt.PS[i];
		const u64 key = sk[i];
		const u32 b = (u32)(key >> 30), L = RUN_MAXL - (u32)(key & RUN_MAXL);
		const u32 r = L - toff, v = rt.run_s[sv[i]] + toff;
		const u32 bs = rt.bstart[b], be = rt.bend[b];
		// q = runs of the bucket that reach level r = sorted keys in [bs, be) not above (b, r)
		const u64 lim = ((u64)b << 30) | (u64)(RUN_MAXL - r);
		u32 a = bs, z = be;
		while (a < z) { const u32 mid = a + ((z - a) >> 1); if (sk[mid] <= lim) a = mid + 1; else z = mid; }
		q = a - bs;
		const u32 idx = i - bs;
		const u32 off = (rt.PS[be] - rt.PS[bs + q]) + q * (r - depth);
		const u32 Mb = rt.PS[be] - rt.PS[bs];
		const u32 pos = (b & 1u) == 0 ? rt.lo[b] + off + idx : rt.lo[b] + Mb - off - q + idx;
		SA[pos] = v;
		if (q == 1) { F[pos] = 2; ISA[v] = pos + 1; }
		else F[pos] = idx == 0 ? 1 : 0;
	}
	const u32 shared = __popc(__ballot_sync(0xffffffffu, q > 1));     // run suffixes in levels that several runs reach
	if (shared && (threadIdx.x & 31) == 0) atomicAdd(&rt.counters[2], shared);
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
template <typename KT>
__global__ void __launch_bounds__(256) k_fwd_flags(const KT* __restrict__ K, u32 n, u8* __restrict__ F)
{
	const u32 j0 = (blockIdx.x * 256 + threadIdx.x) * 4;
	if (j0 >= n) return;
	KT prev = j0 ? K[j0 - 1] : (KT)~K[0];
	u32 w = 0;
	#pragma unroll
	for (int b = 0; b < 4; b++) {
		const u32 j = j0 + b;
		if (j < n) { const KT k = K[j]; w |= (k != prev ? 1u : 0u) << (8 * b); prev = k; }
	}
	if (j0 + 4 <= n) *reinterpret_cast<u32*>(F + j0) = w;
	else for (int b = 0; b < 4 && j0 + b < n; b++) F[j0 + b] = (u8)(w >> (8 * b));
}

// packed records (key << idx_bits | position) in sorted order -> the suffix array and the head flags
__global__ void __launch_bounds__(256) k_fwd_unpack(const u64* __restrict__ K, u32 n, u32 idx_bits, u32* __restrict__ SA, u8* __restrict__ F)
{
	const u32 j0 = (blockIdx.x * 256 + threadIdx.x) * 4;
	if (j0 >= n) return;
	const u64 mask = ((u64)1 << idx_bits) - 1;
	u64 prev = j0 ? K[j0 - 1] >> idx_bits : ~(K[0] >> idx_bits);
	u32 w = 0, v[4] = {0, 0, 0, 0};
	#pragma unroll
	for (int b = 0; b < 4; b++) {
		const u32 j = j0 + b;
		if (j < n) { const u64 r = K[j], k = r >> idx_bits; v[b] = (u32)(r & mask); w |= (k != prev ? 1u : 0u) << (8 * b); prev = k; }
	}
	if (j0 + 4 <= n) { *reinterpret_cast<u32*>(F + j0) = w; *reinterpret_cast<uint4*>(SA + j0) = make_uint4(v[0], v[1], v[2], v[3]); }
	else for (int b = 0; b < 4 && j0 + b < n; b++) { F[j0 + b] = (u8)(w >> (8 * b)); SA[j0 + b] = v[b]; }
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
//   F[j] == 2: a singleton whose rank is already in ISA (run bypass): not read, not ranked; a tile made of nothing else
//                  says so in tile_final[tile] and k_isa_scatter skips it (R = 0 marks such slots in mixed tiles).
__global__ void __launch_bounds__(GS_THREADS) k_grp_apply(const u8* __restrict__ F, const u32* AP, const u32* __restrict__ SA, u32 A,
                                                          const GAgg* __restrict__ agg, u32* __restrict__ ISA, u32* R, u32* __restrict__ VS,
                                                          u32* __restrict__ APn, u32* __restrict__ tile_final)
{
	__shared__ i32 wmh[GS_SUB][8];
	__shared__ u32 wpk[GS_SUB][8];
	const int t = threadIdx.x, lane = t & 31, w = t >> 5;
	const u32 base = blockIdx.x * GS_TILE;
	const GAgg carry0 = agg[blockIdx.x];

	u32 v[GS_SUB], p[GS_SUB], fl[GS_SUB];     // fl: bit0 valid, bit1 head, bit2 survivor, bit3 surviving head, bit4 final (rank already placed)
	int all_final = 1;
	#pragma unroll
	for (int s = 0; s < GS_SUB; s++) {
		const u32 j = base + s * GS_THREADS + t;
		v[s] = 0; p[s] = 0; fl[s] = 0;
		if (j < A) {
			const u32 f = F[j];
			const bool head = f != 0, nhead = (j + 1 == A) || (F[j + 1] != 0);
			fl[s] = 1u | (head ? 2u : 0u) | (!(head && nhead) ? 4u : 0u) | ((head && !nhead) ? 8u : 0u) | (f == 2 ? 16u : 0u);
			p[s] = AP ? (AP[j] & AP_POS) : j;
			if (f != 2) { v[s] = SA[p[s]]; all_final = 0; }
		}
	}
	all_final = __syncthreads_and(all_final);
	if (threadIdx.x == 0 && tile_final) tile_final[blockIdx.x] = (u32)all_final;
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
			if (fl[s] & 16u) { if (R) R[j] = 0; }
			else if (R) { R[j] = (u32)fmh + 1u; if (VS) VS[j] = v[s]; } // ranks leave in slot order; k_isa_scatter places them
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
                                                     int region_log2, u32 tiles, const u32* __restrict__ tile_final)
{
	const u32 region = blockIdx.x / tiles, tile = blockIdx.x % tiles;
	if (tile_final && tile_final[tile]) return;                 // nothing but ranks that are already in place
	const u32 base = tile * 2048 + threadIdx.x;
	u32 v[8], r[8];
	#pragma unroll
	for (int i = 0; i < 8; i++) {
		const u32 j = base + i * 256;
		v[i] = 0xffffffffu; r[i] = 0;
		if (j < A) { v[i] = __ldcs(V + j); r[i] = __ldcs(R + j); }
	}
	#pragma unroll
	for (int i = 0; i < 8; i++) if (v[i] != 0xffffffffu && r[i] != 0 && (region_log2 >= 32 || (v[i] >> region_log2) == region)) ISA[v[i]] = r[i];
}

// ---- 5. doubling rounds ----------------------------------------------------------------------------------
// key2 of suffix v in the round with offset h. (The template parameter selects the periodic-run keys added below.)
struct PerSkip { const u32* rl; const u32* bits; const u8* T; u32 p; u32 first; u32 redirect; u32 deep; };

template <bool PS>
__device__ __forceinline__ u32 key2_of(u32 v, u32 h, u32 n, const u32* __restrict__ ISA, const PerSkip& ps, int* __restrict__ err)
{
	if (PS && ps.redirect) {
		// reduced sort of a periodic block (see "periodic blocks are sorted through their representatives"): a skipped
		// suffix answers with the rank of the representative of its stretch and phase
		u32 x = v + h;
		if (x > n) { dev_fail(err, DE_FWD_RANGE); x = n; }
		if (x < n && ((__ldg(&ps.bits[x >> 5]) >> (x & 31)) & 1u)) x += ps.p * ((__ldg(&ps.rl[x]) - ps.deep) / ps.p);
		return __ldg(&ISA[x]);
	}
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

// ---- 5p. periodic repeats --------------------------------------------------------------------------------
// The run argument holds for any period p. Let r(v) be the length of the longest p-periodic string that starts at v
// (T[v + i] = T[v + i + p] for i < r(v) - p). Two suffixes v, v' that share their first p symbols and have r(v) < r(v')
// agree on r(v) symbols and differ at offset r(v), where v breaks the period and v' does not: their order is the order
// of T[v + r] against T[v + r - p], v's own symbols. So once the groups are h-groups with h >= p, the members of a group
// with r >= h -- a property of the shared h symbols, hence of the whole group or of none of it -- are ordered by
// (break goes down: ascending r | break goes up: descending r), ties by the rank of the suffix at v + r. In key2_of:
//   * pass A (one round, h unchanged): a flagged suffix (bit set: r(v) >= the h of this round) takes key2 = r, or 2n+1-r;
//   * pass B (the same h again) and every later round: a flagged suffix with r(v) >= h takes ISA[v + r(v)], any other
//     suffix the ordinary ISA[v + h]. Either way the key is uniform inside a group (its members have the same r) and the
//     round leaves every suffix at least 2h-ordered, which is what the ordinary keys of the next round rely on.
// A text of period p with sparse defects (the `repetitive` generator of BASELINE configs[2]: p = 1021, one flipped bit
// every 64 Ki) then needs log2(p / depth) ordinary rounds plus these two instead of log2(64 Ki / depth).
// The period is not searched for in the text: after a stable sort the members of a group are in text order, so in a
// periodic text the distance between neighbours of a group IS the period. A sample of those distances is histogrammed
// after the initial step; the mode, if it covers most of the block, is p.
constexpr u32 PER_MAXP = 65536;
constexpr u32 PER_SAMPLE = 64;
__global__ void __launch_bounds__(256) k_per_sample(const u32* __restrict__ AP, const u32* __restrict__ SA, u32 A, u32* __restrict__ hist)
{
	const u32 j = (blockIdx.x * 256 + threadIdx.x) * PER_SAMPLE;
	u32 d = 0;
	if (j + 1 < A && !(AP[j + 1] & AP_HEAD)) {
		const u32 a = SA[AP[j] & AP_POS], b = SA[AP[j + 1] & AP_POS];
		if (b > a && b - a < PER_MAXP) d = b - a;
	}
	const u32 peers = __match_any_sync(0xffffffffu, d);
	if (d && (peers & lanemask_lt()) == 0) atomicAdd(&hist[d], (u32)__popc(peers));
}
// single block: out[0] = most frequent distance, out[1] = its count
__global__ void __launch_bounds__(1024) k_per_pick(const u32* __restrict__ hist, u32* __restrict__ out)
{
	__shared__ u32 sc[32], sd[32];
	const u32 t = threadIdx.x;
	u32 best = 0, bd = 0;
	for (u32 d = 1 + t; d < PER_MAXP; d += 1024) { const u32 c = hist[d]; if (c > best) { best = c; bd = d; } }
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		const u32 oc = __shfl_xor_sync(0xffffffffu, best, o), od = __shfl_xor_sync(0xffffffffu, bd, o);
		if (oc > best || (oc == best && od < bd)) { best = oc; bd = od; }
	}
	if ((t & 31) == 0) { sc[t >> 5] = best; sd[t >> 5] = bd; }
	__syncthreads();
	if (t == 0) {
		for (int k = 1; k < 32; k++) if (sc[k] > best || (sc[k] == best && sd[k] < bd)) { best = sc[k]; bd = sd[k]; }
		out[0] = bd; out[1] = best;
	}
}

// ---- 5q. periodic blocks are sorted through their representatives -----------------------------------------------
// The rounds with h < p are the expensive part of a periodic block: nothing can be skipped in them (members of a group
// share fewer than p symbols), yet nearly all of the block is copies of the same p rotations. When the period is known
// BEFORE the sort they are run on a fraction of the block. With H = the first doubling distance >= p: a suffix is DEEP
// when r(v) >= H (its first H symbols are periodic), and among the deep suffixes of one stretch and phase -- v, v + p,
// v + 2p, ... -- the last one is their REPRESENTATIVE; the others (r(v) >= H + p) are skipped. Representatives and
// non-deep suffixes -- a few percent of the block -- are sorted by the ordinary machinery: initial keys, then the
// rounds h = depth .. H/2, in which a lookup that lands on a skipped suffix x is answered by its representative
// x + p * floor((r(x) - H) / p), whose first H symbols are x's own. That leaves them exactly H-ordered. A skipped suffix
// shares its first H symbols with its representative, so giving it the representative's rank as its key and sorting the
// WHOLE block once by that 3-byte key yields the H-order of all suffixes (ties = H-equal), from which the ordinary flow
// continues at h = H with the repeat-length keys of the previous section. `repetitive(64 MiB)`: the six rounds over
// 67 M suffixes (6.2 ms each) become six rounds over 4 M, and the eight 12-byte radix passes of the initial sort three
// 8-byte ones. The period is found in the text itself: for 64 sampled positions, the distance to the next occurrence
// of the 32 bytes that follow; if half of the samples agree, that distance is p (0.02 ms on a block without a period).
constexpr u32 PROBE_SAMPLES = 64, PROBE_WIN = 32, PROBE_MAXP = 16384;   // (longer periods are still found after the initial step)
// one block per sample: thread t tries the distances t + 1, t + 257, ...; the block stops at the first round that has a hit
__global__ void __launch_bounds__(256) k_per_probe(const u8* __restrict__ T, u32 n, u32* __restrict__ out)
{
	__shared__ u32 s_found;
	const u32 t = threadIdx.x, sample = blockIdx.x;
	const u32 span = PROBE_MAXP + PROBE_WIN;
	if (t == 0) s_found = 0xffffffffu;
	__syncthreads();
	if (n > 2 * span) {
		const u32 a = (u32)(((u64)sample * (u64)(n - span - 1)) / PROBE_SAMPLES);
		const u32 w0 = (u32)T[a] | ((u32)T[a + 1] << 8) | ((u32)T[a + 2] << 16) | ((u32)T[a + 3] << 24);
		for (u32 d0 = 1; d0 < PROBE_MAXP; d0 += 2048) {                  // eight distances per thread between two barriers
			bool any = false;
			#pragma unroll 2
			for (u32 k8 = 0; k8 < 8; k8++) {
				const u32 d = d0 + k8 * 256 + t;
				if (d >= PROBE_MAXP) break;
				const u8* q = T + a + d;
				if (((u32)q[0] | ((u32)q[1] << 8) | ((u32)q[2] << 16) | ((u32)q[3] << 24)) != w0) continue;
				bool hit = true;
				for (u32 k = 4; k < PROBE_WIN; k++) if (q[k] != T[a + k]) { hit = false; break; }
				if (hit) { atomicMin(&s_found, d); any = true; break; }
			}
			if (__syncthreads_or(any ? 1 : 0)) break;
		}
	}
	__syncthreads();
	if (t == 0) out[sample] = s_found == 0xffffffffu ? 0u : s_found;
}

// key of every suffix for the one sort of the whole block: the rank of itself (reduced set) or of its representative
__global__ void __launch_bounds__(256) k_per_expand(const u32* __restrict__ ISA, const u32* __restrict__ RL, const u32* __restrict__ skip, u32 p, u32 deep, u32 n,
                                                    u32* __restrict__ K, u32* __restrict__ V)
{
	const u32 v = blockIdx.x * 256 + threadIdx.x;
	if (v >= n) return;
	u32 x = v;
	if ((skip[v >> 5] >> (v & 31)) & 1u) x += p * ((RL[v] - deep) / p);
	K[v] = ISA[x];
	V[v] = v;
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
#ifndef SG_PAIR_MAX_CFG
#define SG_PAIR_MAX_CFG 1024
#endif
constexpr u32 SG_PAIR_MAX = SG_PAIR_MAX_CFG;             // longest group the all-pairs rank refinement takes on

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
	// The warp totals, in (row, warp) order, are one sequence of 128: warp 0 turns wf into "latest head before this
	// (row, warp)", warp 1 turns wb into "first group end after it" -- 4 entries per lane and one warp scan each -- and
	// every thread then reads its 16 entries instead of looping over all 128 twice.
	constexpr int SG_NW = SG_THREADS / 32, SG_TOT = SG_ITEMS * SG_NW;
	static_assert(SG_TOT == 128, "4 warp totals per lane");
	if (w == 0) {
		u32* f1 = &wf[0][0];
		u32 v[4], run = 0;
		#pragma unroll
		for (int q = 0; q < 4; q++) { v[q] = f1[lane * 4 + q]; run = max(run, v[q]); }
		const u32 inc = (u32)warp_incl_max((i32)run);
		u32 ex = __shfl_up_sync(0xffffffffu, inc, 1);
		if (lane == 0) ex = 0;
		#pragma unroll
		for (int q = 0; q < 4; q++) { f1[lane * 4 + q] = ex; ex = max(ex, v[q]); }
	} else if (w == 1) {
		u32* b1 = &wb[0][0];
		u32 v[4], run = 0x1fffu;
		#pragma unroll
		for (int q = 0; q < 4; q++) { v[q] = b1[SG_TOT - 1 - (lane * 4 + q)]; run = min(run, v[q]); }
		u32 inc = run;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const u32 x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc = min(inc, x); }
		u32 ex = __shfl_up_sync(0xffffffffu, inc, 1);
		if (lane == 0) ex = 0x1fffu;
		#pragma unroll
		for (int q = 0; q < 4; q++) { b1[SG_TOT - 1 - (lane * 4 + q)] = ex; ex = min(ex, v[q]); }
	}
	__syncthreads();
	int big = 0;
	#pragma unroll
	for (int i = 0; i < SG_ITEMS; i++) {
		const u32 gs = max(pk[i] & 0x1fffu, wf[i][w]);
		const u32 ge = min((pk[i] >> 13) & 0x1fffu, wb[i][w]);
		pk[i] = (pk[i] & ~0x3ffffffu) | gs | (ge << 13);
		if ((pk[i] >> 28) & 1u) big |= (ge - gs > SG_PAIR_MAX);
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

// (a batch = the large groups g0.. whose suffixes are numbers x0 .. x0 + count of the back-to-back layout)
template <bool PS>
__global__ void __launch_bounds__(256) k_large_extract(const u32* __restrict__ AP, const u32* __restrict__ SA, const u32* __restrict__ ISA, u32 h, u32 n, int rank_bits,
                                                       const u32* __restrict__ lg_head, const u32* __restrict__ lg_off, u32 ng, u32 g0, u32 x0, u32 count,
                                                       u64* __restrict__ LK, u32* __restrict__ LV, int* __restrict__ err, PerSkip ps)
{
	const u32 x = blockIdx.x * 256 + threadIdx.x;
	if (x >= count) return;
	const u32 g = large_group_of(lg_off, ng, x0 + x);
	const u32 v = SA[AP[lg_head[g] + (x0 + x - lg_off[g])] & AP_POS];
	const u32 k2 = key2_of<PS>(v, h, n, ISA, ps, err);
	LK[x] = ((u64)(g - g0) << rank_bits) | (u64)k2;
	LV[x] = v;
}

__global__ void __launch_bounds__(256) k_large_writeback(const u64* __restrict__ LK, const u32* __restrict__ LV, int rank_bits,
                                                         const u32* __restrict__ lg_head, const u32* __restrict__ lg_off, u32 g0, u32 x0, u32 count,
                                                         const u32* __restrict__ AP, u32* __restrict__ SA, u8* __restrict__ F)
{
	const u32 x = blockIdx.x * 256 + threadIdx.x;
	if (x >= count) return;
	const u64 k = LK[x];
	const u32 g = g0 + (u32)(k >> rank_bits);
	const u32 o = lg_off[g];
	const u32 slot = lg_head[g] + (x0 + x - o);
	SA[AP[slot] & AP_POS] = LV[x];
	F[slot] = (x0 + x == o || LK[x - 1] != k) ? 1 : 0;
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

// The same for blocks whose text no longer fits the L2: one sweep over SA per 64 MiB REGION of the text (the blocks of a
// launch are ordered by region), each taking the bytes that lie in its region and OR-ing them into the zeroed output --
// the gathers stay L2-resident, as the rank placement's stores do (256 MiB block: 4.0 -> 3.3 ms for this step).
__global__ void __launch_bounds__(256) k_fwd_emit_regions(const u8* __restrict__ T, const u32* __restrict__ SA,
                                                          const u32* __restrict__ ISA, i32 n, u8* __restrict__ out, int region_log2, u32 tiles)
{
	const u32 region = blockIdx.x / tiles, tile = blockIdx.x % tiles;
	const i64 o0 = ((i64)tile * 256 + threadIdx.x) * 4;
	if (o0 >= n) return;
	const i64 idx0 = (i64)ISA[0] - 1;
	u32 acc = 0;
	#pragma unroll
	for (int b = 0; b < 4; b++) {
		const i64 o = o0 + b;
		if (o < n) {
			i64 src;
			if (o == 0) src = (i64)n - 1;
			else { const i64 i = (o <= idx0) ? o - 1 : o; src = (i64)__ldcs(SA + i) - 1; }
			if ((u32)(src >> region_log2) == region) acc |= (u32)T[src] << (8 * b);
		}
	}
	if (acc) atomicOr(reinterpret_cast<u32*>(out + o0), acc);          // (bytes beyond n in the last word get zeros OR-ed in: unchanged)
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
	return 6 * Arena::align((N + 2) * 4) + Arena::align((rtiles + 4) * 256 * 4) + Arena::align(256 * 4) +
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
	b.F = d_flags ? d_flags : arena_take<u8>(c, N + 64);
	b.rb.dnext = getenv("JP_BWT_RADIX_NO_DIGIT_BYTES") ? nullptr : b.F;
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
		k_isa_scatter<<<regions * stiles, 256, 0, s>>>(V, R, A, b.ISA, b.isa_region_log2, stiles, b.queue); JP_LAUNCH(c);
		return JP_OK;
	}
	// many regions (blocks over 64 MiB): one radix partition pass buckets the (suffix, rank) pairs by region, then a
	// single ordered sweep
	for (u32 off = 0; off < A; off += pcap) {
		const u32 cnt = std::min(pcap, A - off);
		if (radix_partition_u32(V + off, R + off, pv, pr, cnt, b.isa_region_log2, b.rb.tile_hist, b.rb.totals, s, &c.launches) != 0) { set_error_detail("radix partition setup failed"); return JP_ERR_CUDA; }
		const u32 stiles = (cnt + 2047) / 2048;
		k_isa_scatter<<<stiles, 256, 0, s>>>(pv, pr, cnt, b.ISA, 32, stiles, nullptr); JP_LAUNCH(c);
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
	// (b.queue, idle between the sorting kernels of a round, carries the per-tile "all ranks already placed" words)
	if (!staged) { k_grp_apply<<<tiles, GS_THREADS, 0, s>>>(b.F, APin, b.SA, A, b.agg, b.ISA, nullptr, nullptr, APout, b.queue); JP_LAUNCH(c); }
	else if (APin == nullptr) {
		// initial step: the suffix ids are SA itself; ranks go to the spare unit; VS and the second active buffer are free
		k_grp_apply<<<tiles, GS_THREADS, 0, s>>>(b.F, nullptr, b.SA, A, b.agg, b.ISA, b.X, nullptr, APout, b.queue); JP_LAUNCH(c);
		JP_TRY(place_ranks(c, b, b.SA, b.X, A, b.VS, APout == b.AP[0] ? b.AP[1] : b.AP[0], (u32)(b.usz / 4), s));
	} else {
		// a round: ranks overwrite the (dead) input slots, suffix ids are staged in VS; the spare unit holds the bucketed pairs
		u32* R = const_cast<u32*>(APin);
		k_grp_apply<<<tiles, GS_THREADS, 0, s>>>(b.F, APin, b.SA, A, b.agg, b.ISA, R, b.VS, APout, b.queue); JP_LAUNCH(c);
		const u32 half = (u32)(b.usz / 8);
		JP_TRY(place_ranks(c, b, b.VS, R, A, b.X, b.X + half, half, s));
	}
	JP_KCHECK();
	JP_CUDA(cudaMemcpyAsync(c.h_small + 8, b.counters, 2 * sizeof(u32), cudaMemcpyDeviceToHost, s));
	JP_CUDA(cudaMemcpyAsync(c.h_small, b.err, sizeof(int), cudaMemcpyDeviceToHost, s));
	JP_CUDA(cudaStreamSynchronize(s));
	return JP_OK;
}

// The run bypass keeps its tables in the second arena (only blocks that pass the screen pay for them).
static int run_tabs_alloc(Ctx& c, i32 n, u32 tiles, RunTabs& rt)
{
	const size_t N = (size_t)n;
	rt.cap = (u32)std::max<size_t>(4096, N / 256);
	const size_t cap = rt.cap;
	size_t off = 0;
	auto take = [&](size_t bytes) { const size_t o = off; off += Arena::align(bytes); return o; };
	const size_t o_bits = take((N / 32 + 2) * 4), o_tf = take((size_t)tiles * 4), o_tn = take((size_t)tiles * 4), o_tb = take((size_t)tiles * 4),
	             o_rs = take(cap * 4), o_rl = take(cap * 4), o_k0 = take(cap * 8), o_k1 = take(cap * 8), o_v0 = take(cap * 4), o_v1 = take(cap * 4),
	             o_ps = take((cap + 1) * 4), o_bs = take(512 * 4), o_be = take(512 * 4), o_lo = take(512 * 4), o_sk = take(256 * 8),
	             o_sc = take(257 * 4), o_ct = take(64);
	JP_TRY(arena2_reserve(c, off));
	u8* a = c.arena2.base;
	rt.bits = (u32*)(a + o_bits); rt.tile_first = (u32*)(a + o_tf); rt.tile_next = (u32*)(a + o_tn); rt.tile_base = (u32*)(a + o_tb);
	rt.run_s = (u32*)(a + o_rs); rt.run_L = (u32*)(a + o_rl);
	rt.rk[0] = (u64*)(a + o_k0); rt.rk[1] = (u64*)(a + o_k1); rt.rv[0] = (u32*)(a + o_v0); rt.rv[1] = (u32*)(a + o_v1);
	rt.PS = (u32*)(a + o_ps); rt.bstart = (u32*)(a + o_bs); rt.bend = (u32*)(a + o_be); rt.lo = (u32*)(a + o_lo);
	rt.shift_key = (u64*)(a + o_sk); rt.shift_cum = (u32*)(a + o_sc); rt.counters = (u32*)(a + o_ct);
	return JP_OK;
}

// Repeat lengths r(v) of period p for every position (second arena): bit v of `bits` = r(v) >= thr; returns the number of
// set bits in *count (after the caller's next synchronisation, through h_small[15]).
struct PerTabs { u32* bits; u32* tile_first; u32* tile_next; u32* tile_base; u32* RL; u32* counters; u32* probe; size_t keep; };
static int per_tabs_alloc(Ctx& c, i32 n, u32 tiles, PerTabs& t)
{
	size_t off = 0;
	auto take = [&](size_t bytes) { const size_t o = off; off += Arena::align(bytes); return o; };
	const size_t o_bits = take(((size_t)n / 32 + 2) * 4), o_tf = take((size_t)tiles * 4), o_tn = take((size_t)tiles * 4), o_tb = take((size_t)tiles * 4),
	             o_rl = take((size_t)n * 4), o_ct = take(64);
	t.keep = off;                                                       // everything up to here survives a later growth of the arena
	const size_t o_hist = take((size_t)PER_MAXP * 4 + 64);
	JP_TRY(arena2_reserve(c, off));
	u8* a2 = c.arena2.base;
	t.bits = (u32*)(a2 + o_bits); t.tile_first = (u32*)(a2 + o_tf); t.tile_next = (u32*)(a2 + o_tn); t.tile_base = (u32*)(a2 + o_tb);
	t.RL = (u32*)(a2 + o_rl); t.counters = (u32*)(a2 + o_ct); t.probe = (u32*)(a2 + o_hist);
	return JP_OK;
}
static int per_tabs_fill(Ctx& c, const u8* d_T, i32 n, u32 tiles, u32 p, u32 thr, PerTabs& t, cudaStream_t s)
{
	RunTabs pt = {};
	pt.bits = t.bits; pt.tile_first = t.tile_first; pt.tile_next = t.tile_next; pt.tile_base = t.tile_base; pt.counters = t.counters;
	JP_CUDA(cudaMemsetAsync(t.counters, 0, 64, s));
	k_run_first<<<tiles, 256, 0, s>>>(d_T, (u32)n, p, t.tile_first); JP_LAUNCH(c);
	k_run_scan<<<1, 1024, 0, s>>>(t.tile_first, tiles, t.tile_next); JP_LAUNCH(c);
	k_run_fill<false><<<tiles, 256, 0, s>>>(d_T, (u32)n, p, t.tile_next, thr, pt, t.RL); JP_LAUNCH(c);
	JP_KCHECK();
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
	// period probe (sorting a periodic block through its representatives): the samples land in the radix histogram area
	int want_probe = n >= (1 << 20) ? 1 : 0;
	if (const char* e = getenv("JP_BWT_FWD_REDUCED")) want_probe = atoi(e) != 0 && (u32)n > 2 * (PROBE_MAXP + PROBE_WIN) ? 1 : 0;
	if (want_probe) { k_per_probe<<<PROBE_SAMPLES, 256, 0, s>>>(d_T, (u32)n, b.rb.tile_hist); JP_LAUNCH(c); }
	JP_KCHECK();
	static_assert(PROBE_SAMPLES <= 192, "the samples are read back through the context's pinned words");
	u32* h_probe = reinterpret_cast<u32*>(c.h_small + 64);
	if (want_probe) JP_CUDA(cudaMemcpyAsync(h_probe, b.rb.tile_hist, PROBE_SAMPLES * sizeof(u32), cudaMemcpyDeviceToHost, s));
	JP_CUDA(cudaMemcpyAsync(c.h_small + 16, &b.meta->sigma, 5 * sizeof(i32), cudaMemcpyDeviceToHost, s)); // sigma, bits, depth, key_bits, eq4
	JP_CUDA(cudaStreamSynchronize(s));
	const int bits = c.h_small[17], depth = c.h_small[18], key_bits0 = c.h_small[19];
	if (bits < 1 || bits > 9 || depth < 7 || depth > 63 || key_bits0 < 1 || key_bits0 > 63) { set_error_detail("symbol remap gave bits=%d depth=%d key bits=%d", bits, depth, key_bits0); return JP_ERR_INTERNAL; }
	st->symbol_bits = bits; st->initial_depth = depth;
	const u32 ktiles = (u32)(((size_t)n + KEY_TILE - 1) / KEY_TILE);
	if (cudaFuncSetAttribute(k_seg_sort<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG_SMEM_LIGHT) != cudaSuccess ||
	    cudaFuncSetAttribute(k_seg_sort_radix<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG_SMEM) != cudaSuccess ||
	    cudaFuncSetAttribute(k_seg_sort<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG_SMEM_LIGHT) != cudaSuccess ||
	    cudaFuncSetAttribute(k_seg_sort_radix<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG_SMEM) != cudaSuccess) { set_error_detail("k_seg_sort smem attribute"); return JP_ERR_CUDA; }

	// ---- state of the doubling rounds (shared by the reduced sort of a periodic block and the main loop) ----
	int rank_bits = bit_length((u64)n);                                 // key2 is a rank <= n (a repeat-length key <= 2n + 1)
	int act = 0;
	u32 A = 0, G = 0;
	u64 sectors = 2ull * (u64)n;
	i64 h = depth;
	int rounds = 0;
	PerSkip ps; ps.rl = nullptr; ps.bits = nullptr; ps.T = d_T; ps.p = 1; ps.first = 0; ps.redirect = 0; ps.deep = 0;
	bool periodic = false, reduced = false;
	int pass = 0;                                                       // 0: pass A still to come, 1: pass B next (same h), 2: both done
	i64 h_a = 0;                                                        // the h of pass A: first h >= p
	size_t keep2 = 0;                                                   // bytes at the start of the second arena that later growth must keep
	const bool trace_rounds = getenv("JP_BWT_TRACE_ROUNDS") != nullptr;      // one stderr line per doubling round (host wall time)
	double t_round = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
	std::vector<u32> h_off;
	auto run_rounds = [&](i64 h_stop) -> int {
		while (A > 0 && h < h_stop) {
			if (c.h_small[0] != 0) return map_dev_err(c.h_small[0]);
			if (rounds >= JP_BWT_MAX_ROUNDS || h > (i64)n) { set_error_detail("doubling stuck: round %d h=%lld active=%u", rounds, (long long)h, A); return JP_ERR_INTERNAL; }
			st->active_fraction[rounds] = (float)((double)A / (double)n);
			sectors += 2ull * A;
			double large_frac_now = 0.0;
			const u32* AP = b.AP[act];
			const bool use_ps = reduced || (periodic && h >= h_a);           // repeat-length keys from the first h >= p on; redirected lookups in the reduced sort
			ps.redirect = reduced ? 1u : 0u;
			ps.first = (!reduced && use_ps && pass == 0) ? 1u : 0u;
			// short groups: fused gather + warp-level rank refinement in shared memory; longer ones are queued for the
			// shared-memory radix kernel; groups longer than a window are reported for the large-group route
			const u32 nwin = (A + SG_WIN - 1) / SG_WIN;
			JP_CUDA(cudaMemsetAsync(b.counters + 2, 0, 4 * sizeof(u32), s));
			if (use_ps) k_seg_sort<true><<<nwin, SG_THREADS, SG_SMEM_LIGHT, s>>>(AP, b.SA, A, b.ISA, (u32)h, (u32)n, b.F, b.counters, b.queue, b.win_first, b.win_large, b.err, ps);
			else k_seg_sort<false><<<nwin, SG_THREADS, SG_SMEM_LIGHT, s>>>(AP, b.SA, A, b.ISA, (u32)h, (u32)n, b.F, b.counters, b.queue, b.win_first, b.win_large, b.err, ps);
			JP_LAUNCH(c);
			JP_KCHECK();
			JP_CUDA(cudaMemcpyAsync(c.h_small + 10, b.counters + 2, 2 * sizeof(u32), cudaMemcpyDeviceToHost, s));
			JP_CUDA(cudaStreamSynchronize(s));
			const bool large = c.h_small[10] != 0;
			const u32 queued = (u32)c.h_small[11];
			if (queued) {
				if (use_ps) k_seg_sort_radix<true><<<queued, SG_THREADS, SG_SMEM, s>>>(AP, b.SA, A, b.ISA, (u32)h, (u32)n, rank_bits, b.F, b.queue, b.err, ps);
				else k_seg_sort_radix<false><<<queued, SG_THREADS, SG_SMEM, s>>>(AP, b.SA, A, b.ISA, (u32)h, (u32)n, rank_bits, b.F, b.queue, b.err, ps);
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
				// Scratch of the sort: the spare unit, the staging unit and the idle active buffer hold (8 + 8 + 2 x 4) bytes for
				// up to N/2 suffixes. More than that (blocks that are mostly long repeats) goes through in batches of whole groups;
				// a single group beyond N/2 falls back on the second arena.
				const u32 cap = (u32)(b.usz / 8);
				if (total > cap) {
					h_off.resize((size_t)ng + 1);
					JP_CUDA(cudaMemcpyAsync(h_off.data(), b.lg_off, ((size_t)ng + 1) * 4, cudaMemcpyDeviceToHost, s));
					JP_CUDA(cudaStreamSynchronize(s));
				}
				for (u32 g0 = 0; g0 < ng;) {
					u32 g1 = ng, x0 = 0, count = total;
					if (total > cap) {
						g1 = g0 + 1;
						while (g1 < ng && h_off[g1 + 1] - h_off[g0] <= cap) g1++;
						x0 = h_off[g0]; count = h_off[g1] - x0;
					}
					RadixBuffers lb = b.rb;
					lb.dnext = nullptr;                         // the flag bytes of this round are live in b.F
					if (count <= cap) {
						lb.k[0] = reinterpret_cast<u64*>(b.X); lb.k[1] = reinterpret_cast<u64*>(b.VS);
						lb.v[0] = b.AP[act ^ 1]; lb.v[1] = b.AP[act ^ 1] + b.usz / 8;
					} else {
						const size_t T8 = Arena::align((size_t)count * 8), T4 = Arena::align((size_t)count * 4), base = Arena::align(keep2);
						JP_TRY(arena2_reserve(c, base + 2 * T8 + 2 * T4, keep2));
						u8* a2 = c.arena2.base;
						if (periodic) { ps.bits = (const u32*)a2; ps.rl = (const u32*)(a2 + ((const u8*)ps.rl - (const u8*)ps.bits)); }
						lb.k[0] = reinterpret_cast<u64*>(a2 + base); lb.k[1] = reinterpret_cast<u64*>(a2 + base + T8);
						lb.v[0] = reinterpret_cast<u32*>(a2 + base + 2 * T8); lb.v[1] = reinterpret_cast<u32*>(a2 + base + 2 * T8 + T4);
					}
					if (use_ps) k_large_extract<true><<<(count + 255) / 256, 256, 0, s>>>(AP, b.SA, b.ISA, (u32)h, (u32)n, rank_bits, b.lg_head, b.lg_off, ng, g0, x0, count, lb.k[0], lb.v[0], b.err, ps);
					else k_large_extract<false><<<(count + 255) / 256, 256, 0, s>>>(AP, b.SA, b.ISA, (u32)h, (u32)n, rank_bits, b.lg_head, b.lg_off, ng, g0, x0, count, lb.k[0], lb.v[0], b.err, ps);
					JP_LAUNCH(c);
					const int key_bits = rank_bits + bit_length((u64)(g1 - g0 - 1));
					const int lc = radix_sort_pairs(lb, 0, count, 0, key_bits, s, &c.launches);
					if (lc < 0) { set_error_detail("radix sort setup failed"); return JP_ERR_CUDA; }
					k_large_writeback<<<(count + 255) / 256, 256, 0, s>>>(lb.k[lc], lb.v[lc], rank_bits, b.lg_head, b.lg_off, g0, x0, count, AP, b.SA, b.F); JP_LAUNCH(c);
					JP_KCHECK();
					g0 = g1;
				}
			}
			JP_TRY(group_step(c, b, AP, b.AP[act ^ 1], A, s));
			if (trace_rounds) {
				const double t1 = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
				fprintf(stderr, "[jp_bwt round] %d h=%lld%s active=%u groups=%u -> active=%u groups=%u large_frac=%.4f %.3f ms\n", rounds, (long long)h,
				        reduced ? " (representatives)" : use_ps ? (pass == 0 ? " (repeat lengths)" : " (repeat jumps)") : "", A, G, (u32)c.h_small[8], (u32)c.h_small[9], large_frac_now, t1 - t_round);
				t_round = t1;
			}
			act ^= 1;
			A = (u32)c.h_small[8]; G = (u32)c.h_small[9];
			if (!reduced && use_ps && pass == 0) pass = 1;                   // pass B repeats this h
			else { h *= 2; if (!reduced && use_ps) pass = 2; }
			rounds++;
		}
		return JP_OK;
	};

	// ---- a periodic block is sorted through its representatives ("5q") -------------------------------------------
	bool premode = false;
	PerTabs pt = {};
	u32 per_p = 0;
	if (want_probe) {
		std::sort(h_probe, h_probe + PROBE_SAMPLES);
		u32 best = 0, best_cnt = 0;
		for (u32 i = 0; i < PROBE_SAMPLES;) { u32 j = i; while (j < PROBE_SAMPLES && h_probe[j] == h_probe[i]) j++; if (h_probe[i] != 0 && j - i > best_cnt) { best = h_probe[i]; best_cnt = j - i; } i = j; }
		i64 H = depth;
		while (H < (i64)best) H *= 2;
		if (best >= 2 && best_cnt * 2 >= PROBE_SAMPLES && H * 8 <= (i64)n) {
			per_p = best;
			JP_TRY(per_tabs_alloc(c, n, ktiles, pt));
			JP_TRY(per_tabs_fill(c, d_T, n, ktiles, per_p, (u32)(H + per_p), pt, s));      // bit = skipped: deep, and not the last deep one of its stretch and phase
			JP_CUDA(cudaMemcpyAsync(c.h_small + 14, pt.counters, 2 * sizeof(u32), cudaMemcpyDeviceToHost, s));
			JP_CUDA(cudaStreamSynchronize(s));
			const u32 skipped = (u32)c.h_small[15];
			premode = (u64)skipped * 2 >= (u64)n;                          // worth it when at least half of the block is skipped
			if (premode) {
				const u32 n_red = (u32)n - skipped;
				h_a = H;
				k_run_tile_scan<<<1, 1024, 0, s>>>(pt.tile_base, ktiles, (u32)n); JP_LAUNCH(c);
				k_fwd_keys<true><<<ktiles, 256, 0, s>>>(d_T, n, b.meta, b.rb.k[0], b.rb.v[0], nullptr, 0, pt.bits, pt.tile_base); JP_LAUNCH(c);
				JP_KCHECK();
				JP_CUDA(cudaEventRecord(c.ev[1], s));
				const int cur = radix_sort_pairs(b.rb, 0, n_red, 0, key_bits0, s, &c.launches);
				if (cur < 0) { set_error_detail("radix sort setup failed"); return JP_ERR_CUDA; }
				b.ISA = reinterpret_cast<u32*>(b.unit[cur ? 0 : 2]);
				b.X = reinterpret_cast<u32*>(b.unit[cur ? 1 : 3]);
				b.AP[0] = reinterpret_cast<u32*>(b.unit[cur ? 2 : 0]);
				b.AP[1] = reinterpret_cast<u32*>(b.unit[cur ? 3 : 1]);
				b.SA = b.rb.v[cur]; b.VS = b.rb.v[cur ^ 1];
				k_fwd_flags<u64><<<(n_red + 1023) / 1024, 256, 0, s>>>(b.rb.k[cur], n_red, b.F); JP_LAUNCH(c);
				JP_CUDA(cudaMemsetAsync(b.ISA + n, 0, sizeof(u32), s));
				JP_TRY(group_step(c, b, nullptr, b.AP[0], n_red, s));         // ranks 1..n_red of the reduced set
				A = (u32)c.h_small[8]; G = (u32)c.h_small[9];
				ps.rl = pt.RL; ps.bits = pt.bits; ps.p = per_p; ps.deep = (u32)H;
				reduced = true; act = 0;
				JP_TRY(run_rounds(H));                                        // h = depth .. H/2: exactly H-ordered afterwards
				reduced = false;
				if (c.h_small[0] != 0) return map_dev_err(c.h_small[0]);
				// every suffix takes the rank of its representative; one sort of the whole block by that key
				u32* free_units[5]; int nf = 0;
				for (int i = 0; i < 6; i++) if (reinterpret_cast<u32*>(b.unit[i]) != b.ISA) free_units[nf++] = reinterpret_cast<u32*>(b.unit[i]);
				u32* k32[2] = {free_units[0], free_units[1]};
				u32* v32[2] = {free_units[2], free_units[3]};
				k_per_expand<<<(u32)(((size_t)n + 255) / 256), 256, 0, s>>>(b.ISA, pt.RL, pt.bits, per_p, (u32)H, (u32)n, k32[0], v32[0]); JP_LAUNCH(c);
				const int cc = radix_sort_pairs32(k32, v32, 0, (u32)n, bit_length((u64)n_red + 1), b.rb.tile_hist, b.rb.totals, s, &c.launches);
				if (cc < 0) { set_error_detail("radix sort setup failed"); return JP_ERR_CUDA; }
				JP_KCHECK();
				JP_CUDA(cudaEventRecord(c.ev[2], s));
				k_fwd_flags<u32><<<(u32)(((size_t)n + 1023) / 1024), 256, 0, s>>>(k32[cc], (u32)n, b.F); JP_LAUNCH(c);
				b.SA = v32[cc];
				b.VS = v32[cc ^ 1]; b.X = k32[cc ^ 1]; b.AP[0] = free_units[4]; b.AP[1] = k32[cc];   // (k32[cc] is dead once the flags exist)
				// from here on: the ordinary flow at h = H, with the repeat-length keys
				JP_TRY(per_tabs_fill(c, d_T, n, ktiles, per_p, (u32)H, pt, s));                 // bit = r(v) >= H: the suffixes pass A orders by length
				h = H;
				periodic = true; pass = 0;
				keep2 = pt.keep;
				rank_bits = bit_length(2 * (u64)n + 1);
				st->period = (i32)per_p;
			}
		}
	}

	// Run bypass: worth its detection pass when a visible share of the block sits in long single-symbol runs. The histogram
	// kernel counted the aligned 16-byte vectors of one repeated byte (every run of 31 bytes or more contains one; the
	// 8-space indentation of source text does not): 1/64 of the block in such vectors engages the detection.
	// JP_BWT_FWD_BYPASS=0/1 overrides the screen.
	RunTabs rt = {};
	u32 R = 0, M = 0;
	bool bypass = !premode && (u64)(u32)c.h_small[20] * 16 * 64 >= (u64)n;
	if (const char* e = getenv("JP_BWT_FWD_BYPASS")) bypass = !premode && atoi(e) != 0;
	if (bypass) {
		JP_TRY(run_tabs_alloc(c, n, ktiles, rt));
		JP_CUDA(cudaMemsetAsync(rt.counters, 0, 64, s));
		k_run_first<<<ktiles, 256, 0, s>>>(d_T, (u32)n, 1u, rt.tile_first); JP_LAUNCH(c);
		k_run_scan<<<1, 1024, 0, s>>>(rt.tile_first, ktiles, rt.tile_next); JP_LAUNCH(c);
		k_run_fill<true><<<ktiles, 256, 0, s>>>(d_T, (u32)n, 1u, rt.tile_next, (u32)depth, rt, nullptr); JP_LAUNCH(c);
		JP_KCHECK();
		JP_CUDA(cudaMemcpyAsync(c.h_small + 14, rt.counters, 2 * sizeof(u32), cudaMemcpyDeviceToHost, s));
		JP_CUDA(cudaStreamSynchronize(s));
		R = (u32)c.h_small[14]; M = (u32)c.h_small[15];
		if (R == 0 || R > rt.cap) bypass = false;          // nothing to place, or more runs than the tables hold: plain doubling
	}
	// Context-coded keys for ordinary blocks (the run bypass and the representatives of a periodic block reason about a fixed
	// number of symbols per key and keep the mixed-radix form). JP_BWT_FWD_CTXKEYS=0/1 overrides the size threshold,
	// JP_BWT_FWD_KEYPASSES the number of radix passes.
	const u32 ck_S = (u32)c.h_small[16] + 1;
	const u32 ck_order = ck_S <= CK_MAX_S2 ? 2u : 1u;
	const u32 ck_entries = ck_order == 2 ? ck_S * ck_S * ck_S : ck_S * ck_S;
	bool ctx_keys = !premode && !bypass && n >= (1 << 20);
	if (const char* e = getenv("JP_BWT_FWD_CTXKEYS")) ctx_keys = !premode && !bypass && atoi(e) != 0 && n >= 256;
	if (ctx_keys && Arena::align(((size_t)ck_entries + 31) / 32 * 4) + Arena::align((size_t)ck_entries * 4) + (size_t)ck_entries * 2 > 2 * b.usz) ctx_keys = false;
	int ck_key_bits = 8 * std::min(8, std::max(4, (bit_length((u64)n) + 13 + 7) / 8));
	if (const char* e = getenv("JP_BWT_FWD_KEYPASSES")) ck_key_bits = 8 * std::min(8, std::max(3, atoi(e)));
	if (ck_key_bits > 63) ck_key_bits = 63;
	if (ck_key_bits < (int)(ck_order * (u32)bits) + 16) ctx_keys = false;
	// Packed records: when short keys will do and the positions leave room for them in a 64-bit word (blocks up to 64 MiB),
	// key and position travel as one word and the sort moves 16 bytes per pass and suffix instead of 24.
	const int ck_idx_bits = bit_length((u64)n - 1);
	bool ck_packed = ctx_keys && ck_key_bits <= 40 && 64 - ck_idx_bits >= ck_key_bits - 2 && !getenv("JP_BWT_FWD_KEYPASSES");
	if (const char* e = getenv("JP_BWT_FWD_PACKED")) ck_packed = ctx_keys && atoi(e) != 0 && 64 - ck_idx_bits >= (int)(ck_order * (u32)bits) + 16;
	if (!premode) {
		const u32 n_sorted = bypass ? (u32)n - M : (u32)n;
		if (bypass) {
			st->bypass_suffixes = (i32)M; st->bypass_runs = (i32)R;
			k_run_tile_scan<<<1, 1024, 0, s>>>(rt.tile_base, ktiles, (u32)n); JP_LAUNCH(c);
			k_fwd_keys<true><<<ktiles, 256, 0, s>>>(d_T, n, b.meta, b.rb.k[0], b.rb.v[0], nullptr, 0, rt.bits, rt.tile_base); JP_LAUNCH(c);
		} else if (ctx_keys) {
			// context-coded keys: the tables live in the second key buffer until the first radix pass overwrites it
			CtxTabs ct; ct.S = ck_S; ct.order = ck_order; ct.entries = ck_entries;
			const size_t words = ((size_t)ck_entries + 31) / 32;
			ct.present = reinterpret_cast<u32*>(b.unit[2]);
			ct.counts = ct.present + Arena::align(words * 4) / 4;
			ct.table = reinterpret_cast<u16*>(ct.counts + Arena::align((size_t)ck_entries * 4) / 4);
			JP_CUDA(cudaMemsetAsync(ct.present, 0, Arena::align(words * 4) + Arena::align((size_t)ck_entries * 4), s));
			const u32 every = std::max(1u, (u32)n >> 22);                    // counts from ~4 M sampled positions
			const u32 chunks = ((u32)n + 4095) / 4096;
			const u32 per_sm = (u32)std::max<size_t>(1, std::min<size_t>(6, (200u << 10) / (words * 4 + 1024)));
			if (cudaFuncSetAttribute(k_ctx_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(words * 4)) != cudaSuccess) { set_error_detail("k_ctx_scan smem attribute"); return JP_ERR_CUDA; }
			k_ctx_scan<<<std::min(chunks, (u32)c.sm_count * per_sm), 256, words * 4, s>>>(d_T, (u32)n, b.meta, ct, every); JP_LAUNCH(c);
			k_ctx_codes<<<ck_order == 2 ? ck_S * ck_S : ck_S, 288, 0, s>>>(ct, b.meta); JP_LAUNCH(c);
			if (!getenv("JP_BWT_FWD_KEYPASSES")) {
				// Key length. With a code that fits, a key bit is worth about a bit: log2 n + 13 of them leave a few per cent of
				// the block to the rounds (measured: 5 passes beat 6 and 8 on the order-2 Markov and the uniform 64 MiB blocks).
				// An order-1 code that compresses well says the text has structure, and then it has more than one symbol of
				// context can see (source text: 3.8 of 7.7 bits): all 63 bits pay there.
				unsigned long long* h_ck = reinterpret_cast<unsigned long long*>(c.h_small + 24);
				JP_CUDA(cudaMemcpyAsync(h_ck, &b.meta->ck_bits, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
				JP_CUDA(cudaStreamSynchronize(s));
				const double rate = h_ck[1] ? (double)h_ck[0] / (double)h_ck[1] : (double)bits;
				if (ck_order == 1 && rate < 0.75 * (double)bits) { ck_key_bits = 63; if (!getenv("JP_BWT_FWD_PACKED")) ck_packed = false; }
				st->symbol_bits = (i32)(rate + 0.5);
			}
			if (ck_packed) ck_key_bits = std::min(ck_key_bits, 64 - ck_idx_bits);
			k_fwd_keys_ctx<<<ktiles, 256, 0, s>>>(d_T, n, b.meta, ct, (u32)bits, (u32)ck_key_bits, ck_packed ? (u32)ck_idx_bits : 0u, b.rb.k[0], b.rb.v[0], b.rb.tile_hist,
			                                     rs_stride((u32)radix_tiles((size_t)n)), b.err); JP_LAUNCH(c);
			JP_CUDA(cudaMemcpyAsync(c.h_small + 22, &b.meta->min_depth, sizeof(u32), cudaMemcpyDeviceToHost, s));   // read after the grouping step's sync
		} else {
			k_fwd_keys<false><<<ktiles, 256, 0, s>>>(d_T, n, b.meta, b.rb.k[0], b.rb.v[0], b.rb.tile_hist,
			                                         rs_stride((u32)radix_tiles((size_t)n)), nullptr, nullptr); JP_LAUNCH(c);
		}
		JP_KCHECK();
		JP_CUDA(cudaEventRecord(c.ev[1], s));
		const int cur = (ctx_keys && ck_packed) ? radix_sort_pairs(b.rb, 0, n_sorted, ck_idx_bits, ck_idx_bits + ck_key_bits, s, &c.launches, true, /*keys_only=*/true)
		                                        : radix_sort_pairs(b.rb, 0, n_sorted, 0, ctx_keys ? ck_key_bits : key_bits0, s, &c.launches, /*first_hist_ready=*/!bypass);
		if (cur < 0) { set_error_detail("radix sort setup failed"); return JP_ERR_CUDA; }
		JP_KCHECK();
		JP_CUDA(cudaEventRecord(c.ev[2], s));

		// The key buffers and the id buffers take new roles from here on (the keys are read one last time, for the head flags).
		b.ISA = reinterpret_cast<u32*>(b.unit[cur ? 0 : 2]);      // first half of the other key buffer
		b.X = reinterpret_cast<u32*>(b.unit[cur ? 1 : 3]);        // ... and its second half
		b.AP[0] = reinterpret_cast<u32*>(b.unit[cur ? 2 : 0]);    // the sorted keys' own buffer, dead once the flags exist
		b.AP[1] = reinterpret_cast<u32*>(b.unit[cur ? 3 : 1]);
		if (!bypass) {
			// the sorted suffix ids ARE the suffix array
			if (ctx_keys && ck_packed) {
				b.SA = b.rb.v[0]; b.VS = b.rb.v[1];
				k_fwd_unpack<<<(u32)(((size_t)n + 1023) / 1024), 256, 0, s>>>(b.rb.k[cur], (u32)n, (u32)ck_idx_bits, b.SA, b.F); JP_LAUNCH(c);
			} else {
				b.SA = b.rb.v[cur]; b.VS = b.rb.v[cur ^ 1];
				k_fwd_flags<u64><<<(u32)(((size_t)n + 1023) / 1024), 256, 0, s>>>(b.rb.k[cur], (u32)n, b.F); JP_LAUNCH(c);
			}
		} else {
			// the suffix array is assembled in the other id buffer: sorted suffixes around the buckets, run suffixes inside them
			b.SA = b.rb.v[cur ^ 1]; b.VS = b.rb.v[cur];
			RadixBuffers lb = b.rb;
			lb.k[0] = rt.rk[0]; lb.k[1] = rt.rk[1]; lb.v[0] = rt.rv[0]; lb.v[1] = rt.rv[1]; lb.dnext = nullptr;
			k_run_keys<<<(R + 255) / 256, 256, 0, s>>>(d_T, (u32)n, b.meta, rt, R); JP_LAUNCH(c);
			const int rc = radix_sort_pairs(lb, 0, R, 0, 30 + 9, s, &c.launches);
			if (rc < 0) { set_error_detail("radix sort setup failed"); return JP_ERR_CUDA; }
			k_run_tables<<<1, 1024, 0, s>>>(lb.k[rc], R, (u32)depth, rt); JP_LAUNCH(c);
			k_run_lo<<<1, 256, 0, s>>>(b.rb.k[cur], n_sorted, b.meta, rt); JP_LAUNCH(c);
			k_place_sorted<<<(n_sorted + 255) / 256, 256, 0, s>>>(b.rb.k[cur], b.rb.v[cur], n_sorted, rt, b.SA, b.F); JP_LAUNCH(c);
				k_place_runs<<<(M + 255) / 256, 256, 0, s>>>(lb.k[rc], lb.v[rc], R, M, (u32)depth, rt, b.SA, b.F, b.ISA); JP_LAUNCH(c);
			JP_KCHECK();
			JP_CUDA(cudaMemcpyAsync(c.h_small + 21, rt.counters + 2, sizeof(u32), cudaMemcpyDeviceToHost, s));   // read after the grouping step's sync
		}
	}
	JP_CUDA(cudaMemsetAsync(b.ISA + n, 0, sizeof(u32), s));             // the empty suffix ranks below everything
	JP_TRY(group_step(c, b, nullptr, b.AP[0], (u32)n, s));
	JP_CUDA(cudaEventRecord(c.ev[3], s));
	act = 0;
	A = (u32)c.h_small[8]; G = (u32)c.h_small[9];
	i64 depth0 = depth;                                                 // symbols every initial key is known to cover
	if (ctx_keys) {
		depth0 = (i64)(u32)c.h_small[22];
		if (depth0 < (i64)ck_order || depth0 > (i64)n + 64) { set_error_detail("context-coded keys cover %lld symbols", (long long)depth0); return JP_ERR_INTERNAL; }
		if (depth0 > 64) depth0 = 64;
		if (depth0 < 1) depth0 = 1;
		h = depth0;
		st->initial_depth = (i32)depth0;
		if (trace_rounds) fprintf(stderr, "[jp_bwt keys] context-coded: order %u, %u symbols, %d key bits%s, every key covers >= %lld symbols; active after the sort %u\n", ck_order, ck_S - 1, ck_key_bits, ck_packed ? " (packed with the position)" : "", (long long)depth0, A);
	}

	// Periodic repeats: when most of the block is still unsorted, look for a dominant distance between group neighbours.
	const u32 shared_levels = bypass ? (u32)c.h_small[21] : 0;
	bool run_jumps = shared_levels >= 2048;                              // worth a second detection pass
	if (const char* e = getenv("JP_BWT_FWD_RUNJUMP")) run_jumps = atoi(e) != 0 && shared_levels > 0;
	if (!premode)
	{
		bool look = (u64)A * 4 >= (u64)n * 3 && n >= (1 << 16);
		if (const char* e = getenv("JP_BWT_FWD_PERIODIC")) look = atoi(e) != 0 && A >= 2 * PER_SAMPLE;
		if (look) {
			JP_TRY(arena2_reserve(c, (size_t)PER_MAXP * 4 + 256));          // (the run bypass tables, if any, are dead by now)
			u32* hist = (u32*)c.arena2.base; u32* ct = hist + PER_MAXP;
			JP_CUDA(cudaMemsetAsync(hist, 0, (size_t)PER_MAXP * 4 + 64, s));
			const u32 samples = (A + PER_SAMPLE - 1) / PER_SAMPLE;
			k_per_sample<<<(samples + 255) / 256, 256, 0, s>>>(b.AP[0], b.SA, A, hist); JP_LAUNCH(c);
			k_per_pick<<<1, 1024, 0, s>>>(hist, ct + 2); JP_LAUNCH(c);
			JP_KCHECK();
			JP_CUDA(cudaMemcpyAsync(c.h_small + 14, ct + 2, 2 * sizeof(u32), cudaMemcpyDeviceToHost, s));
			JP_CUDA(cudaStreamSynchronize(s));
			const u32 p = (u32)c.h_small[14], hits = (u32)c.h_small[15];
			h_a = depth0;
			while (h_a < (i64)p) h_a *= 2;
			if (p >= 1 && (u64)hits * PER_SAMPLE * 2 >= (u64)A && h_a * 4 <= (i64)n) {
				periodic = true;
				JP_TRY(per_tabs_alloc(c, n, ktiles, pt));                     // repeat lengths: 4.1 N, only now that they are needed
				JP_TRY(per_tabs_fill(c, d_T, n, ktiles, p, (u32)h_a, pt, s));
				ps.rl = pt.RL; ps.bits = pt.bits; ps.p = p;
				keep2 = pt.keep;
				rank_bits = bit_length(2 * (u64)n + 1);
				st->period = (i32)p;
			}
		}
	}
	if (!periodic && A > 0 && run_jumps) {
		// Run bypass left levels that several runs of equal symbol and class reach (padding to a fixed record size, say):
		// ordinary groups, but ones whose members differ only in what follows their runs. They are already uniform in
		// (class, run length) -- the placement did what pass A does -- so the jump keys of pass B apply from the first round on:
		// a run suffix with r >= h takes the rank of the suffix right after its run, and such a level is resolved as soon
		// as the followers differ instead of after log2(run length / depth) rounds.
		JP_TRY(per_tabs_alloc(c, n, ktiles, pt));                         // (the bypass tables are dead: everything is placed)
		JP_TRY(per_tabs_fill(c, d_T, n, ktiles, 1u, (u32)depth, pt, s));
		ps.rl = pt.RL; ps.bits = pt.bits; ps.p = 1;
		periodic = true; pass = 1; h_a = depth;
		keep2 = pt.keep;
		st->period = 1;
	}
	JP_TRY(run_rounds((i64)1 << 40));
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
	const int emit_log2 = getenv("JP_BWT_FWD_EMIT_REGION_LOG2") ? std::max(8, atoi(getenv("JP_BWT_FWD_EMIT_REGION_LOG2"))) : 26;   // (tests lower it)
	const u32 emit_regions = (u32)(((i64)nlen + (1 << emit_log2) - 1) >> emit_log2), emit_tiles = (u32)(((i64)nlen + 1023) / 1024);
	if (emit_regions >= 2 && emit_regions <= 8 && !getenv("JP_BWT_FWD_EMIT_ONE_SWEEP")) {
		JP_CUDA(cudaMemsetAsync(d_out, 0, (size_t)nlen, s));               // (the head flags that lived here are dead)
		k_fwd_emit_regions<<<emit_regions * emit_tiles, 256, 0, s>>>(d_in, b.SA, b.ISA, nlen, d_out, emit_log2, emit_tiles); JP_LAUNCH(c);
	} else { k_fwd_emit<<<(int)emit_tiles, 256, 0, s>>>(d_in, b.SA, b.ISA, nlen, d_out); JP_LAUNCH(c); }
	k_fwd_trailer<<<1, 128, 0, s>>>(d_in, b.ISA, nlen, len, d_out); JP_LAUNCH(c);
	JP_KCHECK();
	JP_CUDA(cudaEventRecord(c.ev[5], s));
	JP_CUDA(cudaStreamSynchronize(s));
	for (int i = 0; i < 5; i++) JP_CUDA(cudaEventElapsedTime(&st->ms_phase[i], c.ev[i], c.ev[i + 1]));
	JP_CUDA(cudaEventElapsedTime(&st->ms_total, c.ev[0], c.ev[5]));
	st->device_bytes = c.arena.high + c.arena2.high;
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
