/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the Jampack BWT stage.
 *
 * A plain-C restatement of the reference's algorithm for the hot path. It exists to CHECK the
 * CUDA path (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg); nothing in the
 * product (jampack_b200/, include/) may import, link or call it.
 *
 * Parity pin: the reference ships no tests or golden vectors for this path,
 * so this file is pinned against the reference ITSELF, compiled unmodified into oracle/_ref
 * (oracle/Makefile) -- tests/test_oracle.py compares the two on every generator and edge case, and
 * tests/golden/kat.json holds the known-answer hashes produced by the reference (script:
 * tests/golden/make_golden.py).
 *
 * What follows what (all citations into /root/reference):
 *   jpo_bwt_forward  : bwt.cpp:22-65  (sizes :24-30, tail copy :32-33, index sampling :44-48,
 *                      BWT emission :50-56, trailer :57-61). The suffix array itself comes from
 *                      divsufsort.cpp:1721 in the reference; only its CONTRACT is restated here
 *                      (divsufsort.hpp:37-45: plain suffix array, a proper prefix sorts first) with an
 *                      independent Manber-Myers style prefix-doubling sorter -- the result is unique.
 *   jpo_build_map    : bwt.cpp:141-174 (histogram, prefix sum, stable scatter building Map).
 *   jpo_bwt_inverse  : bwt.cpp:72-89 (sizes, tail, indices), :176-183 (chain seeds),
 *                      :261-275 (the walk).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define JPO_UNITS 120 /* format.hpp:26 BWT_UNITS */

/* ---- suffix array: contract of divsufsort(T, SA, n) ------------------------------------------
 * Doubling invariant: before the round with depth h, SA is ordered by the first h symbols of each
 * suffix (end of string smaller than every byte) and rk[i] = index in SA of the first member of
 * i's group. One round = a stable counting sort by rk[i] of the sequence "suffixes ordered by
 * rk[i+h]", which is just SA shifted left by h, preceded by the suffixes that end within h. */
static int jpo_suffix_array(const uint8_t* T, int32_t* SA, int32_t n)
{
	if (n <= 0) return 0;
	int32_t* rk  = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
	int32_t* nrk = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
	int32_t* ptr = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
	int32_t* tmp = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
	if (!rk || !nrk || !ptr || !tmp) { free(rk); free(nrk); free(ptr); free(tmp); return -1; }

	int32_t start[257] = {0}, fill[256];
	for (int32_t i = 0; i < n; i++) start[T[i] + 1]++;
	for (int c = 0; c < 256; c++) start[c + 1] += start[c];
	for (int c = 0; c < 256; c++) fill[c] = start[c];
	for (int32_t i = 0; i < n; i++) SA[fill[T[i]]++] = i;
	int32_t groups = 0;
	for (int c = 0; c < 256; c++) if (start[c + 1] > start[c]) groups++;
	for (int32_t i = 0; i < n; i++) rk[i] = start[T[i]];

	for (int64_t h = 1; groups < n; h *= 2) {
		for (int32_t j = 0; j < n; j++) ptr[j] = j;
		for (int64_t i = (n - h > 0 ? n - h : 0); i < n; i++) tmp[ptr[rk[i]]++] = (int32_t)i;
		for (int32_t j = 0; j < n; j++) {
			int64_t s = (int64_t)SA[j] - h;
			if (s >= 0) tmp[ptr[rk[s]]++] = (int32_t)s;
		}
		int32_t g = 0;
		nrk[tmp[0]] = 0; groups = 1;
		for (int32_t j = 1; j < n; j++) {
			int32_t a = tmp[j - 1], b = tmp[j];
			int32_t ka = (a + h < n) ? rk[a + h] : -1;
			int32_t kb = (b + h < n) ? rk[b + h] : -1;
			if (rk[a] != rk[b] || ka != kb) { g = j; groups++; }
			nrk[b] = g;
		}
		memcpy(SA, tmp, sizeof(int32_t) * (size_t)n);
		int32_t* t = rk; rk = nrk; nrk = t;
	}
	free(rk); free(nrk); free(ptr); free(tmp);
	return 0;
}

int jpo_suffix_array_export(const uint8_t* T, int32_t* SA, int32_t n) { return jpo_suffix_array(T, SA, n); }

/* ---- forward: bwt.cpp:22-65 -------------------------------------------------------------------- */
int jpo_bwt_forward(const uint8_t* T, int32_t Len, uint8_t* Bwt, int32_t* out_len)
{
	*out_len = Len + JPO_UNITS * (int32_t)sizeof(int32_t);          /* :27 */
	int32_t remainder = Len % JPO_UNITS;                             /* :29 */
	int32_t nlen = Len - remainder;                                  /* :30 */
	for (int32_t i = 0; i < remainder; i++) Bwt[nlen + i] = T[nlen + i]; /* :32-33 */
	if (nlen <= 0) return 0;                                         /* :35 -- trailer left untouched */

	int32_t Ind[JPO_UNITS] = {0};
	int32_t* SA = (int32_t*)calloc((size_t)nlen, sizeof(int32_t));
	if (!SA) return -1;
	if (jpo_suffix_array(T, SA, nlen) != 0) { free(SA); return -2; }

	int32_t step = nlen / JPO_UNITS;                                 /* :44 */
	for (int32_t i = 0; i < nlen; i++)                               /* :46-48 */
		if (SA[i] % step == 0) Ind[SA[i] / step] = i;

	Bwt[0] = T[nlen - 1];                                            /* :50 */
	int32_t idx = Ind[0];
	for (int32_t i = 0; i < idx; i++)        Bwt[i + 1] = T[(SA[i] - 1) % nlen]; /* :53-54 */
	for (int32_t i = idx + 1; i < nlen; i++) Bwt[i]     = T[(SA[i] - 1) % nlen]; /* :55-56 */
	for (int k = 0; k < JPO_UNITS; k++) {                            /* :57-61 */
		int32_t v = Ind[k] + 1;
		memcpy(&Bwt[Len + 4 * k], &v, 4);
	}
	free(SA);
	return 0;
}

/* ---- Map (the reference's "LF/rank table"): bwt.cpp:141-174 ------------------------------------ */
int jpo_build_map(const uint8_t* Bwt, int32_t nlen, int32_t idx, int32_t* Map, int32_t* C /*[257]*/)
{
	int32_t count[257] = {0};
	for (int32_t i = 0; i < nlen; i++) count[Bwt[i] + 1]++;          /* :141-167, unrolling dropped */
	for (int i = 1; i < 256; i++) count[i] += count[i - 1];          /* :168-169 => count[c] = #bytes < c */
	if (C) { for (int i = 0; i < 256; i++) C[i] = count[i]; C[256] = nlen; }
	for (int32_t i = 0; i < idx; i++)    Map[count[Bwt[i]]++] = i;     /* :171-172 */
	for (int32_t i = idx; i < nlen; i++) Map[count[Bwt[i]]++] = i + 1; /* :173-174 */
	return 0;
}

/* ---- inverse: bwt.cpp:72-282 ------------------------------------------------------------------- */
int jpo_bwt_inverse(const uint8_t* Bwt, int32_t len_with_trailer, uint8_t* T, int32_t* out_len, int n_units)
{
	int32_t Len = len_with_trailer - JPO_UNITS * (int32_t)sizeof(int32_t); /* :77 */
	if (Len < 0) return -3;
	*out_len = Len;                                                  /* :78 */
	int32_t remainder = Len % JPO_UNITS, nlen = Len - remainder;     /* :80-81 */
	for (int32_t i = 0; i < remainder; i++) T[nlen + i] = Bwt[nlen + i]; /* :82-83 */
	if (nlen <= 0) return 0;
	if (n_units <= 0 || JPO_UNITS % n_units != 0) return -4;         /* :116-132 only ever yields divisors */

	int32_t Ind[JPO_UNITS];
	for (int k = 0; k < JPO_UNITS; k++) memcpy(&Ind[k], &Bwt[Len + 4 * k], 4); /* :87-89 */
	for (int k = 0; k < JPO_UNITS; k++) if (Ind[k] < 1 || Ind[k] > nlen) return -5; /* reference reads OOB instead */

	int32_t* Map = (int32_t*)malloc(sizeof(int32_t) * (size_t)nlen); /* :135 */
	if (!Map) return -1;
	int32_t idx = Ind[0];                                            /* :139 */
	jpo_build_map(Bwt, nlen, idx, Map, 0);

	int32_t step = nlen / n_units;                                   /* :176 */
	for (int j = 0; j < n_units; j++) {                              /* chains are independent */
		int32_t p = Ind[JPO_UNITS / n_units * j];                    /* :180-181 */
		int32_t off = step * j;                                      /* :182-183 */
		for (int32_t i = 0; i != step; i++) {                        /* :265-273 */
			p = Map[p - 1];
			T[i + off] = Bwt[p - (p >= idx)];
		}
	}
	free(Map);
	return 0;
}

/* Naive check used by the tests on tiny inputs: is SA the suffix array of T (prefix sorts first)? */
int jpo_check_suffix_array(const uint8_t* T, const int32_t* SA, int32_t n)
{
	for (int32_t i = 1; i < n; i++) {
		int32_t a = SA[i - 1], b = SA[i];
		int32_t la = n - a, lb = n - b, l = la < lb ? la : lb;
		int c = memcmp(T + a, T + b, (size_t)l);
		if (c > 0 || (c == 0 && la >= lb)) return i;
	}
	return 0;
}

/* ---- second stage, first half: sorted rank coding + RLE0, per 1 MiB chunk -----------------------------------------
 * Restates what Ans::Encode does to a block before the entropy coder (reference ans.cpp:134-160): for every chunk of
 * StackSize = 1 MiB (ans.hpp:33), Postcoder::Encode (rank.cpp:45-90) then RLE::encode (rle.cpp:22-47).
 * freq[256 * k ..], rle[JPO_CHUNK * k ..] and rlen[k] receive chunk k's results. Returns the number of chunks. */
#define JPO_CHUNK (1 << 20)
int jpo_src_rle0(const uint8_t* T, int32_t len, int32_t* freq, uint16_t* rle, int32_t* rlen)
{
	int chunks = 0;
	uint8_t* ranks = (uint8_t*)malloc(JPO_CHUNK);
	if (!ranks) return -1;
	for (int32_t in_p = 0; in_p < len; in_p += JPO_CHUNK, chunks++) {   /* ans.cpp:149-160 */
		const uint8_t* C = T + in_p;
		int32_t n = (in_p + JPO_CHUNK < len) ? JPO_CHUNK : (len - in_p);
		int32_t* F = freq + 256 * chunks;
		uint8_t S2R[256], R2S[256];
		int32_t bucket[256];
		int unique = 0;
		memset(F, 0, 256 * sizeof(int32_t));
		for (int32_t i = 0; i < n; i++) {                             /* rank.cpp:55-64: counts, list in order of first appearance */
			uint8_t s = C[i];
			if (F[s] == 0) { R2S[unique] = s; S2R[s] = (uint8_t)unique; unique++; }
			F[s]++;
		}
		{                                                             /* rank.cpp:15-38, :66-72: buckets by descending count */
			int32_t copy[256], pos = 0;
			memcpy(copy, F, sizeof(copy));
			for (int j = 0; j < 256; j++) {
				int32_t max = 0; int best = 0;
				for (int i = 0; i < 256; i++) if (copy[i] > max) { best = i; max = copy[i]; }
				if (max == 0) break;
				bucket[best] = pos; pos += F[best]; copy[best] = 0;
			}
		}
		for (int32_t i = 0; i < n; i++) {                             /* rank.cpp:74-87: move to front, rank stored in the symbol's bucket */
			uint8_t s = C[i], r = S2R[s];
			ranks[bucket[s]++] = r;
			if (r > 0) {
				do { R2S[r] = R2S[r - 1]; S2R[R2S[r]] = r; } while (0 < --r);
				R2S[0] = s; S2R[s] = 0;
			}
		}
		{                                                             /* rle.cpp:22-47 */
			uint16_t* out = rle + (size_t)JPO_CHUNK * chunks;
			int32_t o = 0;
			for (int32_t i = 0; i < n;) {
				if (ranks[i] == 0) {
					int32_t run = 1;
					while (i + run < n && ranks[i + run] == 0) run++;
					i += run;
					int32_t L = run + 1, msb = 0;
					for (int32_t v = L; v; v >>= 1) msb++;
					msb -= 1;
					while (msb--) out[o++] = (uint16_t)((L >> msb) & 1);
				} else out[o++] = (uint16_t)(ranks[i++] + 1);
			}
			rlen[chunks] = o;
		}
	}
	free(ranks);
	return chunks;
}
