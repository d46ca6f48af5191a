/* Synthetic block generators (BASELINE.json configs) and the FNV-1a hash used to pin known-answer
 * values. Definitions follow SURVEY.md Appendix B exactly (the reference ships no data of its own).
 * Input generation only: neither the product path nor the oracle's algorithm lives here. */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>

static inline uint64_t sm64_next(uint64_t* x)
{
	*x += 0x9E3779B97F4A7C15ull;
	uint64_t z = *x;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

void jps_gen_uniform(uint8_t* T, int64_t n, uint64_t seed)
{
	uint64_t x = seed;
	for (int64_t i = 0; i < n; i++) T[i] = (uint8_t)(sm64_next(&x) >> 56);
}

/* order-2 Markov text over a 64-symbol printable alphabet, geometric choice among 8 successors */
void jps_gen_markov2(uint8_t* T, int64_t n, uint64_t seed)
{
	uint64_t x = seed;
	uint8_t* nxt = (uint8_t*)malloc(64 * 64 * 8);
	for (int i = 0; i < 64 * 64 * 8; i++) nxt[i] = (uint8_t)(sm64_next(&x) >> 58);
	unsigned a = 0, b = 1;
	for (int64_t i = 0; i < n; i++) {
		uint64_t u = sm64_next(&x);
		unsigned k = (unsigned)__builtin_ctzll((u >> 32) | 0x80);
		unsigned s = nxt[((a << 6) | b) * 8 + k];
		T[i] = (uint8_t)(32 + s);
		a = b; b = s;
	}
	free(nxt);
}

/* period-1021 motif over {a,b,c,d}, one flipped bit every 64 KiB */
void jps_gen_repetitive(uint8_t* T, int64_t n, uint64_t seed)
{
	uint64_t x = seed;
	uint8_t m[1021];
	for (int i = 0; i < 1021; i++) m[i] = (uint8_t)('a' + (sm64_next(&x) >> 62));
	for (int64_t i = 0; i < n; i++) T[i] = m[i % 1021];
	for (int64_t i = 65536; i < n; i += 65536) T[i] ^= 1;
}

void jps_gen_alla(uint8_t* T, int64_t n)
{
	for (int64_t i = 0; i < n; i++) T[i] = 'a';
}

/* KAT-A/B/C: T[i] = (i*i + 3i) & 255 ; KAT-E: i%3==0 ? 0 : i%5==0 ? 255 : i&1 */
void jps_gen_kat_quadratic(uint8_t* T, int64_t n)
{
	for (int64_t i = 0; i < n; i++) T[i] = (uint8_t)((i * i + 3 * i) & 255);
}
void jps_gen_kat_extremes(uint8_t* T, int64_t n)
{
	for (int64_t i = 0; i < n; i++) T[i] = (i % 3 == 0) ? 0 : (i % 5 == 0) ? 255 : (uint8_t)(i & 1);
}

uint64_t jps_fnv1a64(const uint8_t* p, int64_t n)
{
	uint64_t h = 0xcbf29ce484222325ull;
	for (int64_t i = 0; i < n; i++) { h ^= p[i]; h *= 0x100000001b3ull; }
	return h;
}
