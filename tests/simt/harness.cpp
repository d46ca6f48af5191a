// harness.cpp -- TEST INFRASTRUCTURE ONLY: the few host-side pieces the device drivers expect from jp_bwt_api.cu
// (error detail, arena, error mapping), restated for the emulator, and C entry points for the tests.
#include "bwt_internal.cuh"
#include <stdarg.h>

namespace jp {

static char g_detail[512];
void set_error_detail(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_detail, sizeof(g_detail), fmt, ap); va_end(ap); }

int map_dev_err(int de)
{
	switch (de) {
	case DE_NONE: return JP_OK;
	case DE_BAD_INDEX: return JP_ERR_BAD_INDEX;
	case DE_CHAIN_LEN: case DE_CHAIN_RANGE: case DE_RANK_LOOP: return JP_ERR_CORRUPT;
	default: return JP_ERR_INTERNAL;
	}
}

int arena_reserve(Ctx& c, size_t total)
{
	if (c.arena.cap - c.arena.off >= total) return JP_OK;
	if (c.arena.off != 0) return JP_ERR_INTERNAL;
	free(c.arena.base);
	c.arena.cap = total + (1u << 20);
	c.arena.base = (u8*)aligned_alloc(256, (c.arena.cap + 255) & ~(size_t)255);
	memset(c.arena.base, 0xA5, c.arena.cap);        // device memory is not zeroed either
	return c.arena.base ? JP_OK : JP_ERR_OOM;
}

int arena2_reserve(Ctx& c, size_t total, size_t keep)
{
	if (total > c.arena2.high) c.arena2.high = total;
	if (c.arena2.cap >= total) return JP_OK;
	const size_t cap = total + 4096;
	u8* fresh = (u8*)aligned_alloc(256, (cap + 255) & ~(size_t)255);
	if (!fresh) return JP_ERR_OOM;
	memset(fresh, 0xA5, cap);
	if (c.arena2.base && keep) memcpy(fresh, c.arena2.base, keep);
	free(c.arena2.base);
	c.arena2.base = fresh; c.arena2.cap = cap;
	return JP_OK;
}

static Ctx& ctx()
{
	static Ctx c;
	static int small[256];
	c.device = 0; c.h_small = small; c.sm_count = 2; c.launches = 0;
	c.arena.reset(); c.arena.high = 0;
	return c;
}

} // namespace jp

extern "C" {

// in: BWT || tail || trailer (len_with_trailer bytes). consume != 0: the input copy may be used as scratch (6N layout).
int emu_inverse(const uint8_t* in, int32_t len_with_trailer, uint8_t* out, int consume, int32_t* stream_chunks, int32_t* launches)
{
	jp::Ctx& c = jp::ctx();
	const size_t n = (size_t)len_with_trailer;
	uint8_t* d_in = (uint8_t*)aligned_alloc(256, (n + 511) & ~(size_t)255);
	uint8_t* d_out = (uint8_t*)aligned_alloc(256, (n + 511) & ~(size_t)255);
	memcpy(d_in, in, n); memset(d_out, 0x5C, n);
	jp_bwt_stats st; memset(&st, 0, sizeof(st));
	const int rc = jp::inverse_device(c, d_in, len_with_trailer, d_out, nullptr, &st, consume ? d_in : nullptr);
	if (len_with_trailer >= JP_BWT_TRAILER_BYTES) memcpy(out, d_out, n - JP_BWT_TRAILER_BYTES);
	if (stream_chunks) *stream_chunks = st.stream_chunks;
	if (launches) *launches = c.launches;
	free(d_in); free(d_out);
	return rc;
}

int emu_suffix_array(const uint8_t* in, int32_t n, int32_t* sa)
{
	jp::Ctx& c = jp::ctx();
	return jp::debug_suffix_array(c, in, n, sa);
}

/* sorted rank coding + RLE0 per 1 MiB chunk: freq[chunks][256], rle (chunk k at k << 20), rlen[chunks] */
int emu_src_rle0(const uint8_t* in, int32_t len, int32_t* freq, uint16_t* rle, int32_t* rlen)
{
	jp::Ctx& c = jp::ctx();
	const size_t n = (size_t)len;
	uint8_t* d_in = (uint8_t*)aligned_alloc(256, (n + 511) & ~(size_t)255);
	memcpy(d_in, in, n);
	jp_bwt_stats st; memset(&st, 0, sizeof(st));
	const int rc = jp::src_rle0_device(c, d_in, len, freq, rle, rlen, nullptr, &st);
	free(d_in);
	return rc;
}

int emu_forward(const uint8_t* in, int32_t len, uint8_t* out, int32_t* rounds, int32_t* launches)
{
	jp::Ctx& c = jp::ctx();
	const size_t n = (size_t)len;
	uint8_t* d_in = (uint8_t*)aligned_alloc(256, (n + 511) & ~(size_t)255);
	uint8_t* d_out = (uint8_t*)aligned_alloc(256, (n + JP_BWT_TRAILER_BYTES + 511) & ~(size_t)255);
	memcpy(d_in, in, n); memset(d_out, 0x5C, n + JP_BWT_TRAILER_BYTES);
	jp_bwt_stats st; memset(&st, 0, sizeof(st));
	const int rc = jp::forward_device(c, d_in, len, d_out, nullptr, &st);
	memcpy(out, d_out, n + JP_BWT_TRAILER_BYTES);
	if (rounds) *rounds = st.rounds;
	if (launches) *launches = c.launches;
	free(d_in); free(d_out);
	return rc;
}

}
