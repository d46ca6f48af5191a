// bwt_internal.cuh -- per-call context (device, stream, workspace arena) shared by the forward and
// inverse drivers and the C-ABI in jp_bwt_api.cu.
#pragma once
#include "common.cuh"
#include "../../include/jp_bwt.h"

namespace jp {

// Longest block an entry point accepts: the reference's buffers are 1.05 * MAX_BLOCKSIZE (jampack.cpp:74-76), but the
// inverse keeps row numbers in 30 bits of an LF entry (bits 30/31 are the anchor and mark flags, bwt_inverse.cu), so a
// block -- in either direction: what the forward emits must be invertible -- stops two rows short of 2^30.
constexpr i64 JP_BWT_MAX_CALL_LEN = ((i64)JP_BWT_MAX_LEN * 105 / 100) < (((i64)1 << 30) - 2) ? ((i64)JP_BWT_MAX_LEN * 105 / 100) : (((i64)1 << 30) - 2);

// Grow-only bump arena: one cudaMalloc per context, re-used by every call that borrows the context
// (cudaMalloc of gigabytes costs milliseconds; a block stage is called once per block).
struct Arena {
	u8*    base = nullptr;
	size_t cap = 0, off = 0, high = 0;
	void reset() { off = 0; }
	static size_t align(size_t x) { return (x + 255) & ~(size_t)255; }
};

struct Ctx {
	int          device = -1;
	cudaStream_t own_stream = nullptr;
	Arena        arena;
	Arena        arena2;              // second, lazily grown allocation: scratch only unusual blocks need (forward: sorting over-long groups)
	int*         h_small = nullptr;   // pinned: error flag + counters read back per round ([0..64)), period probe samples ([64..256))
	u8*          h_stage[2] = {nullptr, nullptr}; // pinned staging for pageable host blocks
	size_t       h_stage_cap = 0;
	u8*          d_in = nullptr;      // device copies of the caller's host blocks (host entry points)
	u8*          d_out = nullptr;
	size_t       d_io_cap = 0;
	cudaEvent_t  ev[12] = {};
	int          launches = 0;
	u32          cur_n = 0;           // block length of the call in flight
	int          sm_count = 0;
	bool         busy = false;
};

// Reserve `total` bytes up front (re-allocating the arena if it is too small), then carve.
int arena_reserve(Ctx& c, size_t total);
// The second arena is used whole (no carving): at least `total` bytes at c.arena2.base afterwards. If it has to grow,
// its first `keep` bytes are carried over (the base pointer changes).
int arena2_reserve(Ctx& c, size_t total, size_t keep = 0);
template <typename T> inline T* arena_take(Ctx& c, size_t count)
{
	size_t bytes = Arena::align(count * sizeof(T));
	T* p = (T*)(c.arena.base + c.arena.off);
	c.arena.off += bytes;
	if (c.arena.off > c.arena.high) c.arena.high = c.arena.off;
	return p;
}

#define JP_LAUNCH(ctx) (++(ctx).launches)
#define JP_KCHECK() JP_CUDA(cudaGetLastError())

// scratch_in: nullptr, or d_in itself when the caller allows the input block to be overwritten (keeps the call at 6N)
int inverse_device(Ctx& c, const u8* d_in, i32 len_with_trailer, u8* d_out, cudaStream_t s, jp_bwt_stats* st, u8* scratch_in);
int forward_device(Ctx& c, const u8* d_in, i32 len, u8* d_out, cudaStream_t s, jp_bwt_stats* st);
int src_rle0_device(Ctx& c, const u8* d_in, i32 len, i32* d_freq, u16* d_rle, i32* d_rlen, cudaStream_t s, jp_bwt_stats* st);

// test hooks
int debug_lf(Ctx& c, const u8* h_in, i32 nlen, i32* h_lf, i32* h_ctable);
int debug_suffix_array(Ctx& c, const u8* h_in, i32 n, i32* h_sa);
double debug_gather_rate(Ctx& c, u64 table_bytes, i32 chains, i32 steps, int dependent);

int map_dev_err(int de);

} // namespace jp
