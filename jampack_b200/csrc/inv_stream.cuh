// inv_stream.cuh -- stream layout and packed sub-chain records of the single-walk inverse, used by bwt_inverse.cu.
// Included INSIDE namespace jp (it opens none), after common.cuh.
#pragma once
constexpr u32 INV_REC_INVALID = 0xffffffffu;
constexpr u32 INV_WALK_BATCH  = 128;
constexpr int ST_ROWS  = 32;
constexpr int ST_CHUNK = ST_ROWS * 32;           // bytes per stream chunk
constexpr int ST_AHEAD = 8;                      // rows before a chunk fills at which the next one is requested
constexpr int PR_DIST_BITS = 24, PR_NXT_BITS = 26;
constexpr u32 PR_DIST_MASK = (1u << PR_DIST_BITS) - 1, PR_NXT_MASK = (1u << PR_NXT_BITS) - 1;
constexpr u32 PR_NXT_INVALID = PR_NXT_MASK;
constexpr u32 PR_LEN_MAX = (1u << (64 - PR_DIST_BITS - PR_NXT_BITS)) - 1;
constexpr u32 ST_NONE = 0xffffffffu;

__device__ __forceinline__ u64 pack3(u32 len, u32 nxt, u32 dist) { return ((u64)len << (PR_DIST_BITS + PR_NXT_BITS)) | ((u64)nxt << PR_DIST_BITS) | dist; }
__device__ __forceinline__ u32 pr_len(u64 r)  { return (u32)(r >> (PR_DIST_BITS + PR_NXT_BITS)); }
__device__ __forceinline__ u32 pr_nxt(u64 r)  { return (u32)(r >> PR_DIST_BITS) & PR_NXT_MASK; }
__device__ __forceinline__ u32 pr_dist(u64 r) { return (u32)r & PR_DIST_MASK; }

struct StreamSpace {
	u8* base0; u32 cap0;           // chunks [0, cap0) live here (the output block) ...
	u8* base1; u32 cap1;           // ... chunks [cap0, cap0 + cap1) here (free part of the consumed input, or workspace)
	u32* chunk_head;               // [walker warp] first chunk of the warp's stream
	u32* chunk_next;               // [chunk] the chunk that follows in the same stream
	u32* batch_head;               // [walker warp] first ticket batch the warp drew (batch = base / WALK_BATCH)
	u32* batch_next;               // [batch] the batch the same warp drew next
	u32  batch_cap;
};
__device__ __forceinline__ u8* chunk_ptr(const StreamSpace& sp, u32 c)
{
	return c < sp.cap0 ? sp.base0 + (size_t)c * ST_CHUNK : sp.base1 + (size_t)(c - sp.cap0) * ST_CHUNK;
}

