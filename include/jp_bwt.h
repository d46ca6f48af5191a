/* jp_bwt.h -- C-ABI of the B200-native Jampack BWT stage (libjpbwt.so).
 *
 * This is the drop-in boundary for the reference's BWT stage. The reference has no FFI of its own
 * for this path: its callers are Jampack::Comp (jampack.cpp:40, Bwt->ForwardBwt(Input, Output)) and
 * Jampack::Decomp (jampack.cpp:50, Bwt->InverseBwt(Input, Output, Option)), on the plug API of
 * format.hpp:37-54 (Buffer{uchar* block; Index* size}, Options). The two entry points below carry
 * exactly what those calls carry -- plain pointers and sizes; jampack_b200/host/bwt_shim.cpp is the
 * ~30-line BlockSort::Bwt replacement that binds them.
 *
 * All arithmetic is integer/byte; results are bit-exact with the reference by construction (the
 * suffix array, hence the BWT and the sampled indices, is mathematically unique).
 *
 * There is NO CPU fallback: every entry point fails with JP_ERR_NO_DEVICE when no CUDA device is
 * usable. Entry points are re-entrant and may be called concurrently from many host threads
 * (the reference calls its stage from an OpenMP team, jampack.cpp:215-219 and :313-317); each call
 * borrows one (device, stream, workspace) context from an internal pool, whole blocks being
 * spread round-robin over the configured devices (no collective: blocks are independent).
 */
#ifndef JP_BWT_H
#define JP_BWT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JP_BWT_UNITS          120   /* format.hpp:26  BWT_UNITS: sampled primary indices per block   */
#define JP_BWT_TRAILER_BYTES  480   /* bwt.cpp:27     BWT_UNITS * sizeof(Index), Index = int32        */
#define JP_BWT_MAX_LEN        (1000 << 20) /* format.hpp:22 MAX_BLOCKSIZE; 1.05x buffers stay < 2^31  */
#define JP_BWT_MAX_ROUNDS     40

enum {
	JP_OK              =  0,
	JP_ERR_ARG         = -1,  /* null pointer, negative or oversize length                            */
	JP_ERR_NO_DEVICE   = -2,  /* no usable CUDA device (there is no CPU path)                          */
	JP_ERR_CUDA        = -3,  /* a CUDA runtime call or kernel failed; see jp_bwt_last_error_detail()  */
	JP_ERR_OOM         = -4,  /* device or pinned-host allocation failed                               */
	JP_ERR_BAD_INDEX   = -5,  /* inverse: a stored primary index is outside [1, nlen] or duplicated    */
	JP_ERR_CORRUPT     = -6,  /* inverse: BWT bytes and indices are inconsistent (chain length check)  */
	JP_ERR_INTERNAL    = -7   /* an internal invariant failed (reported, never silently ignored)       */
};

/* ---- the two stage entry points (HOST buffers) ---------------------------------------------------
 *
 * jp_bwt_forward  replaces BlockSort::Bwt::ForwardBwt (bwt.cpp:22-65) incl. divsufsort (divsufsort.cpp:1721).
 *   in[0..len)            the block as LZ77 left it (jampack.cpp:39)
 *   out[0..len)           BWT bytes of in[0..nlen) followed by the raw tail in[nlen..len), nlen = len - len%120
 *   out[len..len+480)     120 native-endian int32 sampled indices ISA[k*(nlen/120)]+1 -- written iff nlen > 0
 *                         (for len < 120 the reference leaves these bytes untouched, bwt.cpp:35; so do we)
 *   *out_len              len + 480, always (bwt.cpp:27)
 *   `out` must have room for len+480 bytes; `in` and `out` must not overlap; neither is retained.
 *
 * jp_bwt_inverse  replaces BlockSort::Bwt::InverseBwt (bwt.cpp:72-282); Options are not needed: the
 *   output does not depend on the unit/thread count (bwt.cpp:92-132 only shapes the CPU loop).
 *   in[0..len_with_trailer)  exactly what jp_bwt_forward produced; *out_len = len_with_trailer - 480.
 *   All 120 stored indices are used (every decode unit of the format runs concurrently) and, unlike the
 *   reference, validated: out-of-range indices give JP_ERR_BAD_INDEX instead of an out-of-bounds read.
 *
 * Host buffers may be pageable (the reference's calloc/realloc blocks, jampack.cpp:74-76,157-159) or
 * pinned (jp_bwt_host_alloc); pinned buffers are DMA'd directly.
 */
int jp_bwt_forward(const uint8_t* in, int32_t len, uint8_t* out, int32_t* out_len);
int jp_bwt_inverse(const uint8_t* in, int32_t len_with_trailer, uint8_t* out, int32_t* out_len);

/* ---- the same stage with the block already resident in HBM ---------------------------------------
 * d_in / d_out are device pointers on `device`, both 16-byte aligned. The work is enqueued on
 * `stream` (a cudaStream_t, or NULL for the context's own stream) and the call returns after that stream
 * has drained. The context's own stream is NON-BLOCKING: with stream == NULL the caller must have finished
 * producing d_in (and must not be writing d_out) before the call -- work queued on the legacy default stream
 * is not waited for. These are what the `value` leg of bench.py times, and what a device-resident
 * neighbour stage would call. */
int jp_bwt_forward_device(const uint8_t* d_in, int32_t len, uint8_t* d_out, int device, void* stream);
int jp_bwt_inverse_device(const uint8_t* d_in, int32_t len_with_trailer, uint8_t* d_out, int device, void* stream);
/* Same, but d_in is CONSUMED: once the LF table is built the input block is dead (the reference's caller swaps
 * streams and overwrites it, jampack.cpp:50-51), so its bytes hold the sub-chain records and the whole call stays
 * within in + out + 4N table = 6N of device memory (+ N/64 of histograms). The host entry point always works
 * this way on its own device copy. */
int jp_bwt_inverse_device_consume(uint8_t* d_in, int32_t len_with_trailer, uint8_t* d_out, int device, void* stream);

/* ---- the suffix sorter on its own ------------------------------------------------
 * jp_bwt_suffix_array replaces divsufsort(T, SA, n) (divsufsort.cpp:1721, contract divsufsort.hpp:37-45) for its
 * second call site, the -m2 suffix-array match finder of lz77.cpp:141: sa[0..n) receives the suffix array of
 * in[0..n), a proper prefix sorting before its extensions. jampack_b200/host/divsufsort_shim.cpp binds it under the
 * reference's own symbol name. n == 0 is a no-op. Returns JP_OK or an error code. */
int jp_bwt_suffix_array(const uint8_t* in, int32_t n, int32_t* sa);

/* ---- second stage, first half ------------------------------------------------------
 * Sorted rank coding + RLE0 of a stage block, in the 1 MiB chunks of Ans::Encode (ans.cpp:134-160, ans.hpp:33): what
 * Postcoder::Encode (rank.cpp:45-90) and RLE::encode (rle.cpp:22-47) compute per chunk, for every chunk at once. The
 * adaptive rANS that follows (ans.cpp:162-221) is not part of it: a host (or a later kernel) consumes
 *   freq[256 * k ..]            the 256 byte frequencies of chunk k (what WriteHeader stores, ans.cpp:282-286)
 *   rle[JP_ANS_CHUNK * k ..]    the 16-bit RLE0 symbols of chunk k
 *   rlen[k]                     their number
 * for k < ceil(len / JP_ANS_CHUNK). The device form takes the block where jp_bwt_forward_device left it (d_in = that
 * call's d_out, len = its len + 480): the BWT never leaves HBM between the two stages. */
#define JP_ANS_CHUNK (1 << 20)
int jp_src_rle0_device(const uint8_t* d_in, int32_t len, int32_t* d_freq, uint16_t* d_rle, int32_t* d_rlen, int device, void* stream);
int jp_src_rle0(const uint8_t* in, int32_t len, int32_t* freq, uint16_t* rle, int32_t* rlen);

/* ---- device selection (block sharding) ---------------------------------------------
 * Default: every visible device, or the list in the environment variable JP_BWT_DEVICES ("0,1,2").
 * jp_bwt_set_devices replaces the list (n = 0 restores the default). Returns JP_OK or an error. */
int jp_bwt_set_devices(const int* ids, int n);
int jp_bwt_device_count(void);

/* Starts creating the CUDA contexts of all configured devices in the background and returns at once: a host that
 * knows it will call the stage (the shim's static initialiser does) hides the ~1 s per device behind its own start-up. */
int jp_bwt_warmup_async(void);

/* Pinned host blocks for callers that want direct DMA. Pageable blocks need no action: the host
 * entry points page-lock a caller's block in the background once a first call on it has completed (the reference re-uses
 * two blocks per Jampack instance for the whole run, jampack.cpp:74-76, :157-159), so every later call DMAs directly;
 * JP_BWT_HOST_REGISTER=0 disables that. */
void* jp_bwt_host_alloc(uint64_t bytes);
void  jp_bwt_host_free(void* p);
/* A caller that frees (or reallocates) a pageable block it has passed to jp_bwt_forward / jp_bwt_inverse while the
 * library lives on tells it so BEFORE the free: the page-lock taken after its first call is dropped. p == NULL drops all.
 * (The reference keeps its blocks for the whole run and needs no call.) */
void  jp_bwt_host_release(const void* p);

/* ---- diagnostics ----------------------------------------------------------------------------------*/
typedef struct jp_bwt_stats {
	int32_t  direction;          /* 0 forward, 1 inverse, 2 sorted rank coding + RLE0                  */
	int32_t  len, nlen;
	int32_t  device;
	int32_t  kernel_launches;    /* kernels this call launched                                         */
	int32_t  rounds;             /* forward: prefix-doubling rounds after the initial radix bucketing  */
	int32_t  symbol_bits;        /* forward: bits per symbol in the initial key (context-coded keys: the code's mean rate, rounded) */
	int32_t  initial_depth;      /* forward: symbols EVERY initial key covers = the h the doubling starts from */
	int32_t  subchains;          /* inverse: sub-chains the 120 decode units were split into           */
	int32_t  subchain_spacing;   /* inverse: mean sub-chain length (marker spacing m)                  */
	uint64_t device_bytes;       /* workspace bytes held for this call (excl. caller's in/out)         */
	uint64_t random_sectors;     /* counted 32 B random sector touches (SURVEY.md 8d model)            */
	float    ms_total;           /* device time of the whole call (CUDA events)                        */
	float    ms_h2d, ms_d2h;     /* host entry points only                                             */
	float    ms_phase[8];        /* forward: 0 setup + keys 1 initial sort 2 initial ranks 3 rounds 4 emit */
	                             /* inverse: 0 prepare+histogram 1 LF build 2 length walk 3 ranking 4 emit walk */
	                             /*          (single-walk path: 2 decode walk 3 ranking 4 placement + copy)     */
	float    active_fraction[JP_BWT_MAX_ROUNDS]; /* forward: a_r = suffixes still unsorted entering round r */
	int32_t  stream_chunks;      /* inverse: 1 KiB stream chunks the single-walk path used; 0 = two-pass path,  */
	                             /* negative = the stream space overflowed and the two-pass path was rerun      */
	float    large_fraction;     /* forward: share of the block that went through the large-group route, summed */
	                             /* over the rounds                                                              */
	int32_t  radix_tiles;        /* forward: tiles that took the shared-memory radix route, summed over rounds   */
	int32_t  bypass_suffixes;    /* forward: suffixes inside single-symbol runs, placed without sorting          */
	int32_t  bypass_runs;        /* forward: the runs they belong to                                              */
	int32_t  period;             /* forward: period of the repeats ordered by length instead of by doubling; 0 = none */
} jp_bwt_stats;

/* Stats of the last call made by the calling thread. */
int jp_bwt_last_stats(jp_bwt_stats* out);

const char* jp_bwt_strerror(int rc);
/* Detail text (CUDA error string, failing check) of the last failure on the calling thread. */
const char* jp_bwt_last_error_detail(void);
const char* jp_bwt_version(void);

/* ---- test hooks (used by tests/ only; stable enough to script against) ----------------------------
 * jp_bwt_debug_lf: builds the inverse's LF table for in[0..nlen) on device 0 and copies it back
 *   (marker bits stripped): lf[i] = 1 + C[in[i]] + #{j < i : in[j] == in[i]}, ctable[c] = #{in[j] < c}.
 *   This is the inverse permutation of the reference's Map (bwt.cpp:171-174): Map[lf[i]-1] == i + (i >= idx).
 * jp_bwt_debug_gather_rate: random 4-byte gather micro-benchmark over a table of `table_bytes`
 *   (`chains` dependent walkers, `steps` each); returns sectors/s, or a negative error code. */
int jp_bwt_debug_lf(const uint8_t* in, int32_t nlen, int32_t* lf, int32_t* ctable /*[257]*/);
/* jp_bwt_debug_copy: only the host->device and device->host copies a stage call of these sizes makes (no kernels):
 *   the ceiling bench.py reports its end-to-end figure against. */
int jp_bwt_debug_copy(const uint8_t* in, int32_t in_len, uint8_t* out, int32_t out_len);
double jp_bwt_debug_gather_rate(uint64_t table_bytes, int32_t chains, int32_t steps, int dependent);

#ifdef __cplusplus
}
#endif
#endif /* JP_BWT_H */
