/* TEST INFRASTRUCTURE ONLY -- not part of the product path.
 *
 * extern "C" wrapper around the UNMODIFIED reference BWT stage, compiled from the
 * sources where they lie under /root/reference (see oracle/Makefile). The output
 * (oracle/_ref/libjamref.so) is git-ignored; no reference source is copied here.
 *
 * Wrapped: BlockSort::Bwt::ForwardBwt (bwt.cpp:22-65), InverseBwt (bwt.cpp:72-282),
 * through the reference's own Buffer/Options plug API (format.hpp:37-54).
 */
#include "bwt.hpp"
#include <chrono>
#include <omp.h>

extern "C" {

int ref_bwt_forward(const unsigned char* in, int len, unsigned char* out, int* out_len)
{
	int isz = len, osz = 0;
	Buffer I; I.block = const_cast<unsigned char*>(in); I.size = &isz;
	Buffer O; O.block = out; O.size = &osz;
	BlockSort::Bwt b;
	b.ForwardBwt(I, O);
	*out_len = osz;
	return 0;
}

/* `in` is not modified; *in_len_after receives the mutated *Input.size (bwt.cpp:77). */
int ref_bwt_inverse(const unsigned char* in, int len_with_trailer, unsigned char* out, int* out_len,
                    int threads, int* in_len_after)
{
	int isz = len_with_trailer, osz = 0;
	Buffer I; I.block = const_cast<unsigned char*>(in); I.size = &isz;
	Buffer O; O.block = out; O.size = &osz;
	Options opt; memset(&opt, 0, sizeof(opt));
	opt.Threads = threads; opt.Gpu = false; opt.Multiblock = true;
	BlockSort::Bwt b;
	b.InverseBwt(I, O, opt);
	*out_len = osz;
	if (in_len_after) *in_len_after = isz;
	return 0;
}

/* Block-parallel timing shape of Jampack::Compress/Decompress (jampack.cpp:215-219, :313-317):
 * `nblocks` independent blocks, one OpenMP worker each (nested regions stay serial).
 * Returns wall seconds for the whole batch. ins/outs are arrays of block pointers. */
double ref_bwt_forward_batch(const unsigned char* const* ins, const int* lens, unsigned char* const* outs,
                             int nblocks, int threads)
{
	auto t0 = std::chrono::steady_clock::now();
	#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
	for (int b = 0; b < nblocks; b++) { int ol; ref_bwt_forward(ins[b], lens[b], outs[b], &ol); }
	return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

double ref_bwt_inverse_batch(const unsigned char* const* ins, const int* lens_with_trailer, unsigned char* const* outs,
                             int nblocks, int threads, int threads_per_block)
{
	auto t0 = std::chrono::steady_clock::now();
	#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
	for (int b = 0; b < nblocks; b++) { int ol; ref_bwt_inverse(ins[b], lens_with_trailer[b], outs[b], &ol, threads_per_block, 0); }
	return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

/* divsufsort.cpp:1721 itself (second call site: lz77.cpp:141) */
int ref_divsufsort(const unsigned char* T, int* SA, int n) { return divsufsort(T, SA, n); }

int ref_core_count(void) { return (int)GetCoreCount(); }

}

/* Second stage, first half, by the reference's own classes: what Ans::Encode does to each StackSize chunk before the
 * entropy coder (ans.cpp:149-160) -- Postcoder::Encode (rank.cpp:45-90) then RLE::encode (rle.cpp:22-47). */
#include "rank.hpp"
#include "rle.hpp"
extern "C" int ref_src_rle0(const unsigned char* in, int len, int* freq, unsigned short* rle, int* rlen)
{
	const int StackSize = 1 << 20;                       /* ans.hpp:33 */
	Postcoder rank; RLE rle0;
	unsigned char* tmp = (unsigned char*)malloc(StackSize + 16);
	int chunks = 0;
	for (int in_p = 0; in_p < len; in_p += StackSize, chunks++) {
		int n = ((in_p + StackSize) < len) ? StackSize : (len - in_p);
		memcpy(tmp, in + in_p, n);
		tmp[n] = 0xA5;                                   /* rle.cpp:34 reads in[i + run] before it checks the bound */
		rank.Encode(tmp, freq + 256 * chunks, n);
		int rl = n;
		rle0.encode(tmp, rle + (size_t)StackSize * chunks, &rl);
		rlen[chunks] = rl;
	}
	free(tmp);
	return chunks;
}
