/* TEST INFRASTRUCTURE ONLY -- not part of the product path.
 *
 * The reference's own, legacy CUDA inverse (bwt.cpp:8-19 kernel CUDAInverse, bwt.cpp:186-240 host side, taken when
 * Options::Gpu is set): the UNMODIFIED sources compiled by nvcc from where they lie under /root/reference
 * (oracle/Makefile target ref_cuda -> oracle/_ref/libjamref_cuda.so, git-ignored). It is a reported baseline
 * (BASELINE.md plan item 5), never a checker and never on the product path.
 *
 * The reference cannot take this path on a cc >= 7 device as shipped: GetCudaCoreCount() (sys_detect.cpp:105-136) knows
 * cores-per-SM only up to Pascal and answers 0, which bwt.cpp:106-132 turns into a division by zero. The function does
 * consult a cache first -- System::Gpu::Cores, "to skip querying ... once we already know the hardware"
 * (sys_detect.hpp:3-4, sys_detect.cpp:14-19) -- so the wrapper states the hardware there (SMs x 128) before calling
 * in. No reference source is modified or copied.
 */
#include "bwt.hpp"
#include <chrono>
#include <cuda_runtime.h>

namespace System { namespace Gpu { extern int64_t Cores; } }

extern "C" {

/* Returns wall seconds of BlockSort::Bwt::InverseBwt with Options::Gpu = true (host Map build + 6N of PCIe copies +
 * the 120-thread kernel + copy back, exactly what the reference would do), or a negative value on failure. */
double ref_bwt_inverse_legacy_cuda(const unsigned char* in, int len_with_trailer, unsigned char* out, int* out_len)
{
	int dev_count = 0;
	if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count < 1) return -1.0;
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) return -2.0;
	System::Gpu::Cores = (int64_t)prop.multiProcessorCount * 128;
	cudaFree(0);                                   /* context creation is not part of the timed call */
	int isz = len_with_trailer, osz = 0;
	Buffer I; I.block = const_cast<unsigned char*>(in); I.size = &isz;
	Buffer O; O.block = out; O.size = &osz;
	Options opt; memset(&opt, 0, sizeof(opt));
	opt.Threads = 1; opt.Gpu = true; opt.Multiblock = true;
	BlockSort::Bwt b;
	auto t0 = std::chrono::steady_clock::now();
	b.InverseBwt(I, O, opt);
	const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	*out_len = osz;
	return s;
}

}
