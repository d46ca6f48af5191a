// gather_probe.cu -- which load flavour / device limit decides how many DRAM sectors a random 4-byte gather costs?
// Run under: ncu --metrics dram__sectors_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,gpu__time_duration.sum ./gather_probe <gran>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
typedef uint32_t u32; typedef uint64_t u64;
__global__ void fill(u32* tab, u32 n) { u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) tab[i] = ((u32)i * 0x9E3779B1u + 0x7F4A7C15u) & (n - 1); }
template <int V> __device__ __forceinline__ u32 ld(const u32* p)
{
	u32 v;
	if (V == 0) v = *p;
	else if (V == 1) v = __ldg(p);
	else if (V == 2) v = __ldcg(p);
	else if (V == 3) v = __ldcs(p);
	else if (V == 4) v = __ldcv(p);
	else if (V == 5) asm volatile("ld.global.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
	else if (V == 6) asm volatile("ld.global.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
	else if (V == 7) asm volatile("ld.global.L1::evict_first.u32 %0, [%1];" : "=r"(v) : "l"(p));
	else if (V == 8) { u64 pol; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
	                   asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol)); }
	else if (V == 9) asm volatile("ld.global.L2::128B.u32 %0, [%1];" : "=r"(v) : "l"(p));
	else v = *p;
	return v;
}
template <int V> __global__ void __launch_bounds__(256) walk(const u32* __restrict__ tab, u32 n, int steps, u32* sink)
{
	u32 gid = blockIdx.x * blockDim.x + threadIdx.x;
	u32 p = (gid * 2654435761u + 12345u) & (n - 1);
	for (int i = 0; i < steps; i++) p = ld<V>(tab + p);
	if (p == 0xffffffffu) sink[0] = p;
}
template <int V> void run(const char* name, const u32* tab, u32 n, u32* sink)
{
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	const int blocks = 148 * 8, steps = 128;
	walk<V><<<blocks, 256>>>(tab, n, steps, sink);
	cudaEventRecord(a); walk<V><<<blocks, 256>>>(tab, n, steps, sink); cudaEventRecord(b); cudaDeviceSynchronize();
	float ms; cudaEventElapsedTime(&ms, a, b);
	printf("%-28s %8.2f G gathers/s  (%s)\n", name, blocks * 256.0 * steps / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main(int argc, char** argv)
{
	size_t gran = argc > 1 ? atoi(argv[1]) : 0, got = 0;
	size_t mib = argc > 2 ? atoi(argv[2]) : 256;
	if (gran) printf("set limit %zu -> %s\n", gran, cudaGetErrorString(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran)));
	cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity); printf("MaxL2FetchGranularity = %zu, table %zu MiB\n", got, mib);
	u32 n = (u32)((mib << 20) / 4); u32 *tab, *sink; cudaMalloc(&tab, (size_t)n * 4); cudaMalloc(&sink, 64);
	fill<<<(n + 255) / 256, 256>>>(tab, n);
	run<0>("plain ld.global", tab, n, sink); run<1>("ld.global.nc (__ldg)", tab, n, sink); run<2>("ld.cg", tab, n, sink);
	run<3>("ld.cs", tab, n, sink); run<4>("ld.cv", tab, n, sink); run<5>("ld.L2::64B", tab, n, sink);
	run<6>("ld.L1::no_allocate", tab, n, sink); run<7>("ld.L1::evict_first", tab, n, sink); run<8>("ld.L2::cache_hint evict_first", tab, n, sink);
	run<9>("ld.L2::128B", tab, n, sink);
	return 0;
}
