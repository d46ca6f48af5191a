// divsufsort_shim.cpp -- the reference's divsufsort() symbol on top of the B200 suffix sorter.
//
// Link this object INSTEAD of the reference's divsufsort.cpp (together with bwt_shim.cpp instead of bwt.cpp) and the
// -m2 suffix-array match finder of lz77.cpp:134-146 builds its SA on the GPU as well; nothing else of
// divsufsort.hpp (divbwt) is used anywhere in Jampack. Return convention of divsufsort.cpp:1721-1747:
// 0 on success, -1 for bad arguments, -2 otherwise.
#include <stdint.h>
#include "jp_bwt.h"

extern "C" int divsufsort(const unsigned char *T, int *SA, int n)
{
	if(T == 0 || SA == 0 || n < 0)
		return -1;
	if(n == 0)
		return 0;
	return jp_bwt_suffix_array(T, n, (int32_t*)SA) == JP_OK ? 0 : -2;
}
