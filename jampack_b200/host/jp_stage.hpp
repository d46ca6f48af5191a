// jp_stage.hpp -- C++ host mirror of the reference's stage interface for the BWT path, for builds that do
// not have the reference's headers. Same names, argument meaning and error behaviour:
//   Buffer / Options / Index / BWT_UNITS   reference format.hpp:26,32,37-54
//   Error(const char*)                     reference format.hpp:59, format.cpp:6-10 (prints, exit(-1))
//   BlockSort::Bwt::ForwardBwt/InverseBwt  reference bwt.hpp:13-18
// When the reference's own headers are on the include path, bwt_shim.cpp uses those instead and this file
// is not needed: the object then links in place of the reference's bwt.cpp + divsufsort.cpp.
#ifndef JP_STAGE_HPP
#define JP_STAGE_HPP

#include <stdint.h>

#ifndef BWT_UNITS
#define BWT_UNITS 120
#endif

typedef int Index;

struct Buffer
{
	unsigned char *block;
	Index *size;
};

struct Options
{
	Index BlockSize;
	unsigned int MatchFinder;
	unsigned int Threads;
	unsigned int Filters;
	bool Gpu;
	bool Multiblock;
};

extern void Error(const char *string);

namespace BlockSort
{
	class Bwt
	{
		public:
		void ForwardBwt(Buffer Input, Buffer Output);
		void InverseBwt(Buffer Input, Buffer Output, Options Opt);
	};
};

#endif // JP_STAGE_HPP
