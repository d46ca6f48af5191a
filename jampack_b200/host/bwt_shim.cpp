// bwt_shim.cpp -- BlockSort::Bwt on top of the C-ABI of include/jp_bwt.h.
//
// Link this object INSTEAD of the reference's bwt.cpp (and divsufsort.cpp, unless the -m2 match finder of
// lz77.cpp:141 is wanted) and the untouched jampack.cpp consumes the B200 stage: Jampack::Comp calls
// Bwt->ForwardBwt(Input, Output) (jampack.cpp:40), Jampack::Decomp calls Bwt->InverseBwt(Input, Output, Option)
// (jampack.cpp:50). Size conventions and side effects are the reference's (bwt.cpp:27, :77-78); failures go
// through the reference's Error() (format.cpp:6-10), like every other stage.
//
// Build with -DJP_STANDALONE_STAGE to use jp_stage.hpp instead of the reference's headers (then the
// application supplies Error()).
#ifdef JP_STANDALONE_STAGE
#include "jp_stage.hpp"
#else
#include "bwt.hpp"
#endif
#include "jp_bwt.h"

// The CUDA contexts come up while the reference parses its arguments, reads the first batch and runs LZ77 on it.
namespace { struct JpWarmup { JpWarmup() { jp_bwt_warmup_async(); } } jp_warmup_at_load; }

void BlockSort::Bwt::ForwardBwt(Buffer Input, Buffer Output)
{
	int32_t out_len = 0;
	const int rc = jp_bwt_forward(Input.block, *Input.size, Output.block, &out_len);
	*Output.size = *Input.size + (BWT_UNITS * sizeof(Index));   // bwt.cpp:27
	if(rc != JP_OK)
		Error(jp_bwt_strerror(rc));
}

// Opt.Threads / Opt.Gpu only shape the reference's CPU loop (bwt.cpp:92-132); the output does not depend
// on them, and the device runs all 120 decode units at once.
void BlockSort::Bwt::InverseBwt(Buffer Input, Buffer Output, Options Opt)
{
	(void)Opt;
	int32_t out_len = 0;
	const int rc = jp_bwt_inverse(Input.block, *Input.size, Output.block, &out_len);
	*Input.size -= (BWT_UNITS * sizeof(Index));                 // bwt.cpp:77 mutates the caller's size
	*Output.size = *Input.size;                                 // bwt.cpp:78
	if(rc != JP_OK)
		Error(jp_bwt_strerror(rc));
}
