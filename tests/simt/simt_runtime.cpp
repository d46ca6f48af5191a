// simt_runtime.cpp -- TEST INFRASTRUCTURE ONLY (see simt_runtime.hpp): fiber scheduler and collectives.
#include "simt_runtime.hpp"
#undef threadIdx
#undef blockIdx
#undef blockDim
#undef gridDim

namespace simt {

uint3 g_threadIdx, g_blockIdx;
dim3 g_blockDim, g_gridDim;
Block g_block;
unsigned char* g_dynamic_smem = nullptr;
long long g_launches = 0;
static std::vector<unsigned char> g_smem_store;

void fail(const char* what) { fprintf(stderr, "[simt] %s (block %u thread %u)\n", what, g_blockIdx.x, g_threadIdx.x); abort(); }

void complete(Warp& w)
{
	unsigned ballot = 0;
	for (int l = 0; l < 32; l++) if (((w.live >> l) & 1u) && w.val[l]) ballot |= 1u << l;
	for (int l = 0; l < 32; l++) {
		if (!((w.live >> l) & 1u)) continue;
		const int a = w.arg[l];
		int src = l;
		switch (w.op) {
		case OP_BALLOT: w.out[l] = ballot; continue;
		case OP_SYNCWARP: w.out[l] = 0; continue;
		case OP_MATCH: { unsigned m = 0; for (int k = 0; k < 32; k++) if (((w.live >> k) & 1u) && w.val[k] == w.val[l]) m |= 1u << k; w.out[l] = m; continue; }
		case OP_SHFL: src = a & 31; break;
		case OP_SHFL_UP: src = l - a >= 0 ? l - a : l; break;
		case OP_SHFL_DOWN: src = l + a < 32 ? l + a : l; break;
		case OP_SHFL_XOR: src = (l ^ a) & 31; break;
		default: fail("bad collective");
		}
		w.out[l] = ((w.live >> src) & 1u) ? w.val[src] : w.val[l];   // reading an exited lane is undefined on the GPU: keep own value
	}
	w.arrived = 0; w.op = OP_NONE; w.gen++;
}

#if defined(__x86_64__)
// Minimal cooperative context switch: push the callee-saved registers, swap stack pointers, pop, return.
extern "C" void simt_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl simt_switch
.type simt_switch,@function
simt_switch:
	pushq %rbp
	pushq %rbx
	pushq %r12
	pushq %r13
	pushq %r14
	pushq %r15
	movq %rsp, (%rdi)
	movq %rsi, %rsp
	popq %r15
	popq %r14
	popq %r13
	popq %r12
	popq %rbx
	popq %rbp
	ret
.size simt_switch,.-simt_switch
)");
static void* g_main_sp = nullptr;
static void fiber_entry();
static void to_fiber(Fiber& f) { simt_switch(&g_main_sp, f.sp); }
static void to_main(Fiber& f) { simt_switch(&f.sp, g_main_sp); }
static void prepare_fiber(Fiber& f)
{
	// initial frame: six zeroed callee-saved registers, then the return address = fiber_entry, on a 16-byte aligned stack
	uintptr_t top = ((uintptr_t)f.stack + STACK_BYTES) & ~(uintptr_t)15;
	void** sp = (void**)(top - 8);            // after `ret` pops the entry address, rsp % 16 == 8 as at a normal function entry
	*--sp = (void*)fiber_entry;
	for (int i = 0; i < 6; i++) *--sp = nullptr;
	f.sp = sp;
}
#else
static void to_fiber(Fiber& f) { swapcontext(&g_block.main_ctx, &f.ctx); }
static void to_main(Fiber& f) { swapcontext(&f.ctx, &g_block.main_ctx); }
static void fiber_entry();
static void prepare_fiber(Fiber& f)
{
	getcontext(&f.ctx);
	f.ctx.uc_stack.ss_sp = f.stack; f.ctx.uc_stack.ss_size = STACK_BYTES; f.ctx.uc_link = nullptr;
	makecontext(&f.ctx, fiber_entry, 0);
}
#endif

void yield()
{
	Block& b = g_block;
	to_main(b.fibers[b.current]);
}

static void barrier_release(Block& b) { b.bar_or_out = b.bar_or; b.bar_or = 0; b.bar_arrived = 0; b.bar_gen++; }

static void barrier(int pred)
{
	Block& b = g_block;
	if (pred) b.bar_or = 1;
	b.bar_arrived++;
	const unsigned mygen = b.bar_gen;
	if (b.bar_arrived == b.live_threads) barrier_release(b);
	while (b.bar_gen == mygen) yield();
}

static void fiber_entry()
{
	Block& b = g_block;
	const int t = b.current;
	(*b.body)();
	// thread exit: it no longer takes part in collectives or barriers
	b.fibers[t].done = true;
	Warp& w = b.warps[t >> 5];
	w.live &= ~(1u << (t & 31));
	b.live_threads--;
	if (w.live && w.arrived && w.arrived == w.live) complete(w);
	if (b.live_threads && b.bar_arrived && b.bar_arrived == b.live_threads) barrier_release(b);
	to_main(b.fibers[t]);
	fail("a finished fiber was resumed");
}

static void run_block(const std::function<void()>& body, unsigned nthreads)
{
	Block& b = g_block;
	if (b.fibers.size() < nthreads) b.fibers.resize(nthreads);
	b.warps.assign((nthreads + 31) / 32, Warp());
	b.nthreads = b.live_threads = nthreads; b.bar_arrived = 0; b.bar_or = 0; b.body = &body;
	for (unsigned t = 0; t < nthreads; t++) {
		Fiber& f = b.fibers[t];
		if (!f.stack) f.stack = (char*)malloc(STACK_BYTES);
		prepare_fiber(f);
		f.done = false;
		b.warps[t >> 5].live |= 1u << (t & 31);
	}
	unsigned long long idle_sweeps = 0;
	while (b.live_threads) {
		const unsigned before = b.live_threads;
		bool progressed = false;
		for (unsigned t = 0; t < nthreads; t++) {
			if (b.fibers[t].done) continue;
			b.current = (int)t;
			g_threadIdx = uint3{t, 0, 0};
			const unsigned wgen = b.warps[t >> 5].gen, bgen = b.bar_gen, warr = b.warps[t >> 5].arrived, barr = b.bar_arrived;
			to_fiber(b.fibers[t]);
			if (b.fibers[t].done || b.warps[t >> 5].gen != wgen || b.bar_gen != bgen || b.warps[t >> 5].arrived != warr || b.bar_arrived != barr) progressed = true;
		}
		if (b.live_threads != before) progressed = true;
		if (!progressed && ++idle_sweeps > 4) fail("deadlock: every live thread of the block is waiting");
		if (progressed) idle_sweeps = 0;
	}
}

void launch(dim3 grid, dim3 block, size_t smem, cudaStream_t, const std::function<void()>& body)
{
	g_launches++;
	if (block.y != 1 || block.z != 1 || grid.y != 1 || grid.z != 1) fail("only 1-D launches are emulated");
	if (g_smem_store.size() < smem + 64) g_smem_store.resize(smem + 64);
	g_dynamic_smem = g_smem_store.data();
	g_gridDim = grid; g_blockDim = block;
	for (unsigned bx = 0; bx < grid.x; bx++) {
		g_blockIdx = uint3{bx, 0, 0};
		run_block(body, block.x);
	}
}

} // namespace simt

void __syncthreads() { simt::barrier(0); }
int __syncthreads_or(int pred) { simt::barrier(pred); return (int)simt::g_block.bar_or_out; }
