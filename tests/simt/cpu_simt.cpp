/* TEST INFRASTRUCTURE ONLY - fiber scheduler behind cpu_simt.h (x86-64 only). */
#include "cpu_simt.h"

#if !defined(__x86_64__)
#error "the SIMT emulation harness only supports x86-64"
#endif

asm(R"(
.text
.globl simt_ctx_switch
.type simt_ctx_switch,@function
simt_ctx_switch:
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
.size simt_ctx_switch,.-simt_ctx_switch
)");

#undef threadIdx
#undef blockIdx
#undef blockDim
#undef gridDim

namespace simt {

thread_local Block *g_blk = nullptr;
thread_local dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;

static const size_t STACK_BYTES = 256 * 1024;

void yield_to_main()
{
	Block *b = g_blk;
	Fiber &f = b->fibers[b->cur];
	simt_ctx_switch(&f.sp, b->main_sp);
}

static void fiber_entry()
{
	Block *b = g_blk;
	b->body();
	Fiber &f = b->fibers[b->cur];
	f.done = true;
	b->warp_live[b->cur >> 5]--;
	b->blk_live--;
	yield_to_main();
	abort();
}

void warp_barrier()
{
	Block *b = g_blk;
	unsigned w = (unsigned)b->cur >> 5;
	unsigned my = b->warp_gen[w];
	b->warp_arrived[w]++;
	for (;;) {
		if (b->warp_gen[w] != my)
			return;
		if (b->warp_arrived[w] >= b->warp_live[w]) {
			b->warp_arrived[w] = 0;
			b->warp_gen[w]++;
			return;
		}
		yield_to_main();
	}
}

void block_barrier()
{
	Block *b = g_blk;
	unsigned my = b->blk_gen;
	b->blk_arrived++;
	for (;;) {
		if (b->blk_gen != my)
			return;
		if (b->blk_arrived >= b->blk_live) {
			b->blk_arrived = 0;
			b->blk_gen++;
			return;
		}
		yield_to_main();
	}
}

void run_block(Block &b)
{
	g_blk = &b;
	memset(b.warp_arrived, 0, sizeof(b.warp_arrived));
	memset(b.warp_gen, 0, sizeof(b.warp_gen));
	memset(b.warp_live, 0, sizeof(b.warp_live));
	b.blk_arrived = b.blk_gen = 0;
	b.blk_live = b.nthreads;
	for (unsigned i = 0; i < b.nthreads; i++) {
		Fiber &f = b.fibers[i];
		f.done = false;
		b.warp_live[i >> 5]++;
		uintptr_t top = ((uintptr_t)f.stack + STACK_BYTES) & ~(uintptr_t)15;
		uint64_t *sp = (uint64_t *)top;
		*--sp = 0;                         /* fake return address for fiber_entry */
		*--sp = (uint64_t)(uintptr_t)&fiber_entry;
		for (int r = 0; r < 6; r++)
			*--sp = 0;                 /* rbp rbx r12 r13 r14 r15 */
		f.sp = sp;
	}
	unsigned remaining = b.nthreads;
	while (remaining) {
		remaining = 0;
		for (unsigned i = 0; i < b.nthreads; i++) {
			Fiber &f = b.fibers[i];
			if (f.done)
				continue;
			b.cur = (int)i;
			g_threadIdx = dim3(i, 0, 0);
			simt_ctx_switch(&b.main_sp, f.sp);
			if (!f.done)
				remaining++;
		}
	}
	g_blk = nullptr;
}

static thread_local std::vector<uint8_t> g_dyn;
void set_dyn_smem(size_t bytes) { if (g_dyn.size() < bytes + 64) g_dyn.resize(bytes + 64); }
uint8_t *dyn_smem() { return (uint8_t *)(((uintptr_t)g_dyn.data() + 15) & ~(uintptr_t)15); }

void launch(dim3 grid, dim3 block, const std::function<void()> &body)
{
	static thread_local Block blk;
	if (block.x > 2048 || block.y != 1 || block.z != 1 || grid.y != 1 || grid.z != 1) {
		fprintf(stderr, "simt: unsupported launch shape\n");
		abort();
	}
	if (blk.fibers.size() < block.x) {
		size_t old = blk.fibers.size();
		blk.fibers.resize(block.x);
		for (size_t i = old; i < block.x; i++)
			blk.fibers[i].stack = (char *)malloc(STACK_BYTES);
	}
	blk.nthreads = block.x;
	blk.body = body;
	g_blockDim = block;
	g_gridDim = grid;
	for (unsigned bx = 0; bx < grid.x; bx++) {
		g_blockIdx = dim3(bx, 0, 0);
		run_block(blk);
	}
}

}  // namespace simt
