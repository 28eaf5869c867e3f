// cuda_emu.cpp -- TEST INFRASTRUCTURE ONLY.  Fiber scheduler behind cuda_emu.h.
#include "cuda_emu.h"

#include <sys/mman.h>

namespace emu {

thread_local Block *tl_block = nullptr;
thread_local Fiber *tl_fiber = nullptr;
std::mutex g_alloc_mu;
std::set<std::pair<uintptr_t, size_t>> g_dev_allocs;

bool is_device_ptr(const void *p)
{
	std::lock_guard<std::mutex> g(g_alloc_mu);
	uintptr_t a = (uintptr_t)p;
	auto it = g_dev_allocs.upper_bound({a, (size_t)-1});
	if (it == g_dev_allocs.begin()) return false;
	--it;
	return a >= it->first && a < it->first + it->second + 1;
}

static const size_t kStack = 256 * 1024;

// Minimal x86-64 SysV context switch: saves callee-saved registers on the
// current stack, stores rsp to *from, loads rsp from 'to', restores, returns.
extern "C" void emu_switch(void **from, void *to);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
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
.size emu_switch,.-emu_switch
)");

static void fiber_entry();

static void prepare(Fiber &f)
{
	if (!f.stack) {
		f.stack = (uint8_t *)mmap(nullptr, kStack, PROT_READ | PROT_WRITE,
					  MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
		if (f.stack == (uint8_t *)MAP_FAILED) { perror("mmap"); abort(); }
	}
	// initial frame: 6 callee-saved regs + return address (fiber_entry) + alignment pad
	uintptr_t top = ((uintptr_t)f.stack + kStack) & ~(uintptr_t)15;
	void **sp = (void **)top;
	*--sp = nullptr;			// fake return address of fiber_entry (keeps 16B alignment at entry)
	*--sp = (void *)fiber_entry;	// 'ret' target
	for (int i = 0; i < 6; i++) *--sp = nullptr;
	f.sp = sp;
	f.done = false;
}

static void switch_to(Block *b, unsigned next)
{
	Fiber *from = tl_fiber;
	Fiber *to = &b->fibers[next];
	b->cur = next;
	tl_fiber = to;
	emu_switch(&from->sp, to->sp);
}

void yield()
{
	Block *b = tl_block;
	if (++b->spins > 200000000ul) {
		fprintf(stderr, "cuda_emu: deadlock suspected (block %u): a barrier/collective is never completed\n", b->bid.x);
		abort();
	}
	unsigned n = b->nthreads;
	unsigned i = b->cur;
	for (unsigned k = 0; k < n; k++) {
		i = (i + 1 == n) ? 0 : i + 1;
		if (!b->fibers[i].done) {
			if (i != b->cur) switch_to(b, i);
			return;
		}
	}
}

static void fiber_entry()
{
	Block *b = tl_block;
	Fiber *f = tl_fiber;
	(*b->body)();
	f->done = true;
	b->live--;
	b->spins = 0;
	// A finished thread no longer takes part in barriers: if everyone else is
	// already waiting, release them (CUDA semantics for exited threads).
	if (b->live && b->bar_arrived == b->live) {
		b->bar_res[b->bar_gen & 1] = b->bar_acc;
		b->bar_acc = 0;
		b->bar_arrived = 0;
		b->bar_gen++;
	}
	if (b->live == 0) {
		void *dummy;
		emu_switch(&dummy, b->sched_sp);
	}
	for (;;) {
		unsigned n = b->nthreads, i = b->cur;
		for (unsigned k = 0; k < n; k++) {
			i = (i + 1 == n) ? 0 : i + 1;
			if (!b->fibers[i].done) break;
		}
		void *dummy;
		b->cur = i;
		tl_fiber = &b->fibers[i];
		emu_switch(&dummy, tl_fiber->sp);
	}
}

static void run_block(Block &b, unsigned bid, dim3 grid, dim3 block, size_t smem,
		      const std::function<void()> &body)
{
	unsigned nt = block.x * block.y * block.z;
	b.nthreads = nt;
	b.live = nt;
	b.bid = uint3{bid, 0, 0};
	b.bdim = block;
	b.gdim = grid;
	b.body = &body;
	b.bar_arrived = 0;
	b.bar_acc = 0;
	b.spins = 0;
	if (b.fibers.size() < nt) b.fibers.resize(nt);
	b.warps.assign((nt + 31) / 32, Warp());
	if (b.dyn_smem_size < smem + 64) {
		free(b.dyn_smem);
		b.dyn_smem = (uint8_t *)aligned_alloc(128, (smem + 64 + 127) & ~(size_t)127);
		b.dyn_smem_size = smem + 64;
	}
	for (unsigned t = 0; t < nt; t++) {
		Fiber &f = b.fibers[t];
		prepare(f);
		f.tid = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
		f.lane = t & 31;
		f.warp = t >> 5;
	}
	tl_block = &b;
	tl_fiber = &b.fibers[0];
	b.cur = 0;
	Fiber sched;
	// run until the last fiber switches back to sched_sp
	Fiber *first = &b.fibers[0];
	emu_switch(&b.sched_sp, first->sp);
	tl_block = nullptr;
	tl_fiber = nullptr;
	(void)sched;
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body)
{
	unsigned nblocks = grid.x * grid.y * grid.z;
	unsigned nworkers = std::thread::hardware_concurrency();
	if (nworkers == 0) nworkers = 4;
	if (const char *e = getenv("LDB_EMU_THREADS")) nworkers = (unsigned)atoi(e);
	if (nworkers > nblocks) nworkers = nblocks;
	if (nworkers == 0) return;
	std::atomic<unsigned> next{0};
	// Block objects (and their fiber stacks) are pooled across launches.
	static std::mutex pool_mu;
	static std::vector<Block *> pool;
	auto worker = [&]() {
		Block *b = nullptr;
		{
			std::lock_guard<std::mutex> g(pool_mu);
			if (!pool.empty()) { b = pool.back(); pool.pop_back(); }
		}
		if (!b) b = new Block();
		for (;;) {
			unsigned bid = next.fetch_add(1);
			if (bid >= nblocks) break;
			run_block(*b, bid, grid, block, smem, body);
		}
		std::lock_guard<std::mutex> g(pool_mu);
		pool.push_back(b);
	};
	if (nworkers == 1) {
		worker();
		return;
	}
	std::vector<std::thread> th;
	for (unsigned i = 0; i < nworkers; i++) th.emplace_back(worker);
	for (auto &t : th) t.join();
}

} // namespace emu
