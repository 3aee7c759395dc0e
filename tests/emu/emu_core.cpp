// emu_core.cpp — TEST INFRASTRUCTURE: the engine behind tests/emu/cuda_runtime.h.
//
// Execution model.  A grid is a set of CTAs; a CTA is ONE OS thread that runs its CUDA threads as fibers (own stack each, switched
// by a few instructions of assembly) round-robin: a CUDA thread runs until it has to wait (barrier, warp collective, mbarrier,
// spin on global memory), then the next one runs.  Inside a CTA nothing is concurrent, so shared memory needs no atomics; CTAs of
// a grid run on different OS threads, so global-memory atomics are real atomics.  Small grids (<= 64 CTAs) get one OS thread per
// CTA — every CTA is "resident", which is what kernels that wait for each other's CTAs (the step kernel) need; larger grids are
// dealt in index order to a pool (forward progress as on a device: a CTA may wait for lower-numbered ones).
// A CTA that makes no progress for GAS_EMU_DEADLOCK_S seconds (default 60) aborts the process with a per-thread report.
//
// Runtime model.  Streams execute at enqueue time on the calling thread; events are time stamps; a capturing stream (and the
// streams that joined its capture by waiting for one of its events) records closures instead, and a graph launch replays them in
// record order.  Calls that the real runtime refuses while the calling thread is capturing (cudaMalloc, synchronous copies,
// synchronising a capturing stream) are refused here too.
#include <cuda_runtime.h>

#include <execinfo.h>
#include <fcntl.h>
#include <unistd.h>
#include <map>
#include <sched.h>
#include <signal.h>
#include <sys/mman.h>
#include <time.h>

#include <condition_variable>
#include <memory>
#include <thread>

namespace emu {

thread_local FiberInfo *t_fiber = nullptr;
thread_local CtaInfo *t_cta = nullptr;

unsigned long long now_ns() {
	timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (unsigned long long)ts.tv_sec * 1000000000ULL + (unsigned long long)ts.tv_nsec;
}

static int env_int(const char *name, int dflt) {
	const char *e = getenv(name);
	return e && *e ? atoi(e) : dflt;
}
int emulated_sms() {
	static const int n = std::max(1, std::min(64, env_int("GAS_EMU_SMS", 6)));
	return n;
}

// ---- fibers ------------------------------------------------------------------------------------------------------------------
extern "C" void emu_switch(void **save_sp, void *load_sp);
asm(R"(
	.text
	.globl emu_switch
	.type emu_switch, @function
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
	.size emu_switch, .-emu_switch
)");

constexpr size_t kStackBytes = 512 * 1024;

struct WarpState {
	uint64_t dep[32];
	uint64_t res[32];
	int op[32], arg[32], width[32];
	uint32_t want[32];  // member mask of the collective the lane waits in
	bool waiting[32], released[32];
	uint32_t gone; // lanes that have exited or never existed
};
struct BarState {
	int count;
	unsigned gen;
};
struct Fiber {
	FiberInfo info;
	void *sp;
	bool done;
	const char *waiting; // what the thread is blocked on (deadlock report)
	unsigned wait_mask;
	int wait_op;
};
struct Cta {
	CtaInfo info;
	std::vector<Fiber> fibers;
	std::vector<WarpState> warps;
	BarState bars[16];
	unsigned exited;
	void *sched_sp;
	Fiber *cur;
	const std::function<void()> *body;
	const char *kernel;
	bool progress;
};
static thread_local Cta *t_self = nullptr;

static void fiber_entry() {
	Cta *c = t_self;
	(*c->body)();
	Fiber *f = c->cur;
	f->done = true;
	c->exited++;
	c->warps[f->info.warp].gone |= 1u << f->info.lane;
	c->progress = true;
	void *dummy;
	emu_switch(&dummy, c->sched_sp);
	abort(); // never resumed
}

void yield_blocked() {
	Cta *c = t_self;
	emu_switch(&c->cur->sp, c->sched_sp);
}

static inline void wait_note(const char *what) { t_self->cur->waiting = what; }
static inline void wait_done() {
	t_self->cur->waiting = nullptr;
	t_self->progress = true;
}

void syncthreads() {
	Cta *c = t_self;
	BarState &b = c->bars[0];
	const unsigned g = b.gen;
	b.count++;
	wait_note("__syncthreads");
	for (;;) {
		if (b.gen != g) {
			break;
		}
		if (b.count >= (int)(c->info.nthreads - c->exited)) {
			b.count = 0;
			b.gen++;
			break;
		}
		yield_blocked();
	}
	wait_done();
}

void bar_sync(int id, int nthreads) {
	Cta *c = t_self;
	BarState &b = c->bars[id & 15];
	const unsigned g = b.gen;
	b.count++;
	wait_note("bar.sync");
	for (;;) {
		if (b.gen != g) {
			break;
		}
		if (b.count >= nthreads) {
			b.count = 0;
			b.gen++;
			break;
		}
		yield_blocked();
	}
	wait_done();
}

void bar_arrive(int id, int nthreads) {
	Cta *c = t_self;
	BarState &b = c->bars[id & 15];
	b.count++;
	if (b.count >= nthreads) {
		b.count = 0;
		b.gen++;
	}
	c->progress = true;
}

// A collective is identified by its member mask: the lanes that call with the same mask meet.  (Lane pairs inside a warp run
// pair-masked collectives while their neighbours already wait at a full-mask one, so the lowest lane alone does not identify it.)
static void collective_complete(WarpState &w, unsigned mask) {
	const unsigned in = mask & ~w.gone;
	unsigned ballot = 0;
	unsigned sum = 0;
	for (int l = 0; l < 32; l++) {
		if (in >> l & 1u) {
			if (w.dep[l] & 0xffffffffu) {
				ballot |= 1u << l;
			}
			sum += (unsigned)w.dep[l];
		}
	}
	for (int l = 0; l < 32; l++) {
		if (!(in >> l & 1u)) {
			continue;
		}
		const int wd = w.width[l] > 0 && w.width[l] <= 32 ? w.width[l] : 32;
		const int base = l & ~(wd - 1);
		int src = l;
		switch (w.op[l]) {
			case OP_SHFL_IDX: src = base + (w.arg[l] & (wd - 1)); break;
			case OP_SHFL_XOR: src = l ^ w.arg[l]; src = (src >= base && src < base + wd) ? src : l; break;
			case OP_SHFL_UP: src = l - w.arg[l]; src = src >= base ? src : l; break;
			case OP_SHFL_DOWN: src = l + w.arg[l]; src = src < base + wd ? src : l; break;
			default: break;
		}
		switch (w.op[l]) {
			case OP_SHFL_IDX:
			case OP_SHFL_XOR:
			case OP_SHFL_UP:
			case OP_SHFL_DOWN:
				// reading a lane that does not take part is undefined on the device: here it reads that lane's last deposit
				w.res[l] = w.dep[src];
				break;
			case OP_BALLOT: w.res[l] = ballot; break;
			case OP_ANY: w.res[l] = ballot != 0; break;
			case OP_ALL: w.res[l] = ballot == in; break;
			case OP_REDUCE_ADD: w.res[l] = sum; break;
			default: w.res[l] = 0; break;
		}
	}
	for (int l = 0; l < 32; l++) {
		if (in >> l & 1u) {
			w.waiting[l] = false;
			w.released[l] = true;
		}
	}
}

uint64_t warp_collective(unsigned mask, int op, int arg, int width, uint64_t value) {
	Cta *c = t_self;
	Fiber *f = c->cur;
	WarpState &w = c->warps[f->info.warp];
	const int lane = (int)f->info.lane;
	if (!(mask >> lane & 1u)) {
		fprintf(stderr, "[emu] %s: lane %d executes a warp collective whose mask %08x does not name it\n", c->kernel, lane, mask);
		abort();
	}
	w.dep[lane] = value;
	w.op[lane] = op;
	w.arg[lane] = arg;
	w.width[lane] = width;
	w.want[lane] = mask;
	w.released[lane] = false;
	w.waiting[lane] = true;
	wait_note("warp collective");
	f->wait_mask = mask;
	f->wait_op = op;
	for (;;) {
		if (w.released[lane]) {
			break;
		}
		bool all_in = true;
		for (unsigned m = mask & ~w.gone; m; m &= m - 1) {
			const int l = __builtin_ctz(m);
			if (!w.waiting[l] || w.want[l] != mask) {
				all_in = false;
				break;
			}
		}
		if (all_in) {
			collective_complete(w, mask);
			break;
		}
		yield_blocked();
	}
	wait_done();
	return w.res[lane];
}

// ---- stacks --------------------------------------------------------------------------------------------------------------------
static std::mutex g_stack_mu;
static std::vector<std::pair<size_t, unsigned char *>> g_stack_pool;
static unsigned char *stacks_get(size_t bytes) {
	{
		std::lock_guard<std::mutex> lk(g_stack_mu);
		for (size_t i = 0; i < g_stack_pool.size(); i++) {
			if (g_stack_pool[i].first == bytes) {
				unsigned char *p = g_stack_pool[i].second;
				g_stack_pool.erase(g_stack_pool.begin() + (long)i);
				return p;
			}
		}
	}
	void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
	if (p == MAP_FAILED) {
		fprintf(stderr, "[emu] cannot map %zu bytes of fiber stacks\n", bytes);
		abort();
	}
	return (unsigned char *)p;
}
static void stacks_put(size_t bytes, unsigned char *p) {
	std::lock_guard<std::mutex> lk(g_stack_mu);
	if (g_stack_pool.size() < 96) {
		g_stack_pool.emplace_back(bytes, p);
	} else {
		munmap(p, bytes);
	}
}

static std::mutex g_report_mu;
static void deadlock_report(Cta &c) {
	std::unique_lock<std::mutex> lk(g_report_mu);
	fprintf(stderr, "[emu] kernel %s: CTA %u of %u made no progress for too long; threads still alive:\n", c.kernel, c.info.block, c.info.grid);
	// by warp: what its lanes wait for
	for (size_t w = 0; w < c.warps.size(); w++) {
		const char *what[32];
		int n = 0;
		for (int l = 0; l < 32; l++) {
			const size_t t = w * 32 + l;
			if (t < c.fibers.size() && !c.fibers[t].done) {
				what[n++] = c.fibers[t].waiting ? c.fibers[t].waiting : "spinning on memory";
			}
		}
		if (n == 0) {
			continue;
		}
		bool same = true;
		for (int i = 1; i < n; i++) {
			same = same && what[i] == what[0];
		}
		if (same) {
			fprintf(stderr, "    warp %zu: %d lanes: %s\n", w, n, what[0]);
		} else {
			for (int l = 0; l < 32; l++) {
				const size_t t = w * 32 + l;
				if (t < c.fibers.size() && !c.fibers[t].done) {
					const Fiber &f = c.fibers[t];
					if (f.waiting && !strcmp(f.waiting, "warp collective")) {
						fprintf(stderr, "    warp %zu lane %d: warp collective (mask %08x, op %d)\n", w, l, f.wait_mask, f.wait_op);
					} else {
						fprintf(stderr, "    warp %zu lane %d: %s\n", w, l, f.waiting ? f.waiting : "spinning on memory");
					}
				}
			}
		}
	}
	fflush(stderr);
	lk.unlock();
	struct timespec ts = { 3, 0 }; // the other CTAs of the grid get to report too
	nanosleep(&ts, nullptr);
	abort();
}

static void thread_altstack();
// Runs one CTA to completion on the calling OS thread.
static void run_cta(const char *name, unsigned block, unsigned grid, unsigned nthreads, size_t smem, const std::function<void()> &body) {
	static const double deadlock_s = (double)env_int("GAS_EMU_DEADLOCK_S", 60);
	Cta c;
	c.info.block = block;
	c.info.grid = grid;
	c.info.nthreads = nthreads;
	std::unique_ptr<unsigned char[]> dyn(new unsigned char[smem + 256]);
	c.info.dyn_smem = (unsigned char *)(((uintptr_t)dyn.get() + 127) & ~(uintptr_t)127);
	memset(c.info.dyn_smem, 0xcd, smem); // shared memory starts out as garbage
	c.body = &body;
	c.kernel = name;
	c.exited = 0;
	c.progress = false;
	c.cur = nullptr;
	memset(c.bars, 0, sizeof(c.bars));
	c.fibers.resize(nthreads);
	const unsigned nwarps = (nthreads + 31) / 32;
	c.warps.resize(nwarps);
	for (auto &w : c.warps) {
		memset(&w, 0, sizeof(w));
	}
	if (nthreads & 31) {
		c.warps[nwarps - 1].gone = ~0u << (nthreads & 31); // lanes that do not exist
	}
	const size_t stack_bytes = (size_t)nthreads * kStackBytes;
	unsigned char *stacks = stacks_get(stack_bytes);
	for (unsigned t = 0; t < nthreads; t++) {
		Fiber &f = c.fibers[t];
		f.info.tid = t;
		f.info.lane = t & 31;
		f.info.warp = t >> 5;
		f.done = false;
		f.waiting = nullptr;
		// initial frame: six callee-saved registers, then the entry point as return address; the stack pointer is 8 mod 16 at entry
		uintptr_t top = (uintptr_t)(stacks + (size_t)(t + 1) * kStackBytes);
		top &= ~(uintptr_t)15;
		void **sp = (void **)(top - 8);
		*sp = nullptr; // fiber_entry's "return address": ends a backtrace
		*--sp = (void *)fiber_entry;
		for (int i = 0; i < 6; i++) {
			*--sp = nullptr;
		}
		f.sp = sp;
	}
	thread_altstack();
	Cta *saved_self = t_self;
	CtaInfo *saved_cta = t_cta;
	FiberInfo *saved_fiber = t_fiber;
	t_self = &c;
	t_cta = &c.info;
	unsigned long long idle_since = 0;
	while (c.exited < nthreads) {
		c.progress = false;
		for (unsigned t = 0; t < nthreads; t++) {
			Fiber &f = c.fibers[t];
			if (f.done) {
				continue;
			}
			c.cur = &f;
			t_fiber = &f.info;
			emu_switch(&c.sched_sp, f.sp);
		}
		if (c.progress) {
			idle_since = 0;
		} else {
			// everybody waits for something outside this CTA (another CTA, or nothing at all)
			sched_yield();
			const unsigned long long now = now_ns();
			if (!idle_since) {
				idle_since = now;
			} else if ((double)(now - idle_since) * 1e-9 > deadlock_s) {
				deadlock_report(c);
			}
		}
	}
	t_self = saved_self;
	t_cta = saved_cta;
	t_fiber = saved_fiber;
	stacks_put(stack_bytes, stacks);
}

static std::atomic<unsigned long long> g_grids{ 0 }, g_ctas{ 0 };

// a fault inside an emulated kernel: say which kernel / CTA / thread and where (offsets for addr2line -e libgas_b200_emu.so)
static void on_fault(int sig, siginfo_t *si, void *) {
	static std::atomic<int> once{ 0 };
	if (once.fetch_add(1) == 0) {
		Cta *c = t_self;
		fprintf(stderr, "[emu] signal %d at address %p", sig, si ? si->si_addr : nullptr);
		if (c && c->cur) {
			fprintf(stderr, " in kernel %s, CTA %u of %u, thread %u", c->kernel, c->info.block, c->info.grid, c->cur->info.tid);
		}
		fprintf(stderr, "\n");
		void *bt[48];
		const int n = backtrace(bt, 48);
		backtrace_symbols_fd(bt, n, 2);
	}
	signal(sig, SIG_DFL);
	raise(sig);
}
static void thread_altstack() { // per OS thread: a fault on an overflowed fiber stack still gets reported
	static thread_local unsigned char *alt = nullptr;
	if (!alt) {
		alt = new unsigned char[1 << 16];
		stack_t ss{};
		ss.ss_sp = alt;
		ss.ss_size = 1 << 16;
		sigaltstack(&ss, nullptr);
	}
}
static void install_fault_handler() {
	static std::once_flag f;
	std::call_once(f, [] {
		struct sigaction sa{};
		sa.sa_sigaction = on_fault;
		sa.sa_flags = SA_SIGINFO | SA_ONSTACK | SA_NODEFER;
		if (env_int("GAS_EMU_FAULT_HANDLER", 1)) {
			sigaction(SIGSEGV, &sa, nullptr);
			sigaction(SIGBUS, &sa, nullptr);
		}
	});
}

static void run_grid(const char *name, dim3 grid, dim3 block, size_t smem, const std::function<void()> &body) {
	const unsigned n = grid.x * grid.y * grid.z, nt = block.x * block.y * block.z;
	if (n == 0 || nt == 0 || nt > 1024) {
		return;
	}
	install_fault_handler();
	static const int trace = env_int("GAS_EMU_TRACE", 0);
	if (trace) {
		fprintf(stderr, "[emu] grid %s <<<%u, %u, %zu>>>\n", name, n, nt, smem);
	}
	g_grids++;
	g_ctas += n;
	if (n == 1) {
		run_cta(name, 0, 1, nt, smem, body);
		return;
	}
	static const unsigned pool = (unsigned)std::max(1, env_int("GAS_EMU_THREADS", (int)std::max(2u, std::thread::hardware_concurrency())));
	const unsigned workers = n <= 64 ? n : std::min(n, pool);
	std::atomic<unsigned> next{ 0 };
	auto work = [&]() {
		for (;;) {
			const unsigned b = next.fetch_add(1);
			if (b >= n) {
				break;
			}
			run_cta(name, b, n, nt, smem, body);
		}
	};
	std::vector<std::thread> th;
	th.reserve(workers);
	for (unsigned i = 0; i < workers; i++) {
		th.emplace_back(work);
	}
	for (auto &t : th) {
		t.join();
	}
}

// ---- streams, events, capture --------------------------------------------------------------------------------------------------
struct Capture {
	std::vector<std::function<void()>> nodes;
	std::vector<Stream *> joined;
	bool invalid = false;
};
struct Stream {
	Capture *cap = nullptr;
};
struct Event {
	unsigned long long ns = 0;
	Capture *cap = nullptr; // last recorded inside this capture
	bool recorded = false;
};
struct Graph {
	std::vector<std::function<void()>> nodes;
};
static thread_local Capture *t_capturing = nullptr; // capture begun by this thread (cudaStreamCaptureModeThreadLocal)
static thread_local cudaError_t t_last_error = cudaSuccess;

static cudaError_t fail(cudaError_t e) {
	t_last_error = e;
	return e;
}
static cudaError_t unsafe_call(const char *what) {
	if (t_capturing) {
		fprintf(stderr, "[emu] %s while this thread is capturing a stream: the runtime refuses it and invalidates the capture\n", what);
		t_capturing->invalid = true;
		return fail(cudaErrorStreamCaptureUnsupported);
	}
	return cudaSuccess;
}

cudaError_t enqueue_grid(const char *name, dim3 grid, dim3 block, size_t smem, cudaStream_t st, std::function<void()> body) {
	if (block.x * block.y * block.z > 1024 || smem > 227 * 1024) {
		return fail(cudaErrorInvalidConfiguration);
	}
	if (st && st->cap) {
		st->cap->nodes.emplace_back([name, grid, block, smem, body]() { run_grid(name, grid, block, smem, body); });
		return cudaSuccess;
	}
	run_grid(name, grid, block, smem, body);
	return cudaSuccess;
}

} // namespace emu

using namespace emu;

const char *cudaGetErrorString(cudaError_t e) {
	switch (e) {
		case cudaSuccess: return "no error";
		case cudaErrorInvalidValue: return "invalid argument";
		case cudaErrorMemoryAllocation: return "out of memory";
		case cudaErrorInvalidConfiguration: return "invalid configuration argument";
		case cudaErrorNotSupported: return "operation not supported (emulation)";
		case cudaErrorStreamCaptureUnsupported: return "operation not permitted when stream is capturing";
		case cudaErrorStreamCaptureInvalidated: return "operation failed due to a previous error during capture";
		case cudaErrorCapturedEvent: return "operation not permitted on an event last recorded in a capturing stream";
		default: return "unknown error";
	}
}
cudaError_t cudaGetLastError() {
	const cudaError_t e = t_last_error;
	t_last_error = cudaSuccess;
	return e;
}
static int emulated_devices() {
	static const int n = std::max(1, std::min(8, env_int("GAS_EMU_DEVICES", 1)));
	return n;
}
cudaError_t cudaGetDeviceCount(int *n) {
	*n = emulated_devices();
	return cudaSuccess;
}
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
	memset(p, 0, sizeof(*p));
	snprintf(p->name, sizeof(p->name), "emulated sm_100 (tests/emu, %d SMs)", emulated_sms());
	p->major = 10;
	p->minor = 0;
	p->multiProcessorCount = emulated_sms();
	p->l2CacheSize = 4 << 20;
	p->totalGlobalMem = (size_t)8 << 30;
	p->sharedMemPerBlockOptin = 227 * 1024;
	return cudaSuccess;
}
cudaError_t cudaSetDevice(int dev) { return dev >= 0 && dev < emulated_devices() ? cudaSuccess : fail(cudaErrorInvalidValue); }
cudaError_t cudaDeviceSynchronize() { return unsafe_call("cudaDeviceSynchronize"); }
// GAS_EMU_IPC=1 (several emulated ranks, one process each): device allocations are shared mappings of memory files, so that
// cudaIpcGetMemHandle / cudaIpcOpenMemHandle can hand them to the other processes (through /proc/<pid>/fd/<fd>)
struct SharedAlloc {
	int fd;
	size_t bytes;
};
static std::mutex g_alloc_mu;
static std::map<void *, SharedAlloc> g_shared, g_opened;
static bool ipc_mode() {
	static const bool on = env_int("GAS_EMU_IPC", 0) != 0;
	return on;
}
cudaError_t cudaMalloc(void **p, size_t bytes) {
	if (cudaError_t e = unsafe_call("cudaMalloc")) {
		return e;
	}
	void *q = nullptr;
	const size_t n = bytes ? bytes : 1;
	if (ipc_mode()) {
		const int fd = memfd_create("gas_emu_alloc", 0);
		if (fd < 0 || ftruncate(fd, (off_t)n) != 0) {
			return fail(cudaErrorMemoryAllocation);
		}
		q = mmap(nullptr, n, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
		if (q == MAP_FAILED) {
			close(fd);
			return fail(cudaErrorMemoryAllocation);
		}
		std::lock_guard<std::mutex> lk(g_alloc_mu);
		g_shared[q] = SharedAlloc{ fd, n };
	} else if (posix_memalign(&q, 256, n) != 0) {
		return fail(cudaErrorMemoryAllocation);
	}
	// like device memory, allocations are not zeroed: fill with a pattern that shows up when read before it is written
	if (bytes <= ((size_t)64 << 20)) {
		memset(q, 0xa5, bytes);
	}
	*p = q;
	return cudaSuccess;
}
cudaError_t cudaFree(void *p) {
	if (cudaError_t e = unsafe_call("cudaFree")) {
		return e;
	}
	if (ipc_mode()) {
		std::lock_guard<std::mutex> lk(g_alloc_mu);
		auto it = g_shared.find(p);
		if (it != g_shared.end()) {
			munmap(p, it->second.bytes);
			close(it->second.fd);
			g_shared.erase(it);
		}
		return cudaSuccess;
	}
	free(p);
	return cudaSuccess;
}
cudaError_t cudaMemset(void *p, int v, size_t bytes) {
	if (cudaError_t e = unsafe_call("cudaMemset")) {
		return e;
	}
	memset(p, v, bytes);
	return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void *p, int v, size_t bytes, cudaStream_t st) {
	if (st && st->cap) {
		st->cap->nodes.emplace_back([p, v, bytes]() { memset(p, v, bytes); });
		return cudaSuccess;
	}
	memset(p, v, bytes);
	return cudaSuccess;
}
cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind) {
	if (cudaError_t e = unsafe_call("cudaMemcpy")) {
		return e;
	}
	memmove(dst, src, bytes);
	return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind, cudaStream_t st) {
	if (st && st->cap) {
		st->cap->nodes.emplace_back([dst, src, bytes]() { memmove(dst, src, bytes); });
		return cudaSuccess;
	}
	memmove(dst, src, bytes);
	return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *st, unsigned) {
	*st = new Stream();
	return cudaSuccess;
}
cudaError_t cudaStreamDestroy(cudaStream_t st) {
	delete st;
	return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t st) {
	if (st && st->cap) {
		fprintf(stderr, "[emu] cudaStreamSynchronize on a capturing stream\n");
		st->cap->invalid = true;
		return fail(cudaErrorStreamCaptureUnsupported);
	}
	return cudaSuccess;
}
cudaError_t cudaStreamWaitEvent(cudaStream_t st, cudaEvent_t ev, unsigned) {
	if (!st || !ev) {
		return fail(cudaErrorInvalidValue);
	}
	if (ev->cap) {
		if (!st->cap) { // the stream joins the capture the event was recorded in
			st->cap = ev->cap;
			ev->cap->joined.push_back(st);
		} else if (st->cap != ev->cap) {
			return fail(cudaErrorStreamCaptureUnsupported);
		}
	}
	return cudaSuccess; // everything enqueued earlier has already run
}
cudaError_t cudaStreamBeginCapture(cudaStream_t st, cudaStreamCaptureMode) {
	if (!st || st->cap || t_capturing) {
		return fail(cudaErrorInvalidValue);
	}
	st->cap = new Capture();
	t_capturing = st->cap;
	return cudaSuccess;
}
cudaError_t cudaStreamEndCapture(cudaStream_t st, cudaGraph_t *graph) {
	if (!st || !st->cap || st->cap != t_capturing) {
		return fail(cudaErrorInvalidValue);
	}
	Capture *c = st->cap;
	for (Stream *j : c->joined) {
		j->cap = nullptr;
	}
	st->cap = nullptr;
	t_capturing = nullptr;
	cudaError_t e = cudaSuccess;
	if (c->invalid) {
		*graph = nullptr;
		e = fail(cudaErrorStreamCaptureInvalidated);
	} else {
		Graph *g = new Graph();
		g->nodes = std::move(c->nodes);
		*graph = g;
	}
	// events recorded inside the capture stay "captured" until they are recorded again; they must not dangle
	c->nodes.clear();
	c->joined.clear();
	// (the Capture object is leaked on purpose: captured events keep pointing at it, and a later wait on one must not join anything)
	c->invalid = true;
	return e;
}
cudaError_t cudaEventCreate(cudaEvent_t *ev) {
	*ev = new Event();
	return cudaSuccess;
}
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *ev, unsigned) { return cudaEventCreate(ev); }
cudaError_t cudaEventDestroy(cudaEvent_t ev) {
	delete ev;
	return cudaSuccess;
}
cudaError_t cudaEventRecordWithFlags(cudaEvent_t ev, cudaStream_t st, unsigned flags) {
	if (!ev) {
		return fail(cudaErrorInvalidValue);
	}
	if (st && st->cap) {
		if (flags & cudaEventRecordExternal) { // an event-record node: stamps the event when the graph runs
			st->cap->nodes.emplace_back([ev]() {
				ev->ns = now_ns();
				ev->recorded = true;
			});
			ev->cap = nullptr;
		} else {
			ev->cap = st->cap;
		}
		return cudaSuccess;
	}
	ev->cap = nullptr;
	ev->ns = now_ns();
	ev->recorded = true;
	return cudaSuccess;
}
cudaError_t cudaEventRecord(cudaEvent_t ev, cudaStream_t st) { return cudaEventRecordWithFlags(ev, st, cudaEventRecordDefault); }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
	if (!a || !b || !a->recorded || !b->recorded) {
		return fail(cudaErrorInvalidValue);
	}
	if (a->cap || b->cap) {
		return fail(cudaErrorCapturedEvent);
	}
	*ms = (float)((double)(b->ns - a->ns) * 1e-6);
	return cudaSuccess;
}
cudaError_t cudaGraphInstantiate(cudaGraphExec_t *exec, cudaGraph_t graph, unsigned long long) {
	if (!graph) {
		return fail(cudaErrorInvalidValue);
	}
	Graph *g = new Graph();
	g->nodes = graph->nodes;
	*exec = g;
	return cudaSuccess;
}
cudaError_t cudaGraphDestroy(cudaGraph_t graph) {
	delete graph;
	return cudaSuccess;
}
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t exec) {
	delete exec;
	return cudaSuccess;
}
cudaError_t cudaGraphLaunch(cudaGraphExec_t exec, cudaStream_t st) {
	if (!exec || (st && st->cap)) {
		return fail(cudaErrorInvalidValue);
	}
	for (auto &n : exec->nodes) {
		n();
	}
	return cudaSuccess;
}
struct IpcHandle {
	int pid, fd;
	unsigned long long bytes;
};
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) {
	memset(h, 0, sizeof(*h));
	std::lock_guard<std::mutex> lk(g_alloc_mu);
	auto it = g_shared.find(p);
	if (it == g_shared.end()) {
		return fail(cudaErrorNotSupported); // not in IPC mode (GAS_EMU_IPC=1), or not the base of an allocation
	}
	IpcHandle ih{ (int)getpid(), it->second.fd, it->second.bytes };
	memcpy(h->reserved, &ih, sizeof(ih));
	return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) {
	IpcHandle ih;
	memcpy(&ih, h.reserved, sizeof(ih));
	char path[64];
	snprintf(path, sizeof(path), "/proc/%d/fd/%d", ih.pid, ih.fd);
	const int fd = open(path, O_RDWR);
	if (fd < 0) {
		return fail(cudaErrorInvalidValue);
	}
	void *q = mmap(nullptr, (size_t)ih.bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
	close(fd);
	if (q == MAP_FAILED) {
		return fail(cudaErrorMemoryAllocation);
	}
	std::lock_guard<std::mutex> lk(g_alloc_mu);
	g_opened[q] = SharedAlloc{ -1, (size_t)ih.bytes };
	*p = q;
	return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void *p) {
	std::lock_guard<std::mutex> lk(g_alloc_mu);
	auto it = g_opened.find(p);
	if (it == g_opened.end()) {
		return fail(cudaErrorInvalidValue);
	}
	munmap(p, it->second.bytes);
	g_opened.erase(it);
	return cudaSuccess;
}

// ---- statistics for the tests ----------------------------------------------------------------------------------------------------
extern "C" __attribute__((visibility("default"))) void gas_emu_stats(unsigned long long *grids, unsigned long long *ctas) {
	*grids = emu::g_grids.load();
	*ctas = emu::g_ctas.load();
}
