// gas_ptx_emu.h — TEST INFRASTRUCTURE: C++ restatements of the PTX helpers of csrc/gas_ptx.cuh for the CPU emulation (tests/emu).
// Each function does what the instruction it stands for is documented to do, on top of the fiber engine of emu_core.cpp.
#pragma once
#include <cuda_runtime.h> // the stand-in of this directory

static inline unsigned long long gas_globaltimer() { return emu::now_ns(); }
static inline unsigned gas_smid() { return emu::t_cta->block % (unsigned)emu::emulated_sms(); }

typedef uintptr_t gas_smem_addr;
static inline gas_smem_addr gas_smem_u32(const void *p) { return (gas_smem_addr)p; }
#define GAS_DYN_SMEM(type_, align_, name_) type_ *name_ = reinterpret_cast<type_ *>(emu::t_cta->dyn_smem)

// programmatic dependent launch: the emulated stream runs its kernels one after the other, so the wait is already satisfied
#define GAS_GRID_DEP_WAIT() ((void)0)
#define GAS_GRID_DEP_LAUNCH() ((void)0)

// ---- mbarrier: phase bit, arrival count, pending arrivals, pending transaction bytes in the barrier's 64 bits -----------------
namespace emu {
struct MBar {
	uint32_t phase : 1;
	uint32_t expected : 15;
	uint32_t pending : 16;
	int32_t tx;
};
static_assert(sizeof(MBar) == 8, "an mbarrier object is 64 bits");
static inline void mbar_check(MBar *b) {
	if (b->pending == 0 && b->tx == 0) {
		b->phase ^= 1u;
		b->pending = b->expected;
	}
}
} // namespace emu
static inline void gas_mbar_init(uint64_t *bar, uint32_t count) {
	emu::MBar *b = reinterpret_cast<emu::MBar *>(bar);
	b->phase = 0;
	b->expected = count;
	b->pending = count;
	b->tx = 0;
}
static inline void gas_mbar_init_fence() {}
static inline void gas_mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
	emu::MBar *b = reinterpret_cast<emu::MBar *>(bar);
	b->tx += (int32_t)bytes;
	b->pending -= 1;
	emu::mbar_check(b);
}
static inline void gas_mbar_arrive(uint64_t *bar) {
	emu::MBar *b = reinterpret_cast<emu::MBar *>(bar);
	b->pending -= 1;
	emu::mbar_check(b);
}
static inline void gas_mbar_wait(uint64_t *bar, uint32_t parity) {
	// try_wait.parity succeeds once the phase with that parity has completed, i.e. the barrier is in the other phase
	const volatile emu::MBar *b = reinterpret_cast<const volatile emu::MBar *>(bar);
	while (b->phase == (parity & 1u)) {
		emu::yield_blocked();
	}
}
static inline void gas_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
	// cp.async.bulk requires 16-byte aligned addresses and a size that is a multiple of 16: a violation faults on the device
	if ((((uintptr_t)dst | (uintptr_t)src | bytes) & 15u) != 0) {
		fprintf(stderr, "[emu] cp.async.bulk: misaligned copy (dst %p, src %p, %u bytes)\n", dst, src, bytes);
		abort();
	}
	memcpy(dst, src, bytes);
	emu::MBar *b = reinterpret_cast<emu::MBar *>(bar);
	b->tx -= (int32_t)bytes;
	emu::mbar_check(b);
}
static inline uint64_t gas_l2_policy_evict_first() { return 0; }
static inline void gas_bulk_g2s_hint(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t) { gas_bulk_g2s(dst, src, bytes, bar); }

// ---- per-thread async copies (cp.async): executed at issue time; the source is never written while a copy is pending --------
static inline void gas_cp_async_16(void *smem_dst, const void *gsrc) {
	if ((((uintptr_t)smem_dst | (uintptr_t)gsrc) & 15u) != 0) { // a misaligned cp.async faults on the device
		fprintf(stderr, "[emu] cp.async 16: misaligned copy (dst %p, src %p)\n", smem_dst, gsrc);
		abort();
	}
	memcpy(smem_dst, gsrc, 16);
}
static inline void gas_cp_async_8(void *smem_dst, const void *gsrc) {
	if ((((uintptr_t)smem_dst | (uintptr_t)gsrc) & 7u) != 0) {
		fprintf(stderr, "[emu] cp.async 8: misaligned copy (dst %p, src %p)\n", smem_dst, gsrc);
		abort();
	}
	memcpy(smem_dst, gsrc, 8);
}
static inline void gas_cp_async_wait_all() {}

// ---- named barriers ----------------------------------------------------------------------------------------------------
static inline void gas_bar_sync(int id, int nthreads) { emu::bar_sync(id, nthreads); }
#define GAS_BAR_SYNC_IMM(id_, n_) emu::bar_sync((id_), (n_))
#define GAS_BAR_ARRIVE_IMM(id_, n_) emu::bar_arrive((id_), (n_))

// ---- reductions / ordered accesses ---------------------------------------------------------------------------------------
static inline void gas_red_add_v4(float *addr, float a, float b, float c, float d) {
	if (((uintptr_t)addr & 15u) != 0) {
		fprintf(stderr, "[emu] red.global.add.v4.f32: misaligned address %p\n", (void *)addr);
		abort();
	}
	atomicAdd(addr + 0, a);
	atomicAdd(addr + 1, b);
	atomicAdd(addr + 2, c);
	atomicAdd(addr + 3, d);
}
static inline void gas_red_shared_add_f32(gas_smem_addr addr, float v) { *reinterpret_cast<float *>(addr) += v; } // (a CTA's threads share one OS thread)
static inline void gas_red_release_gpu_add_s32(int32_t *p, int v) { __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int gas_atom_add_acq_rel_gpu_s32(int32_t *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline void gas_st_release_gpu_s32(int32_t *p, int v) { __atomic_store_n(p, v, __ATOMIC_SEQ_CST); }
static inline int gas_ld_acquire_gpu_s32(const int32_t *p) { return __atomic_load_n(p, __ATOMIC_SEQ_CST); }
static inline void gas_red_release_sys_add_u64(unsigned long long *p, unsigned long long v) { __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long gas_ld_acquire_sys_u64(const unsigned long long *p) {
	const unsigned long long v = __atomic_load_n(p, __ATOMIC_SEQ_CST);
	emu::yield_blocked(); // (only ever used in spin loops)
	return v;
}

static inline float2 gas_ffma2(const float2 a, const float2 b, const float2 c) {
	float2 d;
	d.x = fmaf(a.x, b.x, c.x);
	d.y = fmaf(a.y, b.y, c.y);
	return d;
}
