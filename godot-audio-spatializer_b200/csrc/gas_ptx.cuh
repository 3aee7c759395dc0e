// gas_ptx.cuh — every inline-PTX instruction the kernels use, one small inline function each (sm_100a).
// Product code.  Having them in one place has a second use: tests/emu/ compiles the kernels' SOURCE with g++ and runs them on the
// CPU (GAS_KERNEL_EMULATION, test infrastructure only, never part of libgas_b200.so); there this header hands over to
// tests/emu/gas_ptx_emu.h, which restates each function in C++.
#pragma once

#ifdef GAS_KERNEL_EMULATION
#include "gas_ptx_emu.h"
#else

#include <stdint.h>

// ---- timers / ids ------------------------------------------------------------------------------------------------------
static __device__ __forceinline__ unsigned long long gas_globaltimer() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}
static __device__ __forceinline__ unsigned gas_smid() {
	unsigned smid;
	asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
	return smid;
}

// ---- shared-memory addresses -------------------------------------------------------------------------------------------
typedef uint32_t gas_smem_addr; // 32-bit shared-window address
static __device__ __forceinline__ gas_smem_addr gas_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// dynamic shared memory of the kernel: GAS_DYN_SMEM(float, 16, s_tile) declares `float s_tile[]`
#define GAS_DYN_SMEM(type_, align_, name_) extern __shared__ __align__(align_) type_ name_[]

// ---- programmatic dependent launch -------------------------------------------------------------------------------------
#define GAS_GRID_DEP_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#define GAS_GRID_DEP_LAUNCH() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")

// ---- mbarriers + 1-D bulk async copies (TMA engine) ----------------------------------------------------------------------
static __device__ __forceinline__ void gas_mbar_init(uint64_t *bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gas_smem_u32(bar)), "r"(count));
}
static __device__ __forceinline__ void gas_mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
static __device__ __forceinline__ void gas_mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gas_smem_u32(bar)), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void gas_mbar_arrive(uint64_t *bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(gas_smem_u32(bar)) : "memory");
}
static __device__ __forceinline__ void gas_mbar_wait(uint64_t *bar, uint32_t parity) {
	uint32_t ok = 0;
	do {
		asm volatile(
				"{\n"
				".reg .pred p;\n"
				"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
				"selp.u32 %0, 1, 0, p;\n"
				"}\n"
				: "=r"(ok)
				: "r"(gas_smem_u32(bar)), "r"(parity)
				: "memory");
	} while (!ok);
}
// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
static __device__ __forceinline__ void gas_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(gas_smem_u32(dst)),
			"l"(src), "r"(bytes), "r"(gas_smem_u32(bar))
			: "memory");
}
static __device__ __forceinline__ uint64_t gas_l2_policy_evict_first() {
	uint64_t p;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
	return p;
}
static __device__ __forceinline__ void gas_bulk_g2s_hint(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(gas_smem_u32(dst)),
			"l"(src), "r"(bytes), "r"(gas_smem_u32(bar)), "l"(policy)
			: "memory");
}

// ---- per-thread async copies global -> shared (SASS: LDGSTS) ---------------------------------------------------------------
static __device__ __forceinline__ void gas_cp_async_16(void *smem_dst, const void *gsrc) {
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(gas_smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
static __device__ __forceinline__ void gas_cp_async_8(void *smem_dst, const void *gsrc) { // (8-byte copies exist in the .ca form only)
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(gas_smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
// all of this thread's async copies so far have landed in shared memory (visible to the thread; a barrier publishes them to the CTA)
static __device__ __forceinline__ void gas_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- named barriers ----------------------------------------------------------------------------------------------------
static __device__ __forceinline__ void gas_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
#define GAS_BAR_SYNC_IMM(id_, n_) asm volatile("bar.sync " #id_ ", %0;" ::"n"(n_) : "memory")
#define GAS_BAR_ARRIVE_IMM(id_, n_) asm volatile("bar.arrive " #id_ ", %0;" ::"n"(n_) : "memory")

// ---- reductions / ordered accesses ---------------------------------------------------------------------------------------
static __device__ __forceinline__ void gas_red_add_v4(float *addr, float a, float b, float c, float d) {
	asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
static __device__ __forceinline__ void gas_red_shared_add_f32(gas_smem_addr addr, float v) {
	asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
static __device__ __forceinline__ void gas_red_release_gpu_add_s32(int32_t *p, int v) {
	asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
static __device__ __forceinline__ int gas_atom_add_acq_rel_gpu_s32(int32_t *p, int v) {
	int old;
	asm volatile("atom.add.acq_rel.gpu.global.s32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
	return old;
}
static __device__ __forceinline__ void gas_st_release_gpu_s32(int32_t *p, int v) {
	asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
static __device__ __forceinline__ int gas_ld_acquire_gpu_s32(const int32_t *p) {
	int v;
	asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
static __device__ __forceinline__ void gas_red_release_sys_add_u64(unsigned long long *p, unsigned long long v) {
	asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
static __device__ __forceinline__ unsigned long long gas_ld_acquire_sys_u64(const unsigned long long *p) {
	unsigned long long v;
	asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}

// ---- packed (L,R) FMA: a * b + c on both halves with one instruction (SASS: FFMA2) ----------------------------------------
#ifndef GAS_USE_FFMA2
#define GAS_USE_FFMA2 1
#endif
static __device__ __forceinline__ float2 gas_ffma2(const float2 a, const float2 b, const float2 c) {
	float2 d;
#if GAS_USE_FFMA2
	asm("{\n"
		".reg .b64 ra, rb, rc, rd;\n"
		"mov.b64 ra, {%2, %3};\n"
		"mov.b64 rb, {%4, %5};\n"
		"mov.b64 rc, {%6, %7};\n"
		"fma.rn.f32x2 rd, ra, rb, rc;\n"
		"mov.b64 {%0, %1}, rd;\n"
		"}\n"
		: "=f"(d.x), "=f"(d.y)
		: "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
#else
	d.x = fmaf(a.x, b.x, c.x);
	d.y = fmaf(a.y, b.y, c.y);
#endif
	return d;
}

#endif // GAS_KERNEL_EMULATION
