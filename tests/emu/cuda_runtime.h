// cuda_runtime.h (tests/emu) — TEST INFRASTRUCTURE.  A stand-in for the CUDA runtime header that lets g++ compile the product's
// .cu files (godot-audio-spatializer_b200/csrc) unchanged and run them on the CPU:
//   * the device vocabulary (threadIdx, __syncthreads, warp shuffles, atomics, __shared__, ...) on top of a fiber-per-CUDA-thread
//     engine (emu_core.cpp): every CTA is one OS thread that schedules its CUDA threads cooperatively; CTAs run concurrently;
//   * the slice of the runtime API the library calls (memory, streams, events, stream capture into graphs, launches), executing
//     every operation at enqueue time (a valid schedule of any correctly ordered program) and replaying captured graphs in order.
// It exists so that the GPU test-suite, smoke() and bench.py's control flow can be executed in a container without a GPU.  It says
// nothing about performance and it is never part of libgas_b200.so.
#pragma once

// system headers first: the qualifier macros below must not leak into them
#include <math.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <new>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#define GAS_KERNEL_EMULATION 1

// ---- qualifiers ----------------------------------------------------------------------------------------------------------
#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))
// (16-byte aligned like nvcc lays out shared arrays: the kernels fill some of them with 128-bit stores)
#define __shared__ static thread_local __attribute__((aligned(16)))

// ---- vector types ----------------------------------------------------------------------------------------------------------
struct __attribute__((aligned(8))) float2 {
	float x, y;
};
struct float3 {
	float x, y, z;
};
struct __attribute__((aligned(16))) float4 {
	float x, y, z, w;
};
struct __attribute__((aligned(8))) int2 {
	int x, y;
};
struct __attribute__((aligned(16))) int4 {
	int x, y, z, w;
};
struct __attribute__((aligned(8))) uint2 {
	unsigned x, y;
};
struct __attribute__((aligned(16))) uint4 {
	unsigned x, y, z, w;
};
struct __attribute__((aligned(16))) double2 {
	double x, y;
};
static inline float2 make_float2(float x, float y) { return float2{ x, y }; }
static inline float3 make_float3(float x, float y, float z) { return float3{ x, y, z }; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{ x, y, z, w }; }
static inline int2 make_int2(int x, int y) { return int2{ x, y }; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{ x, y, z, w }; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{ x, y }; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{ x, y, z, w }; }
static inline double2 make_double2(double x, double y) { return double2{ x, y }; }

struct dim3 {
	unsigned x, y, z;
	dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

// ---- the engine's interface ------------------------------------------------------------------------------------------------
namespace emu {

struct FiberInfo { // what a CUDA thread knows about itself
	unsigned tid, lane, warp;
};
struct CtaInfo {
	unsigned block, nthreads, grid;
	unsigned char *dyn_smem;
};
extern thread_local FiberInfo *t_fiber;
extern thread_local CtaInfo *t_cta;

void yield_blocked(); // the calling CUDA thread cannot proceed: run the others
void syncthreads();
void bar_sync(int id, int nthreads);
void bar_arrive(int id, int nthreads);
enum { OP_SYNC = 0, OP_SHFL_IDX, OP_SHFL_XOR, OP_SHFL_UP, OP_SHFL_DOWN, OP_BALLOT, OP_REDUCE_ADD, OP_ANY, OP_ALL };
uint64_t warp_collective(unsigned mask, int op, int arg, int width, uint64_t value);
unsigned long long now_ns();
int emulated_sms();

template <typename T>
static inline uint64_t to_bits(T v) {
	static_assert(sizeof(T) <= 8, "shuffle payload");
	uint64_t b = 0;
	memcpy(&b, &v, sizeof(T));
	return b;
}
template <typename T>
static inline T from_bits(uint64_t b) {
	T v;
	memcpy(&v, &b, sizeof(T));
	return v;
}

} // namespace emu

// ---- built-in variables ------------------------------------------------------------------------------------------------------
namespace emu {
struct TidX {
	operator unsigned() const { return t_fiber->tid; }
};
struct BidX {
	operator unsigned() const { return t_cta->block; }
};
struct BdimX {
	operator unsigned() const { return t_cta->nthreads; }
};
struct GdimX {
	operator unsigned() const { return t_cta->grid; }
};
struct Zero {
	operator unsigned() const { return 0u; }
};
struct One {
	operator unsigned() const { return 1u; }
};
struct ThreadIdx {
	TidX x;
	Zero y, z;
};
struct BlockIdx {
	BidX x;
	Zero y, z;
};
struct BlockDim {
	BdimX x;
	One y, z;
};
struct GridDim {
	GdimX x;
	One y, z;
};
} // namespace emu
static const emu::ThreadIdx threadIdx{};
static const emu::BlockIdx blockIdx{};
static const emu::BlockDim blockDim{};
static const emu::GridDim gridDim{};
static const int warpSize = 32;

// ---- synchronisation and warp collectives ------------------------------------------------------------------------------------
static inline void __syncthreads() { emu::syncthreads(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::warp_collective(mask, emu::OP_SYNC, 0, 32, 0); }
template <typename T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
	return emu::from_bits<T>(emu::warp_collective(mask, emu::OP_SHFL_IDX, src, width, emu::to_bits(v)));
}
template <typename T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask, int width = 32) {
	return emu::from_bits<T>(emu::warp_collective(mask, emu::OP_SHFL_XOR, lane_mask, width, emu::to_bits(v)));
}
template <typename T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
	return emu::from_bits<T>(emu::warp_collective(mask, emu::OP_SHFL_UP, (int)delta, width, emu::to_bits(v)));
}
template <typename T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
	return emu::from_bits<T>(emu::warp_collective(mask, emu::OP_SHFL_DOWN, (int)delta, width, emu::to_bits(v)));
}
static inline unsigned __ballot_sync(unsigned mask, int pred) { return (unsigned)emu::warp_collective(mask, emu::OP_BALLOT, 0, 32, pred ? 1u : 0u); }
static inline int __any_sync(unsigned mask, int pred) { return (int)emu::warp_collective(mask, emu::OP_ANY, 0, 32, pred ? 1u : 0u); }
static inline int __all_sync(unsigned mask, int pred) { return (int)emu::warp_collective(mask, emu::OP_ALL, 0, 32, pred ? 1u : 0u); }
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) { return (unsigned)emu::warp_collective(mask, emu::OP_REDUCE_ADD, 0, 32, v); }
static inline int __reduce_add_sync(unsigned mask, int v) { return (int)(unsigned)emu::warp_collective(mask, emu::OP_REDUCE_ADD, 0, 32, (unsigned)v); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __nanosleep(unsigned) { emu::yield_blocked(); }

// ---- bit / conversion intrinsics -----------------------------------------------------------------------------------------------
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline unsigned __float_as_uint(float f) { return emu::from_bits<unsigned>(emu::to_bits(f)); }
static inline int __float_as_int(float f) { return emu::from_bits<int>(emu::to_bits(f)); }
static inline float __uint_as_float(unsigned u) { return emu::from_bits<float>(emu::to_bits(u)); }
static inline float __int_as_float(int i) { return emu::from_bits<float>(emu::to_bits(i)); }
static inline unsigned __double2uint_rz(double d) { return d <= 0.0 ? 0u : (d >= 4294967295.0 ? 0xffffffffu : (unsigned)d); } // (C truncates like rz)
static inline float __fmul_rn(float a, float b) {
	volatile float r = a * b;
	return r;
}
static inline float __fdiv_rn(float a, float b) {
	volatile float r = a / b;
	return r;
}
static inline float __fadd_rn(float a, float b) {
	volatile float r = a + b;
	return r;
}
template <typename T>
static inline T __ldcg(const T *p) {
	T v;
	asm volatile("" ::: "memory"); // a fresh read every time (the kernels poll with it)
	memcpy(&v, p, sizeof(T));
	asm volatile("" ::: "memory");
	return v;
}
template <typename T>
static inline T __ldg(const T *p) {
	return *p;
}
static inline size_t __cvta_generic_to_shared(const void *p) { return (size_t)p; }

// min / max over mixed integer types like the CUDA headers provide
template <typename A, typename B>
static inline typename std::common_type<A, B>::type min(A a, B b) {
	typedef typename std::common_type<A, B>::type R;
	return (R)a < (R)b ? (R)a : (R)b;
}
template <typename A, typename B>
static inline typename std::common_type<A, B>::type max(A a, B b) {
	typedef typename std::common_type<A, B>::type R;
	return (R)a > (R)b ? (R)a : (R)b;
}

// ---- atomics (CTAs run on different OS threads) --------------------------------------------------------------------------------
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline float atomicAdd(float *p, float v) {
	uint32_t *ip = reinterpret_cast<uint32_t *>(p);
	uint32_t old = __atomic_load_n(ip, __ATOMIC_RELAXED);
	for (;;) {
		const float f = emu::from_bits<float>(old) + v;
		if (__atomic_compare_exchange_n(ip, &old, emu::from_bits<uint32_t>(emu::to_bits(f)), true, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED)) {
			return emu::from_bits<float>(old);
		}
	}
}
static inline int atomicCAS(int *p, int cmp, int v) {
	__atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
	return cmp;
}
static inline unsigned atomicCAS(unsigned *p, unsigned cmp, unsigned v) {
	__atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
	return cmp;
}
static inline unsigned long long atomicCAS(unsigned long long *p, unsigned long long cmp, unsigned long long v) {
	__atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
	return cmp;
}
static inline int atomicMin(int *p, int v) {
	int old = __atomic_load_n(p, __ATOMIC_RELAXED);
	while (old > v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED)) {
	}
	return old;
}
static inline int atomicMax(int *p, int v) {
	int old = __atomic_load_n(p, __ATOMIC_RELAXED);
	while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED)) {
	}
	return old;
}
static inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicExch(int *p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }

// ---- runtime API ---------------------------------------------------------------------------------------------------------------
typedef int cudaError_t;
enum {
	cudaSuccess = 0,
	cudaErrorInvalidValue = 1,
	cudaErrorMemoryAllocation = 2,
	cudaErrorInvalidConfiguration = 9,
	cudaErrorNotSupported = 801,
	cudaErrorStreamCaptureUnsupported = 900,
	cudaErrorStreamCaptureInvalidated = 901,
	cudaErrorCapturedEvent = 907
};
namespace emu {
struct Stream;
struct Event;
struct Graph;
} // namespace emu
typedef emu::Stream *cudaStream_t;
typedef emu::Event *cudaEvent_t;
typedef emu::Graph *cudaGraph_t;
typedef emu::Graph *cudaGraphExec_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1 };
enum { cudaEventDefault = 0, cudaEventDisableTiming = 2 };
enum { cudaEventRecordDefault = 0, cudaEventRecordExternal = 1 };
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal = 0, cudaStreamCaptureModeThreadLocal = 1, cudaStreamCaptureModeRelaxed = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum { cudaSharedmemCarveoutMaxShared = 100 };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
struct cudaIpcMemHandle_t {
	char reserved[64];
};
struct cudaDeviceProp {
	char name[256];
	int major, minor, multiProcessorCount, l2CacheSize;
	size_t totalGlobalMem, sharedMemPerBlockOptin;
};
enum cudaLaunchAttributeID { cudaLaunchAttributeProgrammaticStreamSerialization = 4, cudaLaunchAttributeProgrammaticEvent = 5 };
struct cudaLaunchAttributeValue {
	int programmaticStreamSerializationAllowed;
	struct {
		cudaEvent_t event;
		int flags;
		int triggerAtBlockStart;
	} programmaticEvent;
};
struct cudaLaunchAttribute {
	cudaLaunchAttributeID id;
	cudaLaunchAttributeValue val;
};
struct cudaLaunchConfig_t {
	dim3 gridDim, blockDim;
	size_t dynamicSmemBytes;
	cudaStream_t stream;
	cudaLaunchAttribute *attrs;
	unsigned numAttrs;
};

const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetLastError();
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int dev);
cudaError_t cudaSetDevice(int dev);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaMalloc(void **p, size_t bytes);
cudaError_t cudaFree(void *p);
cudaError_t cudaMemset(void *p, int v, size_t bytes);
cudaError_t cudaMemsetAsync(void *p, int v, size_t bytes, cudaStream_t st);
cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind);
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t st);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *st, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t st);
cudaError_t cudaStreamSynchronize(cudaStream_t st);
cudaError_t cudaStreamWaitEvent(cudaStream_t st, cudaEvent_t ev, unsigned flags);
cudaError_t cudaStreamBeginCapture(cudaStream_t st, cudaStreamCaptureMode mode);
cudaError_t cudaStreamEndCapture(cudaStream_t st, cudaGraph_t *graph);
cudaError_t cudaEventCreate(cudaEvent_t *ev);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *ev, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t ev);
cudaError_t cudaEventRecord(cudaEvent_t ev, cudaStream_t st);
cudaError_t cudaEventRecordWithFlags(cudaEvent_t ev, cudaStream_t st, unsigned flags);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaGraphInstantiate(cudaGraphExec_t *exec, cudaGraph_t graph, unsigned long long flags);
cudaError_t cudaGraphDestroy(cudaGraph_t graph);
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t exec);
cudaError_t cudaGraphLaunch(cudaGraphExec_t exec, cudaStream_t st);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p);
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void *p);
template <typename F>
static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) {
	return cudaSuccess;
}

namespace emu {
// enqueue a grid: runs it now, or records it when the stream is capturing
cudaError_t enqueue_grid(const char *name, dim3 grid, dim3 block, size_t smem, cudaStream_t st, std::function<void()> thread_body);

template <typename... KArgs, typename... Args>
static inline cudaError_t launch(const char *name, dim3 grid, dim3 block, size_t smem, cudaStream_t st, void (*kernel)(KArgs...), Args &&...args) {
	// arguments are evaluated and copied NOW (like a real launch), not when a captured graph replays
	std::tuple<typename std::decay<KArgs>::type...> packed(static_cast<KArgs>(args)...);
	return enqueue_grid(name, grid, block, smem, st, [kernel, packed]() { std::apply(kernel, packed); });
}
} // namespace emu

template <typename... KArgs, typename... Args>
static inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t *lc, void (*kernel)(KArgs...), Args &&...args) {
	return emu::launch("kernel", lc->gridDim, lc->blockDim, lc->dynamicSmemBytes, lc->stream, kernel, std::forward<Args>(args)...);
}
