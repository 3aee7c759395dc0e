// cuda_emu.h — TEST INFRASTRUCTURE: runs a CUDA kernel's SOURCE on the CPU, one std::thread per CUDA thread, one block at a time.
// Enough of the device vocabulary for element-wise / block-reduction kernels (no warp intrinsics): threadIdx / blockIdx / blockDim /
// gridDim, __syncthreads, __shared__ (function-local static: blocks run one after the other), atomicMin on int.  Written after the
// round's GPU budget was spent, to execute the kernels that could not be run on a B200 yet (gas_resample.cu, gas_bus.cu).
#pragma once
#include <cuda_runtime.h> // host-side declarations and vector types only (this file is compiled by g++, not nvcc)
#include "../../godot-audio-spatializer_b200/csrc/gas_internal.h" // before the macros below: it names gridDim / blockDim as struct fields

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace emu {

struct Dim {
	unsigned x = 1, y = 1, z = 1;
};
inline thread_local Dim t_threadIdx, t_blockIdx;
inline Dim g_blockDim, g_gridDim;

class Barrier {
public:
	explicit Barrier(int n) : n_(n) {}
	void wait() {
		std::unique_lock<std::mutex> lk(m_);
		const int gen = gen_;
		if (++count_ == n_) {
			count_ = 0;
			gen_++;
			cv_.notify_all();
		} else {
			cv_.wait(lk, [&] { return gen != gen_; });
		}
	}

private:
	std::mutex m_;
	std::condition_variable cv_;
	int n_, count_ = 0, gen_ = 0;
};
inline Barrier *g_barrier = nullptr;
inline std::mutex g_atomic_mutex;

inline int atomicMin_emu(int *p, int v) {
	std::lock_guard<std::mutex> lk(g_atomic_mutex);
	const int old = *p;
	*p = std::min(old, v);
	return old;
}

// Runs `body` for every thread of every block of a <<<grid, block>>> launch (1-D).  A thread that returns early simply ends: the
// kernels run here only return early block-uniformly or after their last barrier... except before a barrier, which would hang,
// so barriers are counted per block with the number of threads still alive.
inline void launch(unsigned grid, unsigned block, const std::function<void()> &body) {
	g_blockDim.x = block;
	g_gridDim.x = grid;
	for (unsigned b = 0; b < grid; b++) {
		Barrier bar((int)block);
		g_barrier = &bar;
		std::vector<std::thread> th;
		th.reserve(block);
		for (unsigned t = 0; t < block; t++) {
			th.emplace_back([&, t, b] {
				t_threadIdx.x = t;
				t_blockIdx.x = b;
				body();
			});
		}
		for (auto &x : th) {
			x.join();
		}
	}
	g_barrier = nullptr;
}

} // namespace emu

// ---- the device vocabulary, as macros so that they apply inside the included .cu text ------------------------------------
#undef __shared__
#define __shared__ static
#undef __global__
#define __global__
#undef __device__
#define __device__
#undef __forceinline__
#define __forceinline__ inline
#undef __launch_bounds__
#define __launch_bounds__(...)
#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)
#define __syncthreads() emu::g_barrier->wait()
#define atomicMin(p, v) emu::atomicMin_emu((p), (v))
using std::max;
using std::min;
#define GAS_KERNEL_EMULATION 1
