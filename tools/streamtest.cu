// streamtest — raw read-streaming patterns on this GPU: how fast can 148 persistent CTAs pull V rows of
// 4 KB (gathered through an index list) out of HBM with (a) per-lane LDG.128 into registers, (b) per-warp
// cp.async (LDGSTS) rings, (c) 1-D bulk TMA copies issued by 1..4 producer warps?  Rotates over 8 source
// sets (> L2).  Prints GB/s per variant for linear and shuffled row order.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

static const int V = 16384, ROWB = 4096, SETS = 8;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// (a) registers: warp w of CTA b reads piece (w % wpr) of its rows; U loads in flight per lane
template <int U>
__global__ void __launch_bounds__(256, 1) k_ldg(const float4 *__restrict__ src, const int *__restrict__ order, int v_per_cta, int wpr, float *out) {
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int groups = 8 / wpr, g = warp / wpr, h = warp % wpr;
	const int v0 = blockIdx.x * v_per_cta;
	float4 acc = make_float4(0, 0, 0, 0);
	// each warp covers pieces h, h + wpr, ... of a row (8 pieces of 512 B) for voices g, g + groups, ...
	for (int i = g; i < v_per_cta; i += groups * U) {
		float4 x[U][8];
#pragma unroll
		for (int u = 0; u < U; u++) {
			const int vi = i + u * groups;
			const int row = vi < v_per_cta ? order[v0 + vi] : -1;
#pragma unroll
			for (int p = 0; p < 8; p++) {
				if (p < 8 / wpr) {
					x[u][p] = row >= 0 ? __ldcs(src + (size_t)row * 256 + (h + p * wpr) * 32 + lane) : make_float4(0, 0, 0, 0);
				}
			}
		}
#pragma unroll
		for (int u = 0; u < U; u++) {
#pragma unroll
			for (int p = 0; p < 8; p++) {
				if (p < 8 / wpr) {
					acc.x += x[u][p].x;
					acc.y += x[u][p].y;
					acc.z += x[u][p].z;
					acc.w += x[u][p].w;
				}
			}
		}
	}
	if (acc.x + acc.y + acc.z + acc.w == 1234.5f) {
		out[0] = acc.x;
	}
}

// (b) per-warp cp.async ring, depth D, each warp one 512 B piece per voice (tile = 64*wpr frames)
template <int D>
__global__ void __launch_bounds__(256, 1) k_ring(const float4 *__restrict__ src, const int *__restrict__ order, int v_per_cta, int wpr, float *out) {
	extern __shared__ __align__(128) unsigned char smem[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int groups = 8 / wpr, g = warp / wpr, h = warp % wpr;
	const int v0 = blockIdx.x * v_per_cta;
	unsigned char *ring = smem + warp * D * 512;
	float4 acc = make_float4(0, 0, 0, 0);
	const int n_tiles = 8 / wpr;
	for (int t = 0; t < n_tiles; t++) {
		const int n_mine = (v_per_cta - g + groups - 1) / groups;
		int si = 0, sc = 0;
		auto issue = [&](int i) {
			if (i < n_mine) {
				const int row = order[v0 + g + groups * i];
				asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(ring + si * 512 + lane * 16)),
						"l"(src + (size_t)row * 256 + (t * wpr + h) * 32 + lane)
						: "memory");
			}
			asm volatile("cp.async.commit_group;" ::: "memory");
			if (++si == D) si = 0;
		};
		for (int i = 0; i < D - 1; i++) issue(i);
		for (int i = 0; i < n_mine; i++) {
			issue(i + D - 1);
			asm volatile("cp.async.wait_group %0;" ::"n"(D - 1) : "memory");
			const float4 x = *reinterpret_cast<const float4 *>(ring + sc * 512 + lane * 16);
			acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
			if (++sc == D) sc = 0;
		}
		asm volatile("cp.async.wait_group 0;" ::: "memory");
	}
	if (acc.x + acc.y + acc.z + acc.w == 1234.5f) out[0] = acc.x;
}

// (c) bulk TMA: NPW producer warps, each owns every NPW-th stage of a ring of S stages of VB rows x 4 KB
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
	uint32_t ok = 0;
	do {
		asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
				: "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
	} while (!ok);
}
template <int NPW, int VB, int S>
__global__ void __launch_bounds__(256 + 32 * NPW, 1) k_tma(const float4 *__restrict__ src, const int *__restrict__ order, int v_per_cta, float *out) {
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ __align__(8) uint64_t full[S], empty[S];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (threadIdx.x == 0) {
		for (int s = 0; s < S; s++) {
			asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(1));
			asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[s])), "r"(8));
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	const int v0 = blockIdx.x * v_per_cta;
	const int n_units = (v_per_cta + VB - 1) / VB;
	if (warp >= 8) {
		const int pw = warp - 8;
		for (int u = pw; u < n_units; u += NPW) {
			const int stage = u % S;
			const uint32_t phase = (u / S) & 1;
			const int nv = min(VB, v_per_cta - u * VB);
			mbar_wait(&empty[stage], phase ^ 1u);
			if (lane == 0) {
				asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[stage])), "r"(nv * 4096) : "memory");
			}
			__syncwarp();
			if (lane < nv) {
				const int row = order[v0 + u * VB + lane];
				asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
									 smem_u32(smem + (size_t)stage * VB * 4096 + lane * 4096)),
						"l"(src + (size_t)row * 256), "r"(4096), "r"(smem_u32(&full[stage]))
						: "memory");
			}
		}
	} else {
		float4 acc = make_float4(0, 0, 0, 0);
		for (int u = 0; u < n_units; u++) {
			const int stage = u % S;
			const uint32_t phase = (u / S) & 1;
			const int nv = min(VB, v_per_cta - u * VB);
			mbar_wait(&full[stage], phase);
			for (int v = 0; v < nv; v++) {
				const float4 x = *reinterpret_cast<const float4 *>(smem + (size_t)stage * VB * 4096 + v * 4096 + threadIdx.x * 16);
				acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
			}
			__syncwarp();
			if (lane == 0) {
				asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[stage])) : "memory");
			}
		}
		if (acc.x + acc.y + acc.z + acc.w == 1234.5f) out[0] = acc.x;
	}
}

template <typename L>
static float timeit(cudaStream_t st, L launch) {
	for (int i = 0; i < 3; i++) launch(i);
	cudaStreamSynchronize(st);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	const int reps = 40;
	cudaEventRecord(e0, st);
	for (int i = 0; i < reps; i++) launch(i);
	cudaEventRecord(e1, st);
	cudaStreamSynchronize(st);
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
	return 1e3f * ms / reps; // us per launch (back to back launches, includes ~1-2 us of launch gap)
}

int main() {
	std::vector<float4 *> src(SETS);
	for (auto &p : src) {
		cudaMalloc(&p, (size_t)V * ROWB);
		cudaMemset(p, 0, (size_t)V * ROWB);
	}
	std::vector<int> lin(V), shuf(V);
	for (int i = 0; i < V; i++) lin[i] = shuf[i] = i;
	srand(1);
	std::random_shuffle(shuf.begin(), shuf.end());
	int *d_lin, *d_shuf;
	float *d_out;
	cudaMalloc(&d_lin, V * 4);
	cudaMalloc(&d_shuf, V * 4);
	cudaMalloc(&d_out, 64);
	cudaMemcpy(d_lin, lin.data(), V * 4, cudaMemcpyHostToDevice);
	cudaMemcpy(d_shuf, shuf.data(), V * 4, cudaMemcpyHostToDevice);
	cudaStream_t st;
	cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
	const int nsm = 148, vpc = (V + nsm - 1) / nsm; // 111 (the last CTAs read a few rows twice: clamp)
	std::vector<int> pad(nsm * vpc);
	const double bytes = (double)nsm * vpc * ROWB;
	auto report = [&](const char *name, float us_lin, float us_shuf) {
		printf("%-34s linear %7.2f us %6.0f GB/s | shuffled %7.2f us %6.0f GB/s\n", name, us_lin, bytes / us_lin / 1e3, us_shuf, bytes / us_shuf / 1e3);
	};
	// order arrays padded to nsm*vpc
	int *d_ord[2];
	for (int k = 0; k < 2; k++) {
		for (int i = 0; i < nsm * vpc; i++) pad[i] = (k ? shuf : lin)[i % V];
		cudaMalloc(&d_ord[k], pad.size() * 4);
		cudaMemcpy(d_ord[k], pad.data(), pad.size() * 4, cudaMemcpyHostToDevice);
	}
#define BOTH(name, expr)                                                    \
	{                                                                       \
		float r[2];                                                         \
		for (int k = 0; k < 2; k++) {                                       \
			const int *ord = d_ord[k];                                      \
			r[k] = timeit(st, [&](int i) { const float4 *s = src[i % SETS]; expr; }); \
		}                                                                   \
		report(name, r[0], r[1]);                                           \
	}
	BOTH("ldg regs U=2 wpr=8 (4KB rows)", (k_ldg<2><<<nsm, 256, 0, st>>>(s, ord, vpc, 8, d_out)));
	BOTH("ldg regs U=4 wpr=8", (k_ldg<4><<<nsm, 256, 0, st>>>(s, ord, vpc, 8, d_out)));
	BOTH("ldg regs U=4 wpr=2 (1KB pieces)", (k_ldg<4><<<nsm, 256, 0, st>>>(s, ord, vpc, 2, d_out)));
	BOTH("ldg regs U=8 wpr=1 (512B pieces)", (k_ldg<8><<<nsm, 256, 0, st>>>(s, ord, vpc, 1, d_out)));
	cudaFuncSetAttribute(k_ring<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 24 * 512);
	cudaFuncSetAttribute(k_ring<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 48 * 512);
	BOTH("cp.async ring D=24 wpr=8", (k_ring<24><<<nsm, 256, 8 * 24 * 512, st>>>(s, ord, vpc, 8, d_out)));
	BOTH("cp.async ring D=24 wpr=2", (k_ring<24><<<nsm, 256, 8 * 24 * 512, st>>>(s, ord, vpc, 2, d_out)));
	BOTH("cp.async ring D=48 wpr=8", (k_ring<48><<<nsm, 256, 8 * 48 * 512, st>>>(s, ord, vpc, 8, d_out)));
	BOTH("cp.async ring D=48 wpr=1", (k_ring<48><<<nsm, 256, 8 * 48 * 512, st>>>(s, ord, vpc, 1, d_out)));
#define TMA(NPW, VB, S)                                                                                                \
	cudaFuncSetAttribute(k_tma<NPW, VB, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, VB * S * 4096);               \
	BOTH("tma bulk 4KB npw=" #NPW " vb=" #VB " S=" #S, (k_tma<NPW, VB, S><<<nsm, 256 + 32 * NPW, VB * S * 4096, st>>>(s, ord, vpc, d_out)));
	TMA(1, 8, 6)
	TMA(2, 8, 6)
	TMA(4, 8, 6)
	TMA(1, 4, 12)
	TMA(4, 4, 12)
	TMA(1, 16, 3)
	return 0;
}
