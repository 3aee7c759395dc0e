// fmabench — the K2 consumer inner loop in isolation: 8 consumer warps per SM, one 8-voice x 512-frame stage
// resident in shared memory, no copies.  Reports time per stage for a few shapes of the loop so the FMA side of
// K2 can be compared with the 0.77 us an SM needs to pull a 32 KB stage out of HBM.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ void fma2(float2 &acc, const float2 w, const float2 x) {
	asm("{\n"
		".reg .b64 a, ww, xx;\n"
		"mov.b64 a, {%0, %1};\n"
		"mov.b64 ww, {%2, %3};\n"
		"mov.b64 xx, {%4, %5};\n"
		"fma.rn.f32x2 a, ww, xx, a;\n"
		"mov.b64 {%0, %1}, a;\n"
		"}\n"
		: "+f"(acc.x), "+f"(acc.y)
		: "f"(w.x), "f"(w.y), "f"(x.x), "f"(x.y));
}
__device__ __forceinline__ void fma1(float2 &acc, const float2 w, const float2 x) {
	acc.x = fmaf(w.x, x.x, acc.x);
	acc.y = fmaf(w.y, x.y, acc.y);
}

// MODE 0: FFMA2, loads of U voices first (current K2); 1: same with scalar FFMA; 2: FFMA2, frame-major order
// (all rows of frame 0, then all rows of frame 1); 3: weights read once per voice into registers by LDS.128 pairs and
// x kept as 2 float2, FFMA2 row-major with explicit pairing (wa,x0),(wb,x0),(wa,x1),(wb,x1)
template <int NP, int U, int MODE>
__global__ void __launch_bounds__(256, 1) k_fma(float *out, int stages, int nv) {
	extern __shared__ __align__(16) unsigned char smem[];
	unsigned char *sx = smem;
	unsigned char *sw = smem + 8 * 4096;
	for (int i = threadIdx.x; i < (8 * 4096 + 8 * NP * 8) / 4; i += 256) {
		((float *)smem)[i] = (float)(i & 15) * 0.001f;
	}
	__syncthreads();
	float2 acc[NP][2];
#pragma unroll
	for (int p = 0; p < NP; p++) {
		acc[p][0] = make_float2(0.f, 0.f);
		acc[p][1] = make_float2(0.f, 0.f);
	}
	const int slot = threadIdx.x;
	for (int s = 0; s < stages; s++) {
		for (int v = 0; v + U - 1 < nv; v += U) {
			float4 x[U];
			float4 w[U][(NP + 1) / 2];
#pragma unroll
			for (int u = 0; u < U; u++) {
				const int vv = v + u;
				x[u] = *reinterpret_cast<const float4 *>(sx + (size_t)vv * 4096 + slot * 16);
				const unsigned char *wv = sw + vv * (NP * 8);
#pragma unroll
				for (int p = 0; p < NP; p += 2) {
					w[u][p / 2] = *reinterpret_cast<const float4 *>(wv + p * 8);
				}
			}
#pragma unroll
			for (int u = 0; u < U; u++) {
				const float2 x0 = make_float2(x[u].x, x[u].y), x1 = make_float2(x[u].z, x[u].w);
				if (MODE == 2) {
#pragma unroll
					for (int p = 0; p < NP; p += 2) {
						const float2 wa = make_float2(w[u][p / 2].x, w[u][p / 2].y), wb = make_float2(w[u][p / 2].z, w[u][p / 2].w);
						fma2(acc[p][0], wa, x0);
						fma2(acc[p + 1][0], wb, x0);
					}
#pragma unroll
					for (int p = 0; p < NP; p += 2) {
						const float2 wa = make_float2(w[u][p / 2].x, w[u][p / 2].y), wb = make_float2(w[u][p / 2].z, w[u][p / 2].w);
						fma2(acc[p][1], wa, x1);
						fma2(acc[p + 1][1], wb, x1);
					}
				} else {
#pragma unroll
					for (int p = 0; p < NP; p += 2) {
						const float2 wa = make_float2(w[u][p / 2].x, w[u][p / 2].y), wb = make_float2(w[u][p / 2].z, w[u][p / 2].w);
						if (MODE == 1) {
							fma1(acc[p][0], wa, x0);
							fma1(acc[p][1], wa, x1);
							fma1(acc[p + 1][0], wb, x0);
							fma1(acc[p + 1][1], wb, x1);
						} else {
							fma2(acc[p][0], wa, x0);
							fma2(acc[p][1], wa, x1);
							fma2(acc[p + 1][0], wb, x0);
							fma2(acc[p + 1][1], wb, x1);
						}
					}
				}
			}
		}
		__syncwarp();
	}
	float r = 0.f;
#pragma unroll
	for (int p = 0; p < NP; p++) {
		r += acc[p][0].x + acc[p][0].y + acc[p][1].x + acc[p][1].y;
	}
	if (r == 123.456f) {
		out[threadIdx.x] = r;
	}
}

template <int NP, int U, int MODE>
static void run(const char *name, float *d) {
	const int stages = 2000, nv = 8;
	const size_t smem = 8 * 4096 + 8 * NP * 8;
	cudaFuncSetAttribute(k_fma<NP, U, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	k_fma<NP, U, MODE><<<148, 256, smem>>>(d, 10, nv);
	cudaEventRecord(e0);
	k_fma<NP, U, MODE><<<148, 256, smem>>>(d, stages, nv);
	cudaEventRecord(e1);
	cudaDeviceSynchronize();
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	const double us_stage = 1e3 * ms / stages;
	// FMA lanes per clock per SM at 1.9 GHz: 8 voices * 512 frames * 2 (L,R) * NP per stage
	printf("%-28s NP=%2d U=%d  %.3f us / stage  (%.0f FMA/clk/SM @1.9GHz)  err=%s\n", name, NP, U, us_stage,
			8.0 * 512 * 2 * NP / (us_stage * 1e-6) / 1.9e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
	float *d;
	cudaMalloc(&d, 4096);
	run<8, 4, 0>("ffma2 row-major", d);
	run<8, 4, 1>("ffma scalar", d);
	run<8, 4, 2>("ffma2 frame-major", d);
	run<8, 2, 0>("ffma2 row-major", d);
	run<8, 8, 0>("ffma2 row-major", d);
	run<12, 2, 0>("ffma2 row-major", d);
	run<12, 2, 1>("ffma scalar", d);
	run<12, 4, 0>("ffma2 row-major", d);
	run<20, 2, 0>("ffma2 row-major", d);
	run<20, 2, 1>("ffma scalar", d);
	run<20, 2, 2>("ffma2 frame-major", d);
	run<20, 1, 0>("ffma2 row-major", d);
	run<20, 4, 0>("ffma2 row-major", d);
	run<24, 2, 0>("ffma2 row-major", d);
	return 0;
}
