// calib — per-node cost of back-to-back kernels inside a CUDA graph on this GPU, for a few launch shapes.
#include <cuda_runtime.h>
#include <stdio.h>
__global__ void k_empty(int *p) {
	if (p && threadIdx.x == 9999) {
		*p = 1;
	}
}
__global__ void k_load2(const int *a, const int *b, int *out) { // two dependent global loads
	extern __shared__ int sm[];
	if (threadIdx.x == 0) {
		int i = a[0];
		int v = b[i & 1];
		if (v == 12345) {
			out[0] = v;
		}
	}
}
static float run(cudaStream_t st, int n, void (*launch)(cudaStream_t)) {
	cudaGraph_t g;
	cudaGraphExec_t ge;
	cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
	for (int i = 0; i < n; i++) {
		launch(st);
	}
	cudaStreamEndCapture(st, &g);
	cudaGraphInstantiate(&ge, g, 0);
	for (int i = 0; i < 3; i++) {
		cudaGraphLaunch(ge, st);
	}
	cudaStreamSynchronize(st);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	cudaEventRecord(e0, st);
	for (int i = 0; i < 10; i++) {
		cudaGraphLaunch(ge, st);
	}
	cudaEventRecord(e1, st);
	cudaStreamSynchronize(st);
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	return 1e3f * ms / (10 * n);
}
static int *d;
static bool pdl = false;
template <int G, int B, int SM>
static void launch_shape(cudaStream_t st) {
	cudaLaunchConfig_t lc{};
	lc.gridDim = dim3(G);
	lc.blockDim = dim3(B);
	lc.dynamicSmemBytes = SM;
	lc.stream = st;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	lc.attrs = attr;
	lc.numAttrs = pdl ? 1 : 0;
	cudaLaunchKernelEx(&lc, k_load2, (const int *)d, (const int *)(d + 4), d + 8);
}
int main() {
	cudaMalloc(&d, 4096);
	cudaMemset(d, 0, 4096);
	cudaStream_t st;
	cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
	cudaFuncSetAttribute(k_load2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	for (int p = 0; p < 2; p++) {
		pdl = p;
		printf("pdl=%d  <<<1,32>>> %.2f us   <<<148,288,200KB>>> %.2f us   <<<256,512>>> %.2f us   <<<592,128>>> %.2f us   <<<512,256>>> %.2f us\n", p,
				run(st, 64, launch_shape<1, 32, 0>), run(st, 64, launch_shape<148, 288, 200 * 1024>), run(st, 64, launch_shape<256, 512, 0>),
				run(st, 64, launch_shape<592, 128, 0>), run(st, 64, launch_shape<512, 256, 0>));
	}
	return 0;
}
