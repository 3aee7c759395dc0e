// gas_gain.cu — K1: batched AudioSpatializerInstance3D::calculate_spatialization
// (reference audio_spatializer_3d.cpp:277-489) + update_spatializer_parameters / get_bus_map
// (reference audio_spatializer.cpp:258-324).  A few lanes per emitter (4 by default): the scalar chain
// (transform, attenuation, filter gain, doppler) is evaluated redundantly by the lanes of an emitter, the SPCAP
// speaker gains (one double pow each, the long pole) are dealt across them, and every lane stores its own
// (channel pair, side) elements of the volume tables.  The kernel is a long dependent chain per emitter
// (latency-bound, ~12 us whatever the batch size up to a full GPU), so what matters is enough registers to
// schedule the independent sub-chains side by side (128, no spills) and CTAs small enough to slip in beside
// the mix kernels of the previous block.
//
// Compiled with -fmad=false: the reference mixes float storage with double intermediates (SURVEY Q3,
// Q8, Q10) and the gains must come out the way a scalar x86-64 build produces them, so no contraction.
// float transcendentals are evaluated in double and narrowed, which is correctly rounded in practice
// and therefore agrees with a correctly-rounded libm.
#include "gas_gain.cuh"

namespace {

using namespace gasgain;

// NL lanes per emitter (8, 4, 2 or 1), see gain_emitter.
template <int NL, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_gain(DevTables t, GlobalCfg g, int n, const gas_emitter *__restrict__ emitters,
		int n_listeners, const gas_listener *__restrict__ listeners, const ListenerPre *__restrict__ pre, const gas_area *__restrict__ areas, int n_areas,
		gas_params *__restrict__ out) {
	const int i = (blockIdx.x * blockDim.x + threadIdx.x) / NL;
	const int l = threadIdx.x & (NL - 1);
	const int gbase = threadIdx.x & (32 - NL); // first lane of this emitter's group inside the warp
	const unsigned gm = NL == 32 ? 0xffffffffu : (((1u << NL) - 1u) << gbase);
	if (i >= n) {
		return;
	}
	gain_emitter<NL>(t, g, i, l, gbase, gm, emitters, n_listeners, listeners, pre, areas, n_areas, out);
}

__global__ void k_listener_pre(int n, const gas_listener *__restrict__ listeners, ListenerPre *__restrict__ pre) {
	const int i = threadIdx.x;
	if (i < n) {
		listener_precompute(listeners[i], pre[i]);
	}
}

__global__ void __launch_bounds__(128) k_params_set(DevTables t, GlobalCfg g, int n, const int32_t *__restrict__ ids, const gas_params *__restrict__ params) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) {
		return;
	}
	gas_params p = params[i];
	for (int k = 0; k < GAS_MAX_BUSES_PER_PLAYBACK; k++) {
		p.bus[k] = resolve_bus(g, p.bus[k]);
	}
	commit_params(t, ids[i], p);
}

// proxies (re)registered: current details from the current parameters, previous details empty
// (reference audio_spatializer.cpp:75-95, upstream AudioServer::start_playback_stream)
__global__ void __launch_bounds__(128) k_instance_start(DevTables t, int n, const int32_t *__restrict__ ids) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) {
		return;
	}
	int q = ids[i];
	t.inst_active[q] = 1;
	BusDetails z;
	z.n = 0;
	for (int k = 0; k < GAS_MAX_BUSES_PER_PLAYBACK; k++) {
		z.bus[k] = 0;
		for (int c = 0; c < 4; c++) {
			z.vol[k][c][0] = z.vol[k][c][1] = 0.f;
		}
	}
	t.inst_prev[q] = z; // both parity buffers
	t.inst_prev[t.max_instances + q] = z;
	push_bus_map(t.inst_params[q], inst_mix_channels(t, q), t.inst_cur[q]);
}

} // namespace

cudaError_t launch_gain(gas_ctx *ctx, int n, const gas_emitter *d_em, int n_listeners, const gas_listener *d_l,
		const gas_area *d_areas, gas_params *d_out, cudaStream_t st) {
	if (n <= 0) {
		return cudaSuccess;
	}
	// Launch shape, measured inside the mix step on B200 (16384 emitters, the kernel running beside the mix kernels):
	//   4 lanes x 64-thread CTAs, 128 registers (no spills): 44.7 us / step, 12.3 us alone    <- default
	//   8 lanes x 64-thread CTAs, 128 registers:             45.0 us / step, 14.4 us alone
	//   8 lanes x 256-thread CTAs, 64 registers (spills):    50.7 us / step, 15.1 us alone
	//   2 lanes x 64-thread CTAs:                            49.9 us / step
	// Small CTAs matter more than anything else: the streaming mix kernel leaves ~17 K registers per SM free, which
	// two 64 x 128-register CTAs fill, while one 256-thread CTA has to squeeze into 64 registers per thread to get in.
	static int shape = -1;
	if (shape < 0) {
		const char *e = getenv("GAS_K1_SHAPE"); // experiments: 1 = 8 lanes x 256 threads, 2 = 8 lanes x 64 threads
		shape = e ? atoi(e) : 0;
	}
#define GAS_K1_LAUNCH(NL_, T_, M_)                                                                                          \
	k_gain<NL_, T_, M_><<<(int)(((long long)n * NL_ + T_ - 1) / T_), T_, 0, st>>>(ctx->t, ctx->g, n, d_em, n_listeners, d_l, ctx->d_listener_pre, d_areas, ctx->n_areas_res, d_out)
	switch (shape) {
		case 1: GAS_K1_LAUNCH(8, 256, 4); break;
		case 2: GAS_K1_LAUNCH(8, 64, 8); break;
		case 3: GAS_K1_LAUNCH(1, 32, 16); break; // one lane per emitter: 8x fewer issue slots, the whole grid fits beside the streaming kernel
		case 4: GAS_K1_LAUNCH(2, 32, 16); break;
		case 5: GAS_K1_LAUNCH(2, 64, 8); break;
		case 6: GAS_K1_LAUNCH(4, 32, 16); break;
		default: GAS_K1_LAUNCH(4, 64, 8); break;
	}
#undef GAS_K1_LAUNCH
	ctx->launches++;
	return cudaGetLastError();
}

cudaError_t launch_listener_pre(gas_ctx *ctx, int n_listeners, cudaStream_t st) {
	if (n_listeners <= 0) {
		return cudaSuccess;
	}
	k_listener_pre<<<1, GAS_MAX_LISTENERS, 0, st>>>(n_listeners, ctx->d_listeners, ctx->d_listener_pre);
	ctx->launches++;
	return cudaGetLastError();
}

cudaError_t launch_params_set(gas_ctx *ctx, int n, const int32_t *d_ids, const gas_params *d_params, cudaStream_t st) {
	if (n <= 0) {
		return cudaSuccess;
	}
	k_params_set<<<(n + 127) / 128, 128, 0, st>>>(ctx->t, ctx->g, n, d_ids, d_params);
	ctx->launches++;
	return cudaGetLastError();
}

cudaError_t launch_instance_start(gas_ctx *ctx, int n, const int32_t *d_ids, cudaStream_t st) {
	if (n <= 0) {
		return cudaSuccess;
	}
	k_instance_start<<<(n + 127) / 128, 128, 0, st>>>(ctx->t, n, d_ids);
	ctx->launches++;
	return cudaGetLastError();
}
