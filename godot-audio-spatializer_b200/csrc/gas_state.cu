// gas_state.cu — small bookkeeping kernels: slot (re)initialisation, state export/import.
#include "gas_internal.h"

namespace {

__device__ void params_defaults(gas_params &p) { // spatializer_parameters.h:48, audio_spatializer_3d.h:67-68
	for (int c = 0; c < 4; c++) {
		p.mix_volumes[c][0] = p.mix_volumes[c][1] = 0.f;
	}
	p.pitch_scale = 1.0f;
	p.linear_attenuation = 0.0f;
	p.attenuation_filter_cutoff_hz = 5000.0f;
	p.update_parameters = 0;
	p.n_bus = 0;
	for (int k = 0; k < GAS_MAX_BUSES_PER_PLAYBACK; k++) {
		p.bus[k] = 0;
		for (int c = 0; c < 4; c++) {
			p.bus_volumes[k][c][0] = p.bus_volumes[k][c][1] = 0.f;
		}
	}
}

__device__ void details_clear(BusDetails &d) {
	d.n = 0;
	for (int k = 0; k < GAS_MAX_BUSES_PER_PLAYBACK; k++) {
		d.bus[k] = 0;
		for (int c = 0; c < 4; c++) {
			d.vol[k][c][0] = d.vol[k][c][1] = 0.f;
		}
	}
}

__device__ void instance_reset(DevTables &t, int q, int spat) {
	t.inst_spat[q] = spat;
	params_defaults(t.inst_params[q]);
	t.inst_was_further[q] = 0;
	t.inst_active[q] = 0;
	details_clear(t.inst_cur[q]);
	details_clear(t.inst_prev[q]);
	details_clear(t.inst_prev[t.max_instances + q]);
	{ // kind / mix_channel_mode / effect binding are latched at instantiate() (reference audio_spatializer_3d.cpp:645-652)
		const gas_spatializer &s = t.spat[spat];
		const int mode = s.kind == GAS_SPATIALIZER_EFFECT ? MODE_E : (s.mix_channel_mode ? MODE_B : MODE_A);
		t.inst_mode[q] = mode | ((s.effect_gain_binding + 1) << 8);
	}
	t.inst_fx[q] = t.spat[spat].chain;
	t.inst_threshold[q] = t.threshold_default; // playback_disable_threshold_db = -80 (reference audio_spatializer.h:87)
}

// AudioSpatializer3D defaults, audio_spatializer_3d.h:171-188
__device__ void spat_defaults(gas_spatializer &s) {
	s.kind = GAS_SPATIALIZER_3D;
	s.attenuation_model = GAS_ATTENUATION_INVERSE_DISTANCE;
	s.unit_size = 10.0f;
	s.max_distance = 0.0f;
	s.panning_strength = 1.0f;
	s.area_mask = 1;
	s.emission_angle_enabled = 0;
	s.emission_angle = 45.0f;
	s.emission_angle_filter_attenuation_db = -12.0f;
	s.attenuation_filter_cutoff_hz = 5000.0f;
	s.attenuation_filter_db = -24.0f;
	s.doppler_tracking = GAS_DOPPLER_TRACKING_DISABLED;
	s.doppler_speed_of_sound = 343.0f;
	s.mix_channel_mode = 0;
	s.effect_gain_binding = -1;
	s.chain.n_effects = 0;
	for (int e = 0; e < GAS_MAX_EFFECTS; e++) {
		s.chain.effects[e].mode = GAS_FILTER_HIGHSHELF;
		s.chain.effects[e].cutoff_hz = 2000.f;
		s.chain.effects[e].resonance = 0.5f;
		s.chain.effects[e].gain = 1.f;
		s.chain.effects[e].stages = 1;
	}
}

__global__ void k_defaults(DevTables t, GlobalCfg g) {
	const int tid = blockIdx.x * blockDim.x + threadIdx.x;
	const int n = gridDim.x * blockDim.x;
	for (int i = tid; i < g.max_spatializers; i += n) {
		spat_defaults(t.spat[i]);
	}
	__threadfence();
	for (int q = tid; q < g.max_instances; q += n) {
		t.inst_spat[q] = 0;
		params_defaults(t.inst_params[q]);
		t.inst_was_further[q] = 0;
		t.inst_active[q] = 0;
		details_clear(t.inst_cur[q]);
		details_clear(t.inst_prev[q]);
		details_clear(t.inst_prev[t.max_instances + q]);
		t.inst_mode[q] = MODE_A;
		t.inst_fx[q].n_effects = 0;
		t.inst_threshold[q] = t.threshold_default;
	}
}

__global__ void k_instance_init(DevTables t, int n, const int32_t *__restrict__ ids, const int32_t *__restrict__ spat) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) {
		instance_reset(t, ids[i], spat[i]);
	}
}

__global__ void k_instance_stop(DevTables t, int n, const int32_t *__restrict__ ids) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) {
		t.inst_active[ids[i]] = 0;
	}
}

constexpr int kFxFloats = GAS_MAX_EFFECTS * 2 * GAS_MAX_FILTER_STAGES * 4;

__global__ void k_voice_init(DevTables t, int n, const int32_t *__restrict__ ids) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) {
		return;
	}
	const int v = ids[i];
	for (int k = 0; k < 8; k++) {
		t.vs_prev[(size_t)v * 8 + k] = 0.f;
	}
	gas_processor_state z{};
	for (int k = 0; k < 8; k++) {
		t.vs_proc[(size_t)v * 8 + k] = z;
	}
	for (int k = 0; k < kFxFloats; k++) {
		t.vs_fx[(size_t)v * kFxFloats + k] = 0.f;
	}
	// start_playback_stream (reference audio_spatializer.cpp:57-72): lookahead zeroed, active and has_frames set
	for (int k = 0; k < GAS_LOOKAHEAD_BUFFER_SIZE; k++) {
		t.vs_look[(size_t)v * GAS_LOOKAHEAD_BUFFER_SIZE + k] = gas_frame{ 0.f, 0.f };
	}
	t.vs_life[v] = GAS_VOICE_ACTIVE | GAS_VOICE_HAS_FRAMES;
}

__global__ void k_state_export(DevTables t, int n, const int32_t *__restrict__ ids, gas_voice_state *__restrict__ out) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) {
		return;
	}
	const int v = ids[i];
	gas_voice_state *o = &out[i];
	for (int c = 0; c < 4; c++) {
		o->prev_mix_volumes[c][0] = t.vs_prev[(size_t)v * 8 + c * 2];
		o->prev_mix_volumes[c][1] = t.vs_prev[(size_t)v * 8 + c * 2 + 1];
	}
	for (int k = 0; k < 8; k++) {
		o->filter_processors[k] = t.vs_proc[(size_t)v * 8 + k];
	}
	float *fx = &o->effect_history[0][0][0][0];
	for (int k = 0; k < kFxFloats; k++) {
		fx[k] = t.vs_fx[(size_t)v * kFxFloats + k];
	}
}

__global__ void k_state_import(DevTables t, int n, const int32_t *__restrict__ ids, const gas_voice_state *__restrict__ in) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) {
		return;
	}
	const int v = ids[i];
	const gas_voice_state *o = &in[i];
	for (int c = 0; c < 4; c++) {
		t.vs_prev[(size_t)v * 8 + c * 2] = o->prev_mix_volumes[c][0];
		t.vs_prev[(size_t)v * 8 + c * 2 + 1] = o->prev_mix_volumes[c][1];
	}
	for (int k = 0; k < 8; k++) {
		t.vs_proc[(size_t)v * 8 + k] = o->filter_processors[k];
	}
	const float *fx = &o->effect_history[0][0][0][0];
	for (int k = 0; k < kFxFloats; k++) {
		t.vs_fx[(size_t)v * kFxFloats + k] = fx[k];
	}
}

__global__ void k_params_get(DevTables t, int n, const int32_t *__restrict__ ids, gas_params *__restrict__ out) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) {
		out[i] = t.inst_params[ids[i]];
	}
}

__global__ void k_fx_set(DevTables t, int n, const int32_t *__restrict__ ids, const gas_effect_chain *__restrict__ in) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) {
		t.inst_fx[ids[i]] = in[i];
	}
}

inline int blocks_for(int n) { return (n + 127) / 128; }

} // namespace

#define LAUNCH1(kernel, n, ...)                                  \
	if ((n) <= 0) {                                              \
		return cudaSuccess;                                      \
	}                                                            \
	kernel<<<blocks_for(n), 128, 0, st>>>(ctx->t, n, __VA_ARGS__); \
	ctx->launches++;                                             \
	return cudaGetLastError();

// class table of the block plan: every slot free, aux words unset, the generic classes in the last slots
__global__ void k_plan_defaults(BlockPlan plan) {
	const int i = threadIdx.x;
	if (i >= GAS_MAX_CLASSES) {
		return;
	}
	plan.cls_aux[i] = CLS_AUX_NONE;
	plan.cls_idle[i] = 0;
	unsigned long long key = 0ULL;
	if (i >= GAS_CLS_DYNAMIC) {
		const int k = i - GAS_CLS_DYNAMIC; // mode * 2 + filter
		key = cls_key(PATH_VOICE, k >> 1, CLS_GENERIC | ((k & 1) ? CLS_FILT : 0u), 0, 0u, 0u);
	}
	plan.cls_key[i] = key;
}

cudaError_t launch_defaults(gas_ctx *ctx, cudaStream_t st) {
	k_defaults<<<64, 128, 0, st>>>(ctx->t, ctx->g);
	k_plan_defaults<<<1, GAS_MAX_CLASSES, 0, st>>>(ctx->plan);
	ctx->launches += 2;
	return cudaGetLastError();
}
cudaError_t launch_instance_init(gas_ctx *ctx, int n, const int32_t *d_ids, const int32_t *d_spat, cudaStream_t st) { LAUNCH1(k_instance_init, n, d_ids, d_spat) }
cudaError_t launch_instance_stop(gas_ctx *ctx, int n, const int32_t *d_ids, cudaStream_t st) { LAUNCH1(k_instance_stop, n, d_ids) }
cudaError_t launch_voice_init(gas_ctx *ctx, int n, const int32_t *d_ids, cudaStream_t st) { LAUNCH1(k_voice_init, n, d_ids) }
cudaError_t launch_state_export(gas_ctx *ctx, int n, const int32_t *d_ids, gas_voice_state *d_out, cudaStream_t st) { LAUNCH1(k_state_export, n, d_ids, d_out) }
cudaError_t launch_state_import(gas_ctx *ctx, int n, const int32_t *d_ids, const gas_voice_state *d_in, cudaStream_t st) { LAUNCH1(k_state_import, n, d_ids, d_in) }
cudaError_t launch_params_get(gas_ctx *ctx, int n, const int32_t *d_ids, gas_params *d_out, cudaStream_t st) { LAUNCH1(k_params_get, n, d_ids, d_out) }
cudaError_t launch_fx_set(gas_ctx *ctx, int n, const int32_t *d_ids, const gas_effect_chain *d_in, cudaStream_t st) { LAUNCH1(k_fx_set, n, d_ids, d_in) }
