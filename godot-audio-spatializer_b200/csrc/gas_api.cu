// gas_api.cu — the C ABI (include/gas.h): context, device memory, stream ordering, validation.
// Host code only; every data-path operation is a kernel launched from the other .cu files.
// There is no CPU implementation of anything here: without a CUDA device gas_create fails.
#include "gas_internal.h"

#include <stdarg.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

static thread_local std::string g_create_error;

int gas_fail(gas_ctx *ctx, int status, const char *fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	if (ctx) {
		ctx->err = buf;
	} else {
		g_create_error = buf;
	}
	return status;
}

namespace {

template <typename T>
cudaError_t dev_alloc(T **p, size_t n) {
	cudaError_t e = cudaMalloc((void **)p, n * sizeof(T));
	if (e == cudaSuccess) {
		e = cudaMemset(*p, 0, n * sizeof(T));
	}
	return e;
}

// SpeakerPlacementConfiguration::update_speaker_configuration (reference audio_spatializer_3d.cpp:903-916) with
// the speaker set of :47-55.  Every operation is a separately rounded IEEE float/double operation (volatile
// keeps the host compiler from contracting), so the values are bit-identical to a device evaluation.
void spcap_constants(int speaker_mode, float dir[7][3], float eff[7]) {
	static const float raw[7][3] = { { -1, 0, -1 }, { 1, 0, -1 }, { 0, 0, -1 }, { -1, 0, 1 }, { 1, 0, 1 }, { -1, 0, 0 }, { 1, 0, 0 } };
	const int count = speaker_mode == GAS_SPEAKER_SURROUND_31 ? 3 : (speaker_mode == GAS_SPEAKER_SURROUND_51 ? 5 : (speaker_mode == GAS_SPEAKER_SURROUND_71 ? 7 : 2));
	for (int i = 0; i < 7; i++) { // Vector3::normalized
		volatile float xx = raw[i][0] * raw[i][0], yy = raw[i][1] * raw[i][1], zz = raw[i][2] * raw[i][2];
		volatile float l2 = xx + yy;
		l2 = l2 + zz;
		volatile float len = sqrtf(l2);
		for (int k = 0; k < 3; k++) {
			dir[i][k] = l2 == 0.0f ? 0.0f : raw[i][k] / len;
		}
		eff[i] = 0.f;
	}
	for (int i = 0; i < count; i++) {
		for (int j = 0; j < count; j++) {
			volatile float a = dir[i][0] * dir[j][0], b = dir[i][1] * dir[j][1], c = dir[i][2] * dir[j][2];
			volatile float d = a + b;
			d = d + c;
			volatile double term = 0.5 * (1.0 + (double)d);
			eff[i] = (float)((double)eff[i] + term);
		}
	}
}

void refresh_globals(gas_ctx *ctx) {
	ctx->cfg_epoch++; // graphs captured under the old globals are refused from now on (gas_graph_launch)
	ctx->g.speaker_mode = ctx->cfg.speaker_mode;
	ctx->g.channels = ctx->cfg.speaker_mode + 1;
	ctx->g.num_buses = ctx->cfg.num_buses;
	ctx->g.mix_rate = ctx->cfg.mix_rate;
	ctx->g.global_panning = ctx->cfg.global_panning_strength;
	ctx->g.max_instances = ctx->cfg.max_instances;
	ctx->g.max_voices = ctx->cfg.max_voices;
	ctx->g.max_spatializers = ctx->cfg.max_spatializers;
	spcap_constants(ctx->g.speaker_mode, ctx->g.spk_dir, ctx->g.spk_eff);
}

// Stream ordering between the gain side and the mix side (the reference's physics / audio threads):
// the tiny per-block prologue is the only mix-side kernel that reads instance tables, so gain-side work
// waits for the last prologue and the next prologue waits for the last gain-side work.
int gain_side_begin(gas_ctx *ctx) {
	if (ctx->prologue_pending) {
		GAS_CUDA(ctx, cudaStreamWaitEvent(ctx->s_gain, ctx->ev_prologue_done, 0));
	}
	if (ctx->stream_started_pending) {
		// ... and for the streaming kernel of that block to have all its CTAs resident: the gain kernel's many small CTAs
		// would otherwise take the SMs first and hold the streaming kernel's one-CTA-per-SM grid up by most of their own
		// duration (measured: its CTAs started 8 us late on average); started afterwards they run beside it
		GAS_CUDA(ctx, cudaStreamWaitEvent(ctx->s_gain, ctx->ev_stream_started, 0));
		ctx->stream_started_pending = false;
	}
	return GAS_OK;
}
int gain_side_end(gas_ctx *ctx) {
	GAS_CUDA(ctx, cudaEventRecord(ctx->ev_gain_done, ctx->s_gain));
	ctx->gain_pending = true;
	return GAS_OK;
}

bool ids_valid(const int32_t *ids, int n, int limit) {
	for (int i = 0; i < n; i++) {
		if (ids[i] < 0 || ids[i] >= limit) {
			return false;
		}
	}
	return true;
}

bool spat_valid(const gas_spatializer *s) { // reference audio_spatializer_3d.cpp:670-672,695-697,728-730,737-739,758-760
	if (s->kind != GAS_SPATIALIZER_3D && s->kind != GAS_SPATIALIZER_EFFECT) {
		return false;
	}
	if (!(s->max_distance >= 0.0f)) {
		return false;
	}
	if (!(s->emission_angle >= 0.f && s->emission_angle <= 90.f)) {
		return false;
	}
	if (s->attenuation_model < 0 || s->attenuation_model >= 4) {
		return false;
	}
	if (!(s->panning_strength >= 0.f)) {
		return false;
	}
	if (!(s->doppler_speed_of_sound > 0.f)) {
		return false;
	}
	if (s->chain.n_effects < 0 || s->chain.n_effects > GAS_MAX_EFFECTS) {
		return false;
	}
	return true;
}

// per-kernel timing: a pair of timing events around a launch on the mix stream
int prof_drain(gas_ctx *ctx) {
	if (ctx->prof_used == 0) {
		return GAS_OK;
	}
	GAS_CUDA(ctx, cudaStreamSynchronize(ctx->s_mix));
	GAS_CUDA(ctx, cudaStreamSynchronize(ctx->s_gain));
	for (size_t i = 0; i < ctx->prof_used; i++) {
		float ms = 0.f;
		GAS_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->prof_pairs[i].a, ctx->prof_pairs[i].b));
		ctx->prof_ms[ctx->prof_pairs[i].kind] += ms;
		ctx->prof_n[ctx->prof_pairs[i].kind] += 1;
	}
	ctx->prof_used = 0;
	return GAS_OK;
}

// Inside a capture the timing events become event-record nodes of the graph (cudaEventRecordExternal), so a
// profiled graph reports each kernel's duration as it runs inside the replayed step.

gas_ctx::ProfPair *prof_open(gas_ctx *ctx, int kind, cudaStream_t st = nullptr) {
	st = st ? st : ctx->s_mix;
	if (!ctx->profiling) {
		return nullptr;
	}
	if (ctx->capturing) {
		if (!ctx->gev[kind][0]) {
			if (cudaEventCreate(&ctx->gev[kind][0]) != cudaSuccess || cudaEventCreate(&ctx->gev[kind][1]) != cudaSuccess) {
				return nullptr;
			}
		}
		ctx->gev_used[kind] = true;
		ctx->capture_profiled = true;
		ctx->graph_pair[kind].a = ctx->gev[kind][0];
		ctx->graph_pair[kind].b = ctx->gev[kind][1];
		ctx->graph_pair[kind].kind = kind;
		ctx->graph_pair[kind].st = st;
		cudaEventRecordWithFlags(ctx->graph_pair[kind].a, st, cudaEventRecordExternal);
		return &ctx->graph_pair[kind];
	}
	if (ctx->prof_used >= 3072 && prof_drain(ctx) != GAS_OK) {
		return nullptr;
	}
	if (ctx->prof_used == ctx->prof_pairs.size()) {
		gas_ctx::ProfPair pp{};
		if (cudaEventCreate(&pp.a) != cudaSuccess || cudaEventCreate(&pp.b) != cudaSuccess) {
			return nullptr;
		}
		ctx->prof_pairs.push_back(pp);
	}
	gas_ctx::ProfPair *p = &ctx->prof_pairs[ctx->prof_used++];
	p->kind = kind;
	p->st = st;
	cudaEventRecord(p->a, st);
	return p;
}

void prof_close(gas_ctx *ctx, gas_ctx::ProfPair *p) {
	if (p) {
		cudaEventRecordWithFlags(p->b, p->st, ctx->capturing ? cudaEventRecordExternal : cudaEventRecordDefault);
	}
}

// Outstanding voice-parallel kernels of pipelined steps (side stream): the mix stream waits for them.
int join_voice_stream(gas_ctx *ctx) {
	for (int i = 0; i < GAS_PLAN_DEPTH; i++) {
		if (ctx->block_inflight[i]) {
			GAS_CUDA(ctx, cudaStreamWaitEvent(ctx->s_mix, ctx->ev_block_done[i], 0));
			ctx->block_inflight[i] = false;
		}
	}
	return GAS_OK;
}

// One block, one call: plan (k_plan), stream (step kernel without control work), voice-parallel kernel, in order on the mix stream.
int mix_core(gas_ctx *ctx, int n_voices, const gas_voice *d_voices, const gas_frame *d_src, int src_rows, int src_stride, int frames,
		gas_frame *d_bus, gas_frame *d_peaks) {
	if (ctx->planned.valid) {
		return gas_fail(ctx, GAS_ERR_STATE, "a block planned by gas_step_device is waiting to be streamed: finish the pipelined run first (gas_step_device with next = NULL)");
	}
	int st = join_voice_stream(ctx);
	if (st) {
		return st;
	}
	if (ctx->gain_pending) {
		GAS_CUDA(ctx, cudaStreamWaitEvent(ctx->s_mix, ctx->ev_gain_done, 0));
	}
	if (ctx->comm_pending) { // an exchange on the exchange stream may still read / write bus buffers (inside a capture this is a graph edge)
		GAS_CUDA(ctx, cudaStreamWaitEvent(ctx->s_mix, ctx->ev_comm_done, 0));
		ctx->comm_pending = false;
	}
	gas_ctx::ProfPair *pp = prof_open(ctx, GAS_KERNEL_PROLOGUE);
	GAS_CUDA(ctx, launch_plan(ctx, n_voices, d_voices, src_rows, frames, d_bus, d_peaks, ctx->s_mix));
	prof_close(ctx, pp);
	pp = prof_open(ctx, GAS_KERNEL_NONE); // the timer's own reading: a pair with nothing between its two records
	prof_close(ctx, pp);
	GAS_CUDA(ctx, cudaEventRecord(ctx->ev_prologue_done, ctx->s_mix));
	ctx->prologue_pending = true;
	// every planned block gets exactly one launch of the step kernel and of the voice-parallel kernel (they count blocks on the device)
	pp = prof_open(ctx, GAS_KERNEL_MIX_STREAM);
	GAS_CUDA(ctx, launch_step(ctx, d_src, src_stride, frames, d_bus, nullptr, ctx->s_mix, (ctx->pdl & 2) != 0));
	prof_close(ctx, pp);
	pp = prof_open(ctx, GAS_KERNEL_MIX_VOICE);
	GAS_CUDA(ctx, launch_mix_voice(ctx, d_src, src_stride, frames, d_bus, d_peaks, ctx->s_mix, true));
	prof_close(ctx, pp);
	GAS_CUDA(ctx, cudaEventRecord(ctx->ev_mix_done, ctx->s_mix)); // inside a capture: the edge a later exchange of this block hangs on
	ctx->mix_pending = true;
	return GAS_OK;
}

// Pipelined form: one launch of the step kernel streams the planned block and prepares the next one on its control warps; the
// voice-parallel kernel of the streamed block runs on the side stream, off the chain of step kernels.
int step_core(gas_ctx *ctx, const gas_frame *d_src, int src_stride, const StepNext *next) {
	const bool have = ctx->planned.valid;
	if (!have && !next) {
		return GAS_OK;
	}
	if (ctx->gain_pending) {
		GAS_CUDA(ctx, cudaStreamWaitEvent(ctx->s_mix, ctx->ev_gain_done, 0));
	}
	// an exchange still in flight matters only if it reads the buffer this call clears (next->d_bus) or streams into
	for (int r = 0; r < GAS_PLAN_DEPTH; r++) {
		if (ctx->comm_ring_valid[r] && ((next && ctx->comm_src[r] == next->d_bus) || (have && ctx->comm_src[r] == ctx->planned.bus))) {
			GAS_CUDA(ctx, cudaStreamWaitEvent(ctx->s_mix, ctx->ev_comm_ring[r], 0));
			ctx->comm_ring_valid[r] = false;
		}
	}
	constexpr int D = GAS_PLAN_DEPTH;
	const int i = (int)(ctx->step_count % D), im1 = (int)((ctx->step_count + D - 1) % D), im2 = (int)((ctx->step_count + D - 2) % D);
	if (!have) {
		// first block of a run: gains and plan by the stand-alone kernels, in order on the mix stream
		int st = join_voice_stream(ctx);
		if (st) {
			return st;
		}
		if (next->n_emitters > 0) {
			gas_ctx::ProfPair *pg = prof_open(ctx, GAS_KERNEL_GAIN);
			GAS_CUDA(ctx, launch_gain(ctx, next->n_emitters, next->d_emitters, ctx->n_listeners_res, ctx->d_listeners, ctx->d_areas, nullptr, ctx->s_mix));
			prof_close(ctx, pg);
		}
		gas_ctx::ProfPair *pp = prof_open(ctx, GAS_KERNEL_PROLOGUE);
		GAS_CUDA(ctx, launch_plan(ctx, next->n_voices, next->d_voices, next->src_rows, next->frames, next->d_bus, next->d_peaks, ctx->s_mix));
		prof_close(ctx, pp);
		GAS_CUDA(ctx, cudaEventRecord(ctx->ev_step_done[im1], ctx->s_mix));
		ctx->step_done_valid = true;
	} else {
		// plan slots and output buffers about to be rewritten by this launch's control warps must be free: the voice-parallel kernel
		// of the block before last (its plan slot comes up for clearing), and the last one's if the next block reuses its buffers
		if (ctx->block_inflight[im2]) {
			GAS_CUDA(ctx, cudaStreamWaitEvent(ctx->s_mix, ctx->ev_block_done[im2], 0));
			ctx->block_inflight[im2] = false;
		}
		if (next && ctx->block_inflight[im1] && (ctx->inflight_bus[im1] == next->d_bus || (next->d_peaks && ctx->inflight_peaks[im1] == next->d_peaks))) {
			GAS_CUDA(ctx, cudaStreamWaitEvent(ctx->s_mix, ctx->ev_block_done[im1], 0));
			ctx->block_inflight[im1] = false;
		}
		// the voice-parallel kernel of this block: its plan is complete once the previous step kernel (or the priming plan) is
		if (ctx->step_done_valid) {
			GAS_CUDA(ctx, cudaStreamWaitEvent(ctx->s_voice, ctx->ev_step_done[im1], 0));
		}
		gas_ctx::ProfPair *pp = prof_open(ctx, GAS_KERNEL_MIX_VOICE, ctx->s_voice);
		GAS_CUDA(ctx, launch_mix_voice(ctx, d_src, src_stride, ctx->planned.frames, ctx->planned.bus, ctx->planned.peaks, ctx->s_voice, false));
		prof_close(ctx, pp);
		pp = prof_open(ctx, GAS_KERNEL_MIX_STREAM);
		GAS_CUDA(ctx, launch_step(ctx, d_src, src_stride, ctx->planned.frames, ctx->planned.bus, next, ctx->s_mix, (ctx->pdl & 8) != 0));
		prof_close(ctx, pp);
		GAS_CUDA(ctx, cudaEventRecord(ctx->ev_step_done[i], ctx->s_mix));
		ctx->step_done_valid = true;
		GAS_CUDA(ctx, cudaStreamWaitEvent(ctx->s_voice, ctx->ev_step_done[i], 0));
		GAS_CUDA(ctx, cudaEventRecord(ctx->ev_block_done[i], ctx->s_voice));
		ctx->block_inflight[i] = true;
		ctx->inflight_bus[i] = ctx->planned.bus;
		ctx->inflight_peaks[i] = ctx->planned.peaks;
		ctx->step_count++;
	}
	GAS_CUDA(ctx, cudaEventRecord(ctx->ev_prologue_done, ctx->s_mix)); // gain-side calls wait for the control warps
	ctx->prologue_pending = true;
	ctx->planned.valid = next != nullptr;
	if (next) {
		ctx->planned.n_voices = next->n_voices;
		ctx->planned.frames = next->frames;
		ctx->planned.src_rows = next->src_rows;
		ctx->planned.bus = next->d_bus;
		ctx->planned.peaks = next->d_peaks;
	}
	return GAS_OK;
}

} // namespace

extern "C" {

int gas_abi_version(void) { return GAS_ABI_VERSION; }

size_t gas_abi_sizeof(int32_t id) {
	switch (id) {
		case GAS_STRUCT_FRAME: return sizeof(gas_frame);
		case GAS_STRUCT_EFFECT: return sizeof(gas_effect);
		case GAS_STRUCT_EFFECT_CHAIN: return sizeof(gas_effect_chain);
		case GAS_STRUCT_SPATIALIZER: return sizeof(gas_spatializer);
		case GAS_STRUCT_LISTENER: return sizeof(gas_listener);
		case GAS_STRUCT_AREA: return sizeof(gas_area);
		case GAS_STRUCT_EMITTER: return sizeof(gas_emitter);
		case GAS_STRUCT_PARAMS: return sizeof(gas_params);
		case GAS_STRUCT_VOICE: return sizeof(gas_voice);
		case GAS_STRUCT_PROCESSOR_STATE: return sizeof(gas_processor_state);
		case GAS_STRUCT_VOICE_STATE: return sizeof(gas_voice_state);
		case GAS_STRUCT_CONFIG: return sizeof(gas_config);
		case GAS_STRUCT_VOICE_LIFE: return sizeof(gas_voice_life);
		case GAS_STRUCT_BUS_DESC: return sizeof(gas_bus_desc);
		case GAS_STRUCT_STEP_NEXT: return sizeof(gas_step_next);
		default: return 0;
	}
}

void gas_config_defaults(gas_config *c) {
	if (!c) {
		return;
	}
	c->device = 0;
	c->max_instances = 1024;
	c->max_voices = 1024;
	c->max_frames = 512; // upstream AudioServer buffer size
	c->max_spatializers = 16;
	c->num_buses = 2;
	c->speaker_mode = GAS_SPEAKER_MODE_STEREO;
	c->mix_rate = 44100.0f;
	c->global_panning_strength = 0.5f;
}

void gas_spatializer_defaults(gas_spatializer *s) { // reference audio_spatializer_3d.h:171-188
	if (!s) {
		return;
	}
	memset(s, 0, sizeof(*s));
	s->kind = GAS_SPATIALIZER_3D;
	s->attenuation_model = GAS_ATTENUATION_INVERSE_DISTANCE;
	s->unit_size = 10.0f;
	s->max_distance = 0.0f;
	s->panning_strength = 1.0f;
	s->area_mask = 1;
	s->emission_angle_enabled = 0;
	s->emission_angle = 45.0f;
	s->emission_angle_filter_attenuation_db = -12.0f;
	s->attenuation_filter_cutoff_hz = 5000.0f;
	s->attenuation_filter_db = -24.0f;
	s->doppler_tracking = GAS_DOPPLER_TRACKING_DISABLED;
	s->doppler_speed_of_sound = 343.0f;
	s->mix_channel_mode = 0;
	s->effect_gain_binding = -1;
}

const char *gas_last_error(const gas_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int gas_create(const gas_config *cfg, gas_ctx **out) {
	if (out) {
		*out = nullptr;
	}
	if (!cfg || !out) {
		return gas_fail(nullptr, GAS_ERR_INVALID, "gas_create: null argument");
	}
	if (cfg->max_instances <= 0 || cfg->max_voices <= 0 || cfg->max_frames < 2 || (cfg->max_frames & 1) || cfg->max_spatializers <= 0 ||
			cfg->num_buses < 1 || cfg->num_buses > GAS_MAX_BUSES || cfg->speaker_mode < 0 || cfg->speaker_mode > 3 || !(cfg->mix_rate > 0.f)) {
		return gas_fail(nullptr, GAS_ERR_INVALID, "gas_create: invalid configuration");
	}
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev <= 0) {
		return gas_fail(nullptr, GAS_ERR_NO_DEVICE, "gas_create: no CUDA device (%s); this library has no CPU fallback",
				e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
	}
	if (cfg->device < 0 || cfg->device >= ndev) {
		return gas_fail(nullptr, GAS_ERR_INVALID, "gas_create: device %d out of range (%d devices)", cfg->device, ndev);
	}
	cudaDeviceProp prop;
	e = cudaGetDeviceProperties(&prop, cfg->device);
	if (e != cudaSuccess) {
		return gas_fail(nullptr, GAS_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
	}
	if (prop.major != 10) {
		return gas_fail(nullptr, GAS_ERR_NO_DEVICE, "gas_create: device %d is sm_%d%d; the kernels are built for sm_100a only", cfg->device,
				prop.major, prop.minor);
	}
	e = cudaSetDevice(cfg->device);
	if (e != cudaSuccess) {
		return gas_fail(nullptr, GAS_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
	}
	gas_ctx *ctx = new (std::nothrow) gas_ctx();
	if (!ctx) {
		return gas_fail(nullptr, GAS_ERR_NOMEM, "gas_create: out of host memory");
	}
	ctx->cfg = *cfg;
	ctx->device = cfg->device;
	ctx->num_sms = prop.multiProcessorCount;
	ctx->l2_bytes = prop.l2CacheSize;
	refresh_globals(ctx);
	{
		// programmatic dependent launch of 1 = planner kernel, 2 = step kernel of the block-call form, 4 = voice-parallel kernel of the
		// block-call form, 8 = step kernel of the pipelined form.  Default: 4 (the idle case of the voice-parallel kernel — no filtered
		// voice in the block — then costs no launch latency) and 8 (consecutive step kernels overlap: measured 37.5 -> 36.1 us per step).
		const char *e = getenv("GAS_PDL");
		ctx->pdl = e ? atoi(e) : 12;
		// The voice-parallel kernel needs nothing the streaming kernel produces (both add into the bus buffers with
		// reductions), so with GAS_K3_PARALLEL=1 it runs beside it on its own stream and K2 adds straight into the bus
		// buffers (1 replica; more replicas are folded by the K3 launch, which then has to wait for K2).  Measured on
		// B200: blocks with real K3 work gain 2-15 % (effect chains 191 -> 179 us and 206 -> 186 us, Mode A + filter
		// 170 -> 145 us); a block whose K3 has nothing to do loses 3 us, because the idle K3 CTAs hold the SMs K2's
		// one-CTA-per-SM grid is waiting for (41.7 -> 44.9 us; K2 enqueued first is worse: 51 us).  Whether a block
		// has K3 work is decided on the device (per-voice filter gain), so the host cannot choose per block: off by
		// default, for the caller to turn on for filter / effect-chain heavy scenes.
		e = getenv("GAS_K3_PARALLEL");
		ctx->par_voice = e && atoi(e) != 0;
		e = getenv("GAS_K1_GATE");
		ctx->gain_after_stream = !(e && atoi(e) == 0);
		e = getenv("GAS_K2_SCALED");
		ctx->scaled_classes = !(e && atoi(e) == 0);
		e = getenv("GAS_K3_LEGACY"); // experiments: the round-1 per-warp form of the voice-parallel kernel for every class
		ctx->k3_legacy = e && atoi(e) != 0;
		ctx->replicas = 1; // the step kernel adds straight into the bus buffers (8 replicas + fold measured no faster)
		ctx->par_voice = false;
		e = getenv("GAS_K2_DEBUG"); // the timeline buffer must exist before anything is captured into a graph
		if (e && (atoi(e) & 8)) {
			cudaMalloc((void **)&ctx->d_timeline, 256 * 32 * sizeof(unsigned long long));
		}
		e = getenv("GAS_SKIP"); // experiments only: 1 = no prologue, 2 = no streaming kernel, 4 = no voice-parallel kernel
		ctx->skip = e ? atoi(e) : 0;
	}

	const size_t I = cfg->max_instances, V = cfg->max_voices, F = cfg->max_frames;
	const size_t nscratch_ids = (I > V ? I : V);
	ctx->max_areas = 1024;
	ctx->scratch_bytes = V * sizeof(gas_voice_state);
	if (ctx->scratch_bytes < I * sizeof(gas_params)) {
		ctx->scratch_bytes = I * sizeof(gas_params);
	}
	bool ok = true;
#define ALLOC(ptr, n) ok = ok && (dev_alloc(&(ptr), (n)) == cudaSuccess)
	ALLOC(ctx->t.spat, (size_t)cfg->max_spatializers);
	ALLOC(ctx->t.inst_spat, I);
	ALLOC(ctx->t.inst_params, I);
	ALLOC(ctx->t.inst_was_further, I);
	ALLOC(ctx->t.inst_active, I);
	ALLOC(ctx->t.inst_cur, I);
	ALLOC(ctx->t.inst_prev, 2 * I);
	ALLOC(ctx->t.inst_mode, I);
	ALLOC(ctx->t.inst_seq, I);
	ALLOC(ctx->t.blk, (size_t)BLK_WORDS);
	ctx->t.max_instances = (int32_t)I;
	ALLOC(ctx->t.inst_fx, I);
	ALLOC(ctx->t.vs_prev, V * 8);
	ALLOC(ctx->t.vs_proc, V * 8);
	ALLOC(ctx->t.vs_fx, V * (size_t)(GAS_MAX_EFFECTS * 2 * GAS_MAX_FILTER_STAGES * 4));
	ALLOC(ctx->t.vs_look, V * (size_t)GAS_LOOKAHEAD_BUFFER_SIZE);
	ALLOC(ctx->t.vs_life, V);
	ALLOC(ctx->t.inst_threshold, I);
	ALLOC(ctx->t.vs_src, V);
	ALLOC(ctx->t.vs_start, V);
	ALLOC(ctx->t.vs_pos, V);
	for (int b = 0; b < GAS_MAX_BUSES; b++) {
		ctx->bus_volume_lin[b] = 1.0f;
		ctx->bus_send[b] = 0;
	}
	ctx->max_sources = 4096;
	ALLOC(ctx->d_sources, (size_t)ctx->max_sources);
	ctx->h_sources.assign(ctx->max_sources, SourceDesc{});
	ALLOC(ctx->d_rs_rows, V * F);
	ALLOC(ctx->d_rs_mixed, V);
	ctx->t.max_voices = (int32_t)V;
	ctx->t.threshold_default = expf(-80.0f * (float)0.11512925464970228420089957273422); // upstream Math::db_to_linear(float)
	ALLOC(ctx->d_stage, V * F + 2 * F); // + in/out rows of the per-call entry points
	ALLOC(ctx->d_voices_stage, V);
	ALLOC(ctx->d_mixed, V);
	ALLOC(ctx->d_status, V);
	ALLOC(ctx->plan.cls_key, (size_t)GAS_MAX_CLASSES);
	ALLOC(ctx->plan.cls_aux, (size_t)GAS_MAX_CLASSES);
	ALLOC(ctx->plan.cls_idle, (size_t)GAS_MAX_CLASSES);
	ALLOC(ctx->plan.cls_count, (size_t)GAS_PLAN_DEPTH * GAS_MAX_CLASSES);
	ALLOC(ctx->plan.overflow, (size_t)1);
	ALLOC(ctx->plan.list, (size_t)GAS_PLAN_DEPTH * GAS_MAX_CLASSES * V);
	ALLOC(ctx->plan.k2_rows, (size_t)GAS_PLAN_ROW_DEPTH * GAS_MAX_CLASSES * V * GAS_K2_ROW_FLOATS + 64);
	ALLOC(ctx->plan.rec, (size_t)GAS_PLAN_DEPTH * V);
	ALLOC(ctx->plan.sends, (size_t)GAS_PLAN_DEPTH * V);
	ALLOC(ctx->plan.hdr, (size_t)GAS_PLAN_DEPTH);
	ALLOC(ctx->d_voices, V);
	ALLOC(ctx->d_src, V * F);
	ALLOC(ctx->d_bus, (size_t)GAS_MAX_BUSES * GAS_MAX_CHANNELS_PER_BUS * F);
	ALLOC(ctx->d_peaks, V);
	ALLOC(ctx->d_rep, (size_t)16 * GAS_MAX_BUSES * GAS_MAX_CHANNELS_PER_BUS * F);
	ALLOC(ctx->d_emitters, I);
	ALLOC(ctx->d_listeners, (size_t)GAS_MAX_LISTENERS);
	ALLOC(ctx->d_listener_pre, (size_t)GAS_MAX_LISTENERS);
	ALLOC(ctx->d_areas, (size_t)ctx->max_areas);
	ALLOC(ctx->d_params_out, I);
	ALLOC(ctx->d_ids, nscratch_ids);
	ALLOC(ctx->d_ids2, nscratch_ids);
	ALLOC(ctx->d_ids_mix, nscratch_ids);
	{
		unsigned char *p = nullptr;
		ok = ok && (dev_alloc(&p, ctx->scratch_bytes) == cudaSuccess);
		ctx->d_scratch = p;
		// Staging is per stream: the gain stream (instance_init/start/stop, params_set/get, effect_params_set) and the mix
		// stream (voice_init, state / lifecycle export and import, thresholds) are not ordered against each other, so a
		// shared buffer could be overwritten by one while a kernel of the other has not read it yet.
		size_t mix_bytes = V * (sizeof(gas_voice_state) > sizeof(gas_voice_life) ? sizeof(gas_voice_state) : sizeof(gas_voice_life));
		if (mix_bytes < I * sizeof(float)) {
			mix_bytes = I * sizeof(float);
		}
		p = nullptr;
		ok = ok && (dev_alloc(&p, mix_bytes) == cudaSuccess);
		ctx->d_scratch_mix = p;
	}
#undef ALLOC
	ok = ok && cudaStreamCreateWithFlags(&ctx->s_mix, cudaStreamNonBlocking) == cudaSuccess;
	ok = ok && cudaStreamCreateWithFlags(&ctx->s_gain, cudaStreamNonBlocking) == cudaSuccess;
	ok = ok && cudaStreamCreateWithFlags(&ctx->s_comm, cudaStreamNonBlocking) == cudaSuccess;
	ok = ok && cudaStreamCreateWithFlags(&ctx->s_voice, cudaStreamNonBlocking) == cudaSuccess;
	ok = ok && cudaEventCreateWithFlags(&ctx->ev_stream_started, cudaEventDisableTiming) == cudaSuccess;
	ok = ok && cudaEventCreateWithFlags(&ctx->ev_voice_fork, cudaEventDisableTiming) == cudaSuccess;
	ok = ok && cudaEventCreateWithFlags(&ctx->ev_voice_join, cudaEventDisableTiming) == cudaSuccess;
	ok = ok && cudaEventCreateWithFlags(&ctx->ev_mix_done, cudaEventDisableTiming) == cudaSuccess;
	ok = ok && cudaEventCreateWithFlags(&ctx->ev_comm_done, cudaEventDisableTiming) == cudaSuccess;
	ok = ok && cudaEventCreateWithFlags(&ctx->ev_join2, cudaEventDisableTiming) == cudaSuccess;
	ok = ok && cudaEventCreateWithFlags(&ctx->ev_gain_done, cudaEventDisableTiming) == cudaSuccess;
	ok = ok && cudaEventCreateWithFlags(&ctx->ev_prologue_done, cudaEventDisableTiming) == cudaSuccess;
	ok = ok && cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) == cudaSuccess;
	ok = ok && cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) == cudaSuccess;
	for (int i = 0; i < GAS_PLAN_DEPTH; i++) {
		ok = ok && cudaEventCreateWithFlags(&ctx->ev_step_done[i], cudaEventDisableTiming) == cudaSuccess;
		ok = ok && cudaEventCreateWithFlags(&ctx->ev_block_done[i], cudaEventDisableTiming) == cudaSuccess;
		ok = ok && cudaEventCreateWithFlags(&ctx->ev_comm_ring[i], cudaEventDisableTiming) == cudaSuccess;
	}
	if (ok) {
		ok = launch_defaults(ctx, ctx->s_gain) == cudaSuccess && cudaStreamSynchronize(ctx->s_gain) == cudaSuccess;
	}
	if (!ok) {
		cudaError_t le = cudaGetLastError();
		int st = gas_fail(nullptr, le == cudaErrorMemoryAllocation ? GAS_ERR_NOMEM : GAS_ERR_CUDA, "gas_create: device setup failed (%s)",
				cudaGetErrorString(le));
		gas_destroy(ctx);
		return st;
	}
	*out = ctx;
	return GAS_OK;
}

void gas_destroy(gas_ctx *ctx) {
	if (!ctx) {
		return;
	}
	cudaSetDevice(ctx->device);
	if (ctx->s_mix) {
		cudaStreamSynchronize(ctx->s_mix);
	}
	if (ctx->s_gain) {
		cudaStreamSynchronize(ctx->s_gain);
	}
	if (ctx->s_comm) {
		cudaStreamSynchronize(ctx->s_comm);
	}
	gas_comm_close(ctx);
	void *ptrs[] = { ctx->t.spat, ctx->t.inst_spat, ctx->t.inst_params, ctx->t.inst_was_further, ctx->t.inst_active, ctx->t.inst_cur,
		ctx->t.inst_prev, ctx->t.inst_mode, ctx->t.blk, ctx->t.inst_fx, ctx->plan.sends, ctx->t.vs_prev, ctx->t.vs_proc, ctx->t.vs_fx, ctx->plan.cls_key, ctx->plan.cls_aux, ctx->plan.cls_idle, ctx->plan.cls_count,
		ctx->plan.overflow, ctx->plan.list, ctx->plan.k2_rows, ctx->plan.rec, ctx->d_voices, ctx->d_src, ctx->d_bus,
		ctx->d_peaks, ctx->d_rep, ctx->d_emitters, ctx->d_listeners, ctx->d_areas, ctx->d_params_out, ctx->d_ids, ctx->d_ids2, ctx->d_scratch,
		ctx->d_exchange, ctx->d_comm_seq, ctx->d_comm_ticket, ctx->t.vs_look, ctx->t.vs_life, ctx->t.inst_threshold, ctx->d_stage, ctx->d_voices_stage,
		ctx->d_mixed, ctx->d_status, ctx->d_ids_mix, ctx->d_scratch_mix, ctx->plan.hdr, ctx->d_listener_pre, ctx->t.inst_seq, ctx->t.vs_src, ctx->t.vs_start, ctx->t.vs_pos, ctx->d_sources, ctx->d_rs_rows, ctx->d_rs_mixed };
	for (void *p : ptrs) {
		if (p) {
			cudaFree(p);
		}
	}
	for (auto &sd : ctx->h_sources) {
		if (sd.pcm) {
			cudaFree((void *)sd.pcm);
		}
	}
	if (ctx->ev_gain_done) {
		cudaEventDestroy(ctx->ev_gain_done);
	}
	if (ctx->ev_prologue_done) {
		cudaEventDestroy(ctx->ev_prologue_done);
	}
	if (ctx->ev_fork) {
		cudaEventDestroy(ctx->ev_fork);
	}
	if (ctx->ev_join) {
		cudaEventDestroy(ctx->ev_join);
	}
	if (ctx->s_voice) {
		cudaStreamSynchronize(ctx->s_voice);
		cudaStreamDestroy(ctx->s_voice);
	}
	for (cudaEvent_t e : { ctx->ev_mix_done, ctx->ev_comm_done, ctx->ev_join2, ctx->ev_voice_fork, ctx->ev_voice_join, ctx->ev_stream_started }) {
		if (e) {
			cudaEventDestroy(e);
		}
	}
	for (int i = 0; i < GAS_PLAN_DEPTH; i++) {
		if (ctx->ev_step_done[i]) {
			cudaEventDestroy(ctx->ev_step_done[i]);
		}
		if (ctx->ev_block_done[i]) {
			cudaEventDestroy(ctx->ev_block_done[i]);
		}
		if (ctx->ev_comm_ring[i]) {
			cudaEventDestroy(ctx->ev_comm_ring[i]);
		}
	}
	if (ctx->s_comm) {
		cudaStreamDestroy(ctx->s_comm);
	}

	for (auto &g : ctx->graphs) {
		if (g.exec) {
			cudaGraphExecDestroy(g.exec);
		}
	}
	for (auto &pp : ctx->prof_pairs) {
		cudaEventDestroy(pp.a);
		cudaEventDestroy(pp.b);
	}
	for (int k = 0; k < GAS_KERNEL_KINDS; k++) {
		for (int e = 0; e < 2; e++) {
			if (ctx->gev[k][e]) {
				cudaEventDestroy(ctx->gev[k][e]);
			}
		}
	}
	if (ctx->s_mix) {
		cudaStreamDestroy(ctx->s_mix);
	}
	if (ctx->s_gain) {
		cudaStreamDestroy(ctx->s_gain);
	}
	delete ctx;
}

#define ENTER(ctx)                                                          \
	if (!(ctx)) {                                                           \
		return gas_fail(nullptr, GAS_ERR_INVALID, "%s: null context", __func__); \
	}                                                                       \
	std::lock_guard<std::mutex> _lk((ctx)->mu);                             \
	GAS_CUDA(ctx, cudaSetDevice((ctx)->device))

int gas_set_speaker_mode(gas_ctx *ctx, int32_t mode) {
	ENTER(ctx);
	if (mode < 0 || mode > 3) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_set_speaker_mode: mode %d out of range", mode);
	}
	ctx->cfg.speaker_mode = mode;
	refresh_globals(ctx);
	return GAS_OK;
}

int gas_set_mix_rate(gas_ctx *ctx, float hz) {
	ENTER(ctx);
	if (!(hz > 0.f)) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_set_mix_rate: rate must be positive");
	}
	ctx->cfg.mix_rate = hz;
	refresh_globals(ctx);
	return GAS_OK;
}

int gas_set_global_panning_strength(gas_ctx *ctx, float s) {
	ENTER(ctx);
	ctx->cfg.global_panning_strength = s;
	refresh_globals(ctx);
	return GAS_OK;
}

int gas_get_channel_count(const gas_ctx *ctx) { return ctx ? ctx->cfg.speaker_mode + 1 : 0; }

int gas_spatializer_set(gas_ctx *ctx, int32_t slot, const gas_spatializer *s) {
	ENTER(ctx);
	if (slot < 0 || slot >= ctx->cfg.max_spatializers || !s) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_spatializer_set: bad slot or null");
	}
	if (!spat_valid(s)) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_spatializer_set: property out of range (max_distance>=0, emission_angle in [0,90], attenuation_model<4, panning_strength>=0, doppler_speed_of_sound>0)");
	}
	int st = gain_side_begin(ctx);
	if (st) {
		return st;
	}
	GAS_CUDA(ctx, cudaMemcpyAsync(ctx->t.spat + slot, s, sizeof(*s), cudaMemcpyHostToDevice, ctx->s_gain));
	return gain_side_end(ctx);
}

int gas_instance_init(gas_ctx *ctx, int32_t n, const int32_t *instances, const int32_t *spatializers) {
	ENTER(ctx);
	if (n < 0 || n > ctx->cfg.max_instances || (n > 0 && (!instances || !spatializers)) || !ids_valid(instances, n, ctx->cfg.max_instances) ||
			!ids_valid(spatializers, n, ctx->cfg.max_spatializers)) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_instance_init: bad instance or spatializer slot");
	}
	int st = gain_side_begin(ctx);
	if (st) {
		return st;
	}
	GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_ids, instances, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->s_gain));
	GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_ids2, spatializers, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->s_gain));
	GAS_CUDA(ctx, launch_instance_init(ctx, n, ctx->d_ids, ctx->d_ids2, ctx->s_gain));
	for (int i = 0; i < n; i++) {
		if (instances[i] + 1 > ctx->inst_hwm) {
			ctx->inst_hwm = instances[i] + 1;
		}
	}
	return gain_side_end(ctx);
}

static int instance_list_op(gas_ctx *ctx, int32_t n, const int32_t *instances, bool start, const char *who) {
	if (n < 0 || n > ctx->cfg.max_instances || (n > 0 && !instances) || !ids_valid(instances, n, ctx->cfg.max_instances)) {
		return gas_fail(ctx, GAS_ERR_INVALID, "%s: bad instance slot", who);
	}
	int st = gain_side_begin(ctx);
	if (st) {
		return st;
	}
	GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_ids, instances, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->s_gain));
	if (start) {
		GAS_CUDA(ctx, launch_instance_start(ctx, n, ctx->d_ids, ctx->s_gain));
		for (int i = 0; i < n; i++) {
			if (instances[i] + 1 > ctx->inst_hwm) {
				ctx->inst_hwm = instances[i] + 1;
			}
		}
	} else {
		GAS_CUDA(ctx, launch_instance_stop(ctx, n, ctx->d_ids, ctx->s_gain));
	}
	return gain_side_end(ctx);
}

int gas_instance_start(gas_ctx *ctx, int32_t n, const int32_t *instances) {
	ENTER(ctx);
	return instance_list_op(ctx, n, instances, true, "gas_instance_start");
}

int gas_instance_stop(gas_ctx *ctx, int32_t n, const int32_t *instances) {
	ENTER(ctx);
	return instance_list_op(ctx, n, instances, false, "gas_instance_stop");
}

int gas_voice_init(gas_ctx *ctx, int32_t n, const int32_t *voices) {
	ENTER(ctx);
	if (n < 0 || n > ctx->cfg.max_voices || (n > 0 && !voices) || !ids_valid(voices, n, ctx->cfg.max_voices)) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_voice_init: bad voice slot");
	}
	GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_ids_mix, voices, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->s_mix));
	GAS_CUDA(ctx, launch_voice_init(ctx, n, ctx->d_ids_mix, ctx->s_mix));
	return GAS_OK;
}

static int gain_common(gas_ctx *ctx, int32_t n, const gas_emitter *d_em, int32_t n_listeners, const gas_listener *listeners, int32_t n_areas,
		const gas_area *areas, gas_params *d_out) {
	if (listeners) {
		if (n_listeners < 0 || n_listeners > GAS_MAX_LISTENERS) {
			return gas_fail(ctx, GAS_ERR_INVALID, "gas_gain_compute: 0..%d listeners", GAS_MAX_LISTENERS);
		}
		if (ctx->capturing) {
			return gas_fail(ctx, GAS_ERR_STATE, "gas_gain_compute: host listeners cannot be uploaded while capturing; use gas_listeners_set");
		}
		if (n_listeners > 0) {
			GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_listeners, listeners, n_listeners * sizeof(gas_listener), cudaMemcpyHostToDevice, ctx->s_gain));
			GAS_CUDA(ctx, launch_listener_pre(ctx, n_listeners, ctx->s_gain));
		}
		ctx->n_listeners_res = n_listeners;
	} else {
		n_listeners = ctx->n_listeners_res; // resident copy (gas_listeners_set)
	}
	if (areas) {
		if (n_areas < 0 || n_areas > ctx->max_areas) {
			return gas_fail(ctx, GAS_ERR_INVALID, "gas_gain_compute: 0..%d areas", ctx->max_areas);
		}
		if (ctx->capturing) {
			return gas_fail(ctx, GAS_ERR_STATE, "gas_gain_compute: host areas cannot be uploaded while capturing; use gas_areas_set");
		}
		if (n_areas > 0) {
			GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_areas, areas, n_areas * sizeof(gas_area), cudaMemcpyHostToDevice, ctx->s_gain));
		}
		ctx->n_areas_res = n_areas;
	}
	gas_ctx::ProfPair *pp = prof_open(ctx, GAS_KERNEL_GAIN, ctx->s_gain);
	GAS_CUDA(ctx, launch_gain(ctx, n, d_em, n_listeners, ctx->d_listeners, ctx->d_areas, d_out, ctx->s_gain));
	prof_close(ctx, pp);
	return GAS_OK;
}

int gas_listeners_set(gas_ctx *ctx, int32_t n_listeners, const gas_listener *listeners) {
	ENTER(ctx);
	if (n_listeners < 0 || n_listeners > GAS_MAX_LISTENERS || (n_listeners > 0 && !listeners) || ctx->capturing) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_listeners_set: 0..%d listeners, not while capturing", GAS_MAX_LISTENERS);
	}
	int st = gain_side_begin(ctx);
	if (st) {
		return st;
	}
	if (n_listeners > 0) {
		GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_listeners, listeners, n_listeners * sizeof(gas_listener), cudaMemcpyHostToDevice, ctx->s_gain));
		GAS_CUDA(ctx, launch_listener_pre(ctx, n_listeners, ctx->s_gain));
	}
	ctx->n_listeners_res = n_listeners;
	return gain_side_end(ctx);
}

int gas_areas_set(gas_ctx *ctx, int32_t n_areas, const gas_area *areas) {
	ENTER(ctx);
	if (n_areas < 0 || n_areas > ctx->max_areas || (n_areas > 0 && !areas) || ctx->capturing) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_areas_set: 0..%d areas, not while capturing", ctx->max_areas);
	}
	int st = gain_side_begin(ctx);
	if (st) {
		return st;
	}
	if (n_areas > 0) {
		GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_areas, areas, n_areas * sizeof(gas_area), cudaMemcpyHostToDevice, ctx->s_gain));
	}
	ctx->n_areas_res = n_areas;
	return gain_side_end(ctx);
}

int gas_gain_compute(gas_ctx *ctx, int32_t n, const gas_emitter *emitters, int32_t n_listeners, const gas_listener *listeners,
		int32_t n_areas, const gas_area *areas, gas_params *out_params) {
	{
		ENTER(ctx);
		if (n < 0 || n > ctx->cfg.max_instances || (n > 0 && !emitters)) {
			return gas_fail(ctx, GAS_ERR_INVALID, "gas_gain_compute: 0..max_instances emitters");
		}
		const int32_t area_limit = areas ? n_areas : ctx->n_areas_res; // without an areas argument the resident copy is used
		for (int i = 0; i < n; i++) {
			const gas_emitter &e = emitters[i];
			if (e.instance < 0 || e.instance >= ctx->cfg.max_instances || e.spatializer < 0 || e.spatializer >= ctx->cfg.max_spatializers ||
					e.area >= area_limit) {
				return gas_fail(ctx, GAS_ERR_INVALID, "gas_gain_compute: emitter %d references a bad instance/spatializer/area", i);
			}
		}
		int st = gain_side_begin(ctx);
		if (st) {
			return st;
		}
		GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_emitters, emitters, n * sizeof(gas_emitter), cudaMemcpyHostToDevice, ctx->s_gain));
		st = gain_common(ctx, n, ctx->d_emitters, n_listeners, listeners, n_areas, areas, out_params ? ctx->d_params_out : nullptr);
		if (st) {
			return st;
		}
		if (out_params && n > 0) {
			GAS_CUDA(ctx, cudaMemcpyAsync(out_params, ctx->d_params_out, n * sizeof(gas_params), cudaMemcpyDeviceToHost, ctx->s_gain));
		}
		st = gain_side_end(ctx);
		if (st) {
			return st;
		}
	}
	if (out_params) {
		GAS_CUDA(ctx, cudaStreamSynchronize(ctx->s_gain));
	}
	return GAS_OK;
}

int gas_gain_compute_device(gas_ctx *ctx, int32_t n, const gas_emitter *d_emitters, int32_t n_listeners, const gas_listener *listeners,
		int32_t n_areas, const gas_area *areas, gas_params *d_out_params) {
	ENTER(ctx);
	if (n < 0 || n > ctx->cfg.max_instances || (n > 0 && !d_emitters)) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_gain_compute_device: 0..max_instances emitters");
	}
	int st = gain_side_begin(ctx);
	if (st) {
		return st;
	}
	st = gain_common(ctx, n, d_emitters, n_listeners, listeners, n_areas, areas, d_out_params);
	if (st) {
		return st;
	}
	return gain_side_end(ctx);
}

int gas_params_set(gas_ctx *ctx, int32_t n, const int32_t *instances, const gas_params *params) {
	ENTER(ctx);
	if (n < 0 || n > ctx->cfg.max_instances || (n > 0 && (!instances || !params)) || !ids_valid(instances, n, ctx->cfg.max_instances)) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_params_set: bad instance slot");
	}
	for (int i = 0; i < n; i++) {
		if (params[i].n_bus < 0 || params[i].n_bus > GAS_MAX_BUSES_PER_PLAYBACK) {
			return gas_fail(ctx, GAS_ERR_INVALID, "gas_params_set: n_bus must be 0..%d", GAS_MAX_BUSES_PER_PLAYBACK);
		}
	}
	int st = gain_side_begin(ctx);
	if (st) {
		return st;
	}
	GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_ids, instances, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->s_gain));
	GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_scratch, params, n * sizeof(gas_params), cudaMemcpyHostToDevice, ctx->s_gain));
	GAS_CUDA(ctx, launch_params_set(ctx, n, ctx->d_ids, (const gas_params *)ctx->d_scratch, ctx->s_gain));
	return gain_side_end(ctx);
}

int gas_params_get(gas_ctx *ctx, int32_t n, const int32_t *instances, gas_params *out) {
	{
		ENTER(ctx);
		if (n < 0 || n > ctx->cfg.max_instances || (n > 0 && (!instances || !out)) || !ids_valid(instances, n, ctx->cfg.max_instances)) {
			return gas_fail(ctx, GAS_ERR_INVALID, "gas_params_get: bad instance slot");
		}
		GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_ids, instances, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->s_gain));
		GAS_CUDA(ctx, launch_params_get(ctx, n, ctx->d_ids, (gas_params *)ctx->d_scratch, ctx->s_gain));
		if (n > 0) {
			GAS_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_scratch, n * sizeof(gas_params), cudaMemcpyDeviceToHost, ctx->s_gain));
		}
	}
	GAS_CUDA(ctx, cudaStreamSynchronize(ctx->s_gain));
	return GAS_OK;
}

int gas_effect_params_set(gas_ctx *ctx, int32_t n, const int32_t *instances, const gas_effect_chain *chains) {
	ENTER(ctx);
	if (n < 0 || n > ctx->cfg.max_instances || (n > 0 && (!instances || !chains)) || !ids_valid(instances, n, ctx->cfg.max_instances)) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_effect_params_set: bad instance slot");
	}
	for (int i = 0; i < n; i++) {
		if (chains[i].n_effects < 0 || chains[i].n_effects > GAS_MAX_EFFECTS) {
			return gas_fail(ctx, GAS_ERR_INVALID, "gas_effect_params_set: n_effects must be 0..%d", GAS_MAX_EFFECTS);
		}
	}
	int st = gain_side_begin(ctx);
	if (st) {
		return st;
	}
	GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_ids, instances, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->s_gain));
	GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_scratch, chains, n * sizeof(gas_effect_chain), cudaMemcpyHostToDevice, ctx->s_gain));
	GAS_CUDA(ctx, launch_fx_set(ctx, n, ctx->d_ids, (const gas_effect_chain *)ctx->d_scratch, ctx->s_gain));
	return gain_side_end(ctx);
}

int gas_mix_block(gas_ctx *ctx, int32_t n_voices, const gas_voice *voices, const gas_frame *src, int32_t src_rows, int32_t frames,
		gas_frame *bus_out, gas_frame *peaks) {
	{
		ENTER(ctx);
		if (n_voices < 0 || n_voices > ctx->cfg.max_voices || (n_voices > 0 && !voices)) {
			return gas_fail(ctx, GAS_ERR_INVALID, "gas_mix_block: 0..max_voices voices");
		}
		if (frames < 2 || (frames & 1) || frames > ctx->cfg.max_frames) {
			return gas_fail(ctx, GAS_ERR_INVALID, "gas_mix_block: frames must be even and in [2, max_frames]");
		}
		if (src_rows < 0 || src_rows > ctx->cfg.max_voices || (src_rows > 0 && !src)) {
			return gas_fail(ctx, GAS_ERR_INVALID, "gas_mix_block: 0..max_voices source rows");
		}
		for (int i = 0; i < n_voices; i++) {
			const gas_voice &v = voices[i];
			if (v.voice < 0 || v.voice >= ctx->cfg.max_voices || v.instance < 0 || v.instance >= ctx->cfg.max_instances || v.src_row >= src_rows) {
				return gas_fail(ctx, GAS_ERR_INVALID, "gas_mix_block: voice %d references a bad voice/instance slot or source row", i);
			}
		}
		if (n_voices > 0) {
			GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_voices, voices, n_voices * sizeof(gas_voice), cudaMemcpyHostToDevice, ctx->s_mix));
		}
		if (src_rows > 0) {
			GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_src, src, (size_t)src_rows * frames * sizeof(gas_frame), cudaMemcpyHostToDevice, ctx->s_mix));
		}
		int st = mix_core(ctx, n_voices, ctx->d_voices, ctx->d_src, src_rows, frames, frames, ctx->d_bus, peaks ? ctx->d_peaks : nullptr);
		if (st) {
			return st;
		}
		if (bus_out) {
			GAS_CUDA(ctx, cudaMemcpyAsync(bus_out, ctx->d_bus, (size_t)ctx->cfg.num_buses * (ctx->cfg.speaker_mode + 1) * frames * sizeof(gas_frame),
					cudaMemcpyDeviceToHost, ctx->s_mix));
		}
		if (peaks && n_voices > 0) {
			GAS_CUDA(ctx, cudaMemcpyAsync(peaks, ctx->d_peaks, n_voices * sizeof(gas_frame), cudaMemcpyDeviceToHost, ctx->s_mix));
		}
	}
	GAS_CUDA(ctx, cudaStreamSynchronize(ctx->s_mix));
	return GAS_OK;
}

int gas_mix_block_device(gas_ctx *ctx, int32_t n_voices, const gas_voice *d_voices, const gas_frame *d_src, int32_t src_rows,
		int32_t src_row_stride, int32_t frames, gas_frame *d_bus_out, gas_frame *d_peaks) {
	ENTER(ctx);
	if (n_voices < 0 || n_voices > ctx->cfg.max_voices || (n_voices > 0 && (!d_voices || !d_src)) || !d_bus_out) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_mix_block_device: bad voice count or null pointer");
	}
	if (frames < 2 || (frames & 1) || frames > ctx->cfg.max_frames || src_row_stride < frames || (src_row_stride & 1)) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_mix_block_device: frames/stride must be even, frames <= max_frames, stride >= frames");
	}
	if (((uintptr_t)d_src & 15u) || ((uintptr_t)d_bus_out & 15u)) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_mix_block_device: source and bus buffers must be 16-byte aligned");
	}
	return mix_core(ctx, n_voices, d_voices, d_src, src_rows, src_row_stride, frames, d_bus_out, d_peaks);
}

// ---- pipelined form -----------------------------------------------------------------------------------------------
int gas_step_device(gas_ctx *ctx, const gas_frame *d_src, int32_t src_row_stride, const gas_step_next *next) {
	ENTER(ctx);
	StepNext nx{};
	if (next) {
		if (next->n_voices < 0 || next->n_voices > ctx->cfg.max_voices || (next->n_voices > 0 && !next->d_voices) || !next->d_bus_out) {
			return gas_fail(ctx, GAS_ERR_INVALID, "gas_step_device: next block: bad voice count or null pointer");
		}
		if (next->frames < 2 || (next->frames & 1) || next->frames > ctx->cfg.max_frames || ((uintptr_t)next->d_bus_out & 15u)) {
			return gas_fail(ctx, GAS_ERR_INVALID, "gas_step_device: next block: frames must be even and <= max_frames, bus buffers 16-byte aligned");
		}
		if (next->n_emitters < 0 || next->n_emitters > ctx->cfg.max_instances || (next->n_emitters > 0 && !next->d_emitters)) {
			return gas_fail(ctx, GAS_ERR_INVALID, "gas_step_device: next block: bad emitter count or null pointer");
		}
		if (next->n_emitters > 0 && ctx->n_listeners_res <= 0) {
			return gas_fail(ctx, GAS_ERR_STATE, "gas_step_device: gains need resident listeners (gas_listeners_set)");
		}
		if (ctx->planned.valid && (next->d_bus_out == ctx->planned.bus || (next->d_peaks && next->d_peaks == ctx->planned.peaks))) {
			return gas_fail(ctx, GAS_ERR_INVALID, "gas_step_device: the next block must not share output buffers with the block being streamed");
		}
		nx.n_emitters = next->n_emitters;
		nx.d_emitters = next->d_emitters;
		nx.n_voices = next->n_voices;
		nx.d_voices = next->d_voices;
		nx.src_rows = next->src_rows;
		nx.frames = next->frames;
		nx.d_bus = next->d_bus_out;
		nx.d_peaks = next->d_peaks;
	}
	if (ctx->planned.valid) {
		if ((ctx->planned.n_voices > 0 && !d_src) || src_row_stride < ctx->planned.frames || (src_row_stride & 1) || ((uintptr_t)d_src & 15u)) {
			return gas_fail(ctx, GAS_ERR_INVALID, "gas_step_device: source rows of the planned block: null / misaligned pointer or stride < frames");
		}
	}
	return step_core(ctx, d_src, src_row_stride, next ? &nx : nullptr);
}

int gas_step_join_device(gas_ctx *ctx) {
	ENTER(ctx);
	return join_voice_stream(ctx);
}

// ---- stream form: voice lifecycle around the block path (gas_life.cu) ---------------------------------------------
static int stream_core(gas_ctx *ctx, int n_voices, const gas_voice *d_voices, const gas_frame *d_src, int src_rows, int src_stride, int frames,
		const int32_t *d_mixed, gas_frame *d_bus, int32_t *d_status) {
	// stage: lookahead splice / end fade into the staging rows + rewritten voice list; then the ordinary block path with
	// peaks; then the deactivation pass
	GAS_CUDA(ctx, launch_life_stage(ctx, n_voices, d_voices, d_mixed, d_src, src_rows, src_stride, frames, ctx->d_voices_stage, ctx->d_stage, frames,
			ctx->s_mix));
	int st = mix_core(ctx, n_voices, ctx->d_voices_stage, ctx->d_stage, n_voices, frames, frames, d_bus, ctx->d_peaks);
	if (st) {
		return st;
	}
	GAS_CUDA(ctx, launch_life_post(ctx, n_voices, ctx->d_voices_stage, ctx->d_peaks, d_status, ctx->s_mix));
	GAS_CUDA(ctx, cudaEventRecord(ctx->ev_mix_done, ctx->s_mix));
	return GAS_OK;
}

int gas_mix_block_stream(gas_ctx *ctx, int32_t n_voices, const gas_voice *voices, const gas_frame *src, int32_t src_rows, int32_t frames,
		const int32_t *mixed_frames, gas_frame *bus_out, int32_t *status_out) {
	{
		ENTER(ctx);
		if (n_voices < 0 || n_voices > ctx->cfg.max_voices || (n_voices > 0 && (!voices || !mixed_frames))) {
			return gas_fail(ctx, GAS_ERR_INVALID, "gas_mix_block_stream: 0..max_voices voices, one frame count per voice");
		}
		if (frames < 2 || (frames & 1) || frames > ctx->cfg.max_frames) {
			return gas_fail(ctx, GAS_ERR_INVALID, "gas_mix_block_stream: frames must be even and in [2, max_frames]");
		}
		if (src_rows < 0 || src_rows > ctx->cfg.max_voices || (src_rows > 0 && !src)) {
			return gas_fail(ctx, GAS_ERR_INVALID, "gas_mix_block_stream: 0..max_voices source rows");
		}
		for (int i = 0; i < n_voices; i++) {
			const gas_voice &v = voices[i];
			if (v.voice < 0 || v.voice >= ctx->cfg.max_voices || v.instance < 0 || v.instance >= ctx->cfg.max_instances || v.src_row >= src_rows ||
					mixed_frames[i] < 0 || mixed_frames[i] > frames) {
				return gas_fail(ctx, GAS_ERR_INVALID, "gas_mix_block_stream: voice %d references a bad slot / row or has a frame count outside [0, frames]", i);
			}
		}
		if (n_voices > 0) {
			GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_voices, voices, n_voices * sizeof(gas_voice), cudaMemcpyHostToDevice, ctx->s_mix));
			GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_mixed, mixed_frames, n_voices * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->s_mix));
		}
		if (src_rows > 0) {
			GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_src, src, (size_t)src_rows * frames * sizeof(gas_frame), cudaMemcpyHostToDevice, ctx->s_mix));
		}
		int st = stream_core(ctx, n_voices, ctx->d_voices, ctx->d_src, src_rows, frames, frames, ctx->d_mixed, ctx->d_bus, ctx->d_status);
		if (st) {
			return st;
		}
		if (bus_out) {
			GAS_CUDA(ctx, cudaMemcpyAsync(bus_out, ctx->d_bus, (size_t)ctx->cfg.num_buses * (ctx->cfg.speaker_mode + 1) * frames * sizeof(gas_frame),
					cudaMemcpyDeviceToHost, ctx->s_mix));
		}
		if (status_out && n_voices > 0) {
			GAS_CUDA(ctx, cudaMemcpyAsync(status_out, ctx->d_status, n_voices * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->s_mix));
		}
	}
	GAS_CUDA(ctx, cudaStreamSynchronize(ctx->s_mix));
	return GAS_OK;
}

int gas_mix_block_stream_device(gas_ctx *ctx, int32_t n_voices, const gas_voice *d_voices, const gas_frame *d_src, int32_t src_rows,
		int32_t src_row_stride, int32_t frames, const int32_t *d_mixed_frames, gas_frame *d_bus_out, int32_t *d_status_out) {
	ENTER(ctx);
	if (n_voices < 0 || n_voices > ctx->cfg.max_voices || (n_voices > 0 && (!d_voices || !d_mixed_frames)) || !d_bus_out) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_mix_block_stream_device: bad voice count or null pointer");
	}
	if (frames < 2 || (frames & 1) || frames > ctx->cfg.max_frames || src_row_stride < frames || (src_row_stride & 1)) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_mix_block_stream_device: frames/stride must be even, frames <= max_frames, stride >= frames");
	}
	if (((uintptr_t)d_src & 7u) || ((uintptr_t)d_bus_out & 15u)) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_mix_block_stream_device: source rows must be 8-byte, bus buffers 16-byte aligned");
	}
	return stream_core(ctx, n_voices, d_voices, d_src, src_rows, src_row_stride, frames, d_mixed_frames, d_bus_out, d_status_out);
}

// ---- bus graph after the mix --------------------------------------------------------------------
int gas_bus_layout_set(gas_ctx *ctx, int32_t n_buses, const gas_bus_desc *buses) {
	ENTER(ctx);
	if (n_buses != ctx->cfg.num_buses || !buses || ctx->capturing) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_bus_layout_set: one descriptor per bus of the context (%d), not while capturing", ctx->cfg.num_buses);
	}
	// upstream AudioServer::_mix_step: with any bus soloed, only soloed buses and the buses on their send chains are audible
	bool solo_mode = false, soloed[GAS_MAX_BUSES] = {};
	int send[GAS_MAX_BUSES] = {};
	for (int b = 0; b < n_buses; b++) {
		int t = buses[b].send;
		send[b] = (b > 0 && t >= 0 && t < b) ? t : 0; // an invalid send (unknown bus, or one that is not to the left) goes to Master
		solo_mode = solo_mode || buses[b].solo != 0;
	}
	if (solo_mode) {
		for (int b = 0; b < n_buses; b++) {
			if (buses[b].solo) {
				int i = b;
				soloed[i] = true;
				while (i != 0) {
					i = send[i];
					soloed[i] = true;
				}
			}
		}
	}
	for (int b = 0; b < n_buses; b++) {
		float v = expf(buses[b].volume_db * (float)0.11512925464970228420089957273422); // Math::db_to_linear(float)
		if (solo_mode ? !soloed[b] : buses[b].mute != 0) {
			v = 0.0f;
		}
		ctx->bus_volume_lin[b] = v;
		ctx->bus_send[b] = send[b];
	}
	return GAS_OK;
}

int gas_bus_graph_device(gas_ctx *ctx, gas_frame *d_bus, int32_t frames) {
	ENTER(ctx);
	if (!d_bus || frames < 2 || (frames & 1) || frames > ctx->cfg.max_frames || ((uintptr_t)d_bus & 15u)) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_bus_graph_device: bad buffer or frame count");
	}
	int st = join_voice_stream(ctx); // voice-parallel kernels of pipelined steps may still be adding to the buffers
	if (st) {
		return st;
	}
	if (ctx->comm_pending) {
		GAS_CUDA(ctx, cudaStreamWaitEvent(ctx->s_mix, ctx->ev_comm_done, 0));
		ctx->comm_pending = false;
	}
	GAS_CUDA(ctx, launch_bus_graph(ctx, d_bus, frames, ctx->s_mix));
	GAS_CUDA(ctx, cudaEventRecord(ctx->ev_mix_done, ctx->s_mix));
	ctx->mix_pending = true;
	return GAS_OK;
}

int gas_bus_graph(gas_ctx *ctx, gas_frame *bus_inout, int32_t frames) {
	{
		ENTER(ctx);
		if (!bus_inout || frames < 2 || (frames & 1) || frames > ctx->cfg.max_frames || ctx->capturing) {
			return gas_fail(ctx, GAS_ERR_INVALID, "gas_bus_graph: bad buffer or frame count (or capture in progress)");
		}
		const size_t bytes = (size_t)ctx->cfg.num_buses * (ctx->cfg.speaker_mode + 1) * frames * sizeof(gas_frame);
		int st = join_voice_stream(ctx);
		if (st) {
			return st;
		}
		GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_bus, bus_inout, bytes, cudaMemcpyHostToDevice, ctx->s_mix));
		GAS_CUDA(ctx, launch_bus_graph(ctx, ctx->d_bus, frames, ctx->s_mix));
		GAS_CUDA(ctx, cudaMemcpyAsync(bus_inout, ctx->d_bus, bytes, cudaMemcpyDeviceToHost, ctx->s_mix));
	}
	GAS_CUDA(ctx, cudaStreamSynchronize(ctx->s_mix));
	return GAS_OK;
}

// ---- device-resident sources + resampler -----------------------------------------------------------
int gas_source_set(gas_ctx *ctx, int32_t slot, const gas_frame *frames, int32_t n_frames, float sample_rate, int32_t loop) {
	ENTER(ctx);
	if (slot < 0 || slot >= ctx->max_sources || n_frames < 0 || (n_frames > 0 && !frames) || !(sample_rate > 0.f) || ctx->capturing) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_source_set: slot 0..%d, frames, sample_rate > 0, not while capturing", ctx->max_sources - 1);
	}
	if (n_frames > 0 && n_frames < 128) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_source_set: a source holds at least 128 frames (one internal buffer of the resampler)");
	}
	GAS_CUDA(ctx, cudaStreamSynchronize(ctx->s_mix)); // nothing in flight reads the slot
	SourceDesc &sd = ctx->h_sources[slot];
	if (sd.pcm) {
		cudaFree((void *)sd.pcm);
		sd = SourceDesc{};
	}
	if (n_frames > 0) {
		gas_frame *d = nullptr;
		if (cudaMalloc((void **)&d, (size_t)n_frames * sizeof(gas_frame)) != cudaSuccess) {
			return gas_fail(ctx, GAS_ERR_NOMEM, "gas_source_set: out of device memory");
		}
		GAS_CUDA(ctx, cudaMemcpyAsync(d, frames, (size_t)n_frames * sizeof(gas_frame), cudaMemcpyHostToDevice, ctx->s_mix));
		sd.pcm = d;
		sd.n_frames = n_frames;
		sd.loop = loop ? 1 : 0;
		sd.sample_rate = sample_rate;
	}
	GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_sources + slot, &sd, sizeof(SourceDesc), cudaMemcpyHostToDevice, ctx->s_mix));
	GAS_CUDA(ctx, cudaStreamSynchronize(ctx->s_mix));
	return GAS_OK;
}

int gas_voice_play(gas_ctx *ctx, int32_t n, const int32_t *voices, const int32_t *sources, const int32_t *start_frames) {
	ENTER(ctx);
	if (n < 0 || n > ctx->cfg.max_voices || (n > 0 && (!voices || !sources)) || !ids_valid(voices, n, ctx->cfg.max_voices) || ctx->capturing) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_voice_play: bad voice slot (or capture in progress)");
	}
	for (int i = 0; i < n; i++) {
		const int s = sources[i];
		const int st = start_frames ? start_frames[i] : 0;
		if (s < -1 || s >= ctx->max_sources || (s >= 0 && !ctx->h_sources[s].pcm) || st < 0) {
			return gas_fail(ctx, GAS_ERR_INVALID, "gas_voice_play: voice %d: source slot %d is empty or out of range", voices[i], s);
		}
		if (s >= 0 && !ctx->h_sources[s].loop && ctx->h_sources[s].n_frames - st < 128) {
			return gas_fail(ctx, GAS_ERR_INVALID, "gas_voice_play: voice %d: fewer than 128 frames behind the start frame", voices[i]);
		}
	}
	if (n == 0) {
		return GAS_OK;
	}
	int32_t *d = ctx->d_ids_mix; // [max(V, I)] ints: voices | sources | starts need 3 n <= 3 V: staged through the scratch buffer instead
	(void)d;
	int32_t *dv = (int32_t *)ctx->d_scratch_mix, *ds = dv + n, *dst = ds + n;
	GAS_CUDA(ctx, cudaMemcpyAsync(dv, voices, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->s_mix));
	GAS_CUDA(ctx, cudaMemcpyAsync(ds, sources, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->s_mix));
	if (start_frames) {
		GAS_CUDA(ctx, cudaMemcpyAsync(dst, start_frames, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->s_mix));
	}
	GAS_CUDA(ctx, launch_voice_play(ctx, n, dv, ds, start_frames ? dst : nullptr, ctx->s_mix));
	GAS_CUDA(ctx, cudaStreamSynchronize(ctx->s_mix)); // the host arrays may be reused
	return GAS_OK;
}

int gas_resample_block_device(gas_ctx *ctx, int32_t n_voices, const gas_voice *d_voices, int32_t frames, gas_frame *d_rows, int32_t row_stride,
		int32_t src_rows, int32_t *d_mixed_frames) {
	ENTER(ctx);
	if (n_voices < 0 || n_voices > ctx->cfg.max_voices || (n_voices > 0 && (!d_voices || !d_rows)) || frames < 1 || frames > ctx->cfg.max_frames ||
			row_stride < frames || src_rows < 0) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_resample_block_device: bad voice count, frame count, stride or null pointer");
	}
	if (ctx->gain_pending) { // the pitch scale comes from the instance's current parameters
		GAS_CUDA(ctx, cudaStreamWaitEvent(ctx->s_mix, ctx->ev_gain_done, 0));
	}
	GAS_CUDA(ctx, launch_resample(ctx, n_voices, d_voices, frames, d_rows, row_stride, src_rows, d_mixed_frames, ctx->s_mix));
	GAS_CUDA(ctx, cudaEventRecord(ctx->ev_prologue_done, ctx->s_mix)); // later gain-side calls wait: they rewrite the pitch scale
	ctx->prologue_pending = true;
	return GAS_OK;
}

int gas_mix_block_resident(gas_ctx *ctx, int32_t n_voices, const gas_voice *voices, int32_t frames, gas_frame *bus_out, int32_t *status_out) {
	{
		ENTER(ctx);
		if (n_voices < 0 || n_voices > ctx->cfg.max_voices || (n_voices > 0 && !voices) || !bus_out) {
			return gas_fail(ctx, GAS_ERR_INVALID, "gas_mix_block_resident: bad voice count or null pointer");
		}
		if (frames < GAS_LOOKAHEAD_BUFFER_SIZE || (frames & 1) || frames > ctx->cfg.max_frames) {
			return gas_fail(ctx, GAS_ERR_INVALID, "gas_mix_block_resident: frames must be even, >= %d and <= max_frames", GAS_LOOKAHEAD_BUFFER_SIZE);
		}
		if (ctx->capturing) {
			return gas_fail(ctx, GAS_ERR_STATE, "gas_mix_block_resident: host pointers cannot be used while capturing");
		}
		if (ctx->planned.valid) {
			return gas_fail(ctx, GAS_ERR_STATE, "gas_mix_block_resident: a block planned by gas_step_device is waiting to be streamed");
		}
		for (int i = 0; i < n_voices; i++) {
			const gas_voice &v = voices[i];
			if (v.voice < 0 || v.voice >= ctx->cfg.max_voices || v.instance < 0 || v.instance >= ctx->cfg.max_instances || v.src_row < -1 ||
					v.src_row >= ctx->cfg.max_voices) {
				return gas_fail(ctx, GAS_ERR_INVALID, "gas_mix_block_resident: voice %d references a bad slot or row", i);
			}
		}
		const size_t bus_frames = (size_t)ctx->cfg.num_buses * (ctx->cfg.speaker_mode + 1) * frames;
		if (n_voices > 0) {
			GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_voices, voices, n_voices * sizeof(gas_voice), cudaMemcpyHostToDevice, ctx->s_mix));
		}
		if (ctx->gain_pending) {
			GAS_CUDA(ctx, cudaStreamWaitEvent(ctx->s_mix, ctx->ev_gain_done, 0));
		}
		// source rows come from the resident PCM: row src_row of the internal row buffer for every voice
		GAS_CUDA(ctx, launch_resample(ctx, n_voices, ctx->d_voices, frames, ctx->d_rs_rows, frames, ctx->cfg.max_voices, ctx->d_rs_mixed, ctx->s_mix));
		int st = stream_core(ctx, n_voices, ctx->d_voices, ctx->d_rs_rows, ctx->cfg.max_voices, frames, frames, ctx->d_rs_mixed, ctx->d_bus, ctx->d_status);
		if (st) {
			return st;
		}
		GAS_CUDA(ctx, cudaMemcpyAsync(bus_out, ctx->d_bus, bus_frames * sizeof(gas_frame), cudaMemcpyDeviceToHost, ctx->s_mix));
		if (status_out && n_voices > 0) {
			GAS_CUDA(ctx, cudaMemcpyAsync(status_out, ctx->d_status, n_voices * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->s_mix));
		}
	}
	GAS_CUDA(ctx, cudaStreamSynchronize(ctx->s_mix));
	return GAS_OK;
}

int gas_set_playback_disable_threshold_db(gas_ctx *ctx, int32_t n, const int32_t *instances, const float *db) {
	ENTER(ctx);
	if (n < 0 || n > ctx->cfg.max_instances || (n > 0 && (!instances || !db)) || !ids_valid(instances, n, ctx->cfg.max_instances)) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_set_playback_disable_threshold_db: bad instance slot");
	}
	std::vector<float> lin((size_t)(n > 0 ? n : 1));
	for (int i = 0; i < n; i++) {
		lin[(size_t)i] = expf(db[i] * (float)0.11512925464970228420089957273422); // upstream Math::db_to_linear(float), on the host like the reference
	}
	// the mix stream owns the lifecycle tables and has staging buffers of its own
	GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_ids_mix, instances, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->s_mix));
	GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_scratch_mix, lin.data(), n * sizeof(float), cudaMemcpyHostToDevice, ctx->s_mix));
	GAS_CUDA(ctx, launch_threshold_set(ctx, n, ctx->d_ids_mix, (const float *)ctx->d_scratch_mix, ctx->s_mix));
	GAS_CUDA(ctx, cudaStreamSynchronize(ctx->s_mix)); // `lin` is pageable and about to go out of scope
	return GAS_OK;
}

int gas_voice_life_export(gas_ctx *ctx, int32_t n, const int32_t *voices, gas_voice_life *out) {
	{
		ENTER(ctx);
		if (n < 0 || n > ctx->cfg.max_voices || (n > 0 && (!voices || !out)) || !ids_valid(voices, n, ctx->cfg.max_voices)) {
			return gas_fail(ctx, GAS_ERR_INVALID, "gas_voice_life_export: bad voice slot");
		}
		GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_ids_mix, voices, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->s_mix));
		GAS_CUDA(ctx, launch_life_export(ctx, n, ctx->d_ids_mix, (gas_voice_life *)ctx->d_scratch_mix, ctx->s_mix));
		if (n > 0) {
			GAS_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_scratch_mix, n * sizeof(gas_voice_life), cudaMemcpyDeviceToHost, ctx->s_mix));
		}
	}
	GAS_CUDA(ctx, cudaStreamSynchronize(ctx->s_mix));
	return GAS_OK;
}

int gas_voice_life_import(gas_ctx *ctx, int32_t n, const int32_t *voices, const gas_voice_life *in) {
	ENTER(ctx);
	if (n < 0 || n > ctx->cfg.max_voices || (n > 0 && (!voices || !in)) || !ids_valid(voices, n, ctx->cfg.max_voices)) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_voice_life_import: bad voice slot");
	}
	GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_ids_mix, voices, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->s_mix));
	GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_scratch_mix, in, n * sizeof(gas_voice_life), cudaMemcpyHostToDevice, ctx->s_mix));
	GAS_CUDA(ctx, launch_life_import(ctx, n, ctx->d_ids_mix, (const gas_voice_life *)ctx->d_scratch_mix, ctx->s_mix));
	GAS_CUDA(ctx, cudaStreamSynchronize(ctx->s_mix));
	return GAS_OK;
}

int gas_status_flags(gas_ctx *ctx, uint32_t *out_flags) {
	{
		ENTER(ctx);
		if (!out_flags) {
			return gas_fail(ctx, GAS_ERR_INVALID, "gas_status_flags: null output");
		}
	}
	GAS_CUDA(ctx, cudaStreamSynchronize(ctx->s_mix));
	int32_t overflow = 0;
	GAS_CUDA(ctx, cudaMemcpy(&overflow, ctx->plan.overflow, sizeof(overflow), cudaMemcpyDeviceToHost));
	if (overflow) {
		GAS_CUDA(ctx, cudaMemset(ctx->plan.overflow, 0, sizeof(overflow)));
	}
	*out_flags = overflow ? GAS_STATUS_CLASS_OVERFLOW : 0u;
	return GAS_OK;
}

// ---- the per-call virtuals on one voice (gas_single.cu) ---------------------------------------------------------------
static int single_call(gas_ctx *ctx, int32_t instance, int32_t voice, int32_t channel, gas_frame *out, const gas_frame *src, int32_t frames, const char *who) {
	if (instance < 0 || instance >= ctx->cfg.max_instances || voice < 0 || voice >= ctx->cfg.max_voices || !out || !src) {
		return gas_fail(ctx, GAS_ERR_INVALID, "%s: bad instance / voice slot or null buffer", who);
	}
	if (frames < 1 || frames > ctx->cfg.max_frames) {
		return gas_fail(ctx, GAS_ERR_INVALID, "%s: frames must be in [1, max_frames]", who);
	}
	// ordered after the gain side (the instance's parameters) and on the mix stream (the voice's playback data)
	if (ctx->gain_pending) {
		GAS_CUDA(ctx, cudaStreamWaitEvent(ctx->s_mix, ctx->ev_gain_done, 0));
	}
	gas_frame *d_in = ctx->d_stage, *d_out = ctx->d_stage + ctx->cfg.max_frames;
	GAS_CUDA(ctx, cudaMemcpyAsync(d_in, src, (size_t)frames * sizeof(gas_frame), cudaMemcpyHostToDevice, ctx->s_mix));
	GAS_CUDA(ctx, launch_single_voice(ctx, instance, voice, channel, d_out, d_in, frames, ctx->s_mix));
	GAS_CUDA(ctx, cudaMemcpyAsync(out, d_out, (size_t)frames * sizeof(gas_frame), cudaMemcpyDeviceToHost, ctx->s_mix));
	GAS_CUDA(ctx, cudaEventRecord(ctx->ev_prologue_done, ctx->s_mix)); // later gain-side work waits for this read of the parameters
	ctx->prologue_pending = true;
	return GAS_OK;
}

int gas_process_frames(gas_ctx *ctx, int32_t instance, int32_t voice, gas_frame *out, const gas_frame *src, int32_t frames) {
	{
		ENTER(ctx);
		int st = single_call(ctx, instance, voice, -1, out, src, frames, "gas_process_frames");
		if (st) {
			return st;
		}
	}
	GAS_CUDA(ctx, cudaStreamSynchronize(ctx->s_mix));
	return GAS_OK;
}

int gas_mix_channel(gas_ctx *ctx, int32_t instance, int32_t voice, int32_t channel, gas_frame *out, const gas_frame *src, int32_t frames) {
	{
		ENTER(ctx);
		if (channel < 0 || channel >= GAS_MAX_CHANNELS_PER_BUS) { // ERR_FAIL_INDEX of get_filter_processor, audio_spatializer_3d.cpp:888
			return gas_fail(ctx, GAS_ERR_INVALID, "gas_mix_channel: channel must be in [0, %d)", GAS_MAX_CHANNELS_PER_BUS);
		}
		int st = single_call(ctx, instance, voice, channel, out, src, frames, "gas_mix_channel");
		if (st) {
			return st;
		}
	}
	GAS_CUDA(ctx, cudaStreamSynchronize(ctx->s_mix));
	return GAS_OK;
}

int gas_sync(gas_ctx *ctx) {
	if (!ctx) {
		return gas_fail(nullptr, GAS_ERR_INVALID, "gas_sync: null context");
	}
	GAS_CUDA(ctx, cudaSetDevice(ctx->device));
	GAS_CUDA(ctx, cudaStreamSynchronize(ctx->s_gain));
	GAS_CUDA(ctx, cudaStreamSynchronize(ctx->s_comm));
	GAS_CUDA(ctx, cudaStreamSynchronize(ctx->s_mix));
	GAS_CUDA(ctx, cudaStreamSynchronize(ctx->s_voice));
	return GAS_OK;
}

void *gas_mix_stream(gas_ctx *ctx) { return ctx ? (void *)ctx->s_mix : nullptr; }
void *gas_gain_stream(gas_ctx *ctx) { return ctx ? (void *)ctx->s_gain : nullptr; }
uint64_t gas_kernel_launches(const gas_ctx *ctx) { return ctx ? ctx->launches : 0; }
// experiments only (not in gas.h): device pointer of the K2 timeline buffer, [CTA][8] uint64
extern "C" GAS_API void *gas_debug_timeline(gas_ctx *ctx) { return ctx ? (void *)ctx->d_timeline : nullptr; }
// experiments only (not in gas.h): the routing-class table after a synchronise — keys[128], counts[2][128] (the two most recent plan slots)
extern "C" GAS_API int gas_debug_classes(gas_ctx *ctx, unsigned long long *keys, int32_t *counts) {
	if (!ctx || !keys || !counts) {
		return GAS_ERR_INVALID;
	}
	cudaSetDevice(ctx->device);
	cudaDeviceSynchronize();
	cudaMemcpy(keys, ctx->plan.cls_key, GAS_MAX_CLASSES * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
	int32_t p = 0;
	cudaMemcpy(&p, ctx->t.blk + BLK_P, sizeof(int32_t), cudaMemcpyDeviceToHost);
	for (int i = 0; i < 2; i++) { // counts[0] = the last planned block, counts[1] = the one before
		const int slot = (p - 1 - i) & (GAS_PLAN_DEPTH - 1);
		cudaMemcpy(counts + i * GAS_MAX_CLASSES, ctx->plan.cls_count + slot * GAS_MAX_CLASSES, GAS_MAX_CLASSES * sizeof(int32_t), cudaMemcpyDeviceToHost);
	}
	return GAS_OK;
}

int gas_voice_state_export(gas_ctx *ctx, int32_t n, const int32_t *voices, gas_voice_state *out) {
	{
		ENTER(ctx);
		if (n < 0 || n > ctx->cfg.max_voices || (n > 0 && (!voices || !out)) || !ids_valid(voices, n, ctx->cfg.max_voices)) {
			return gas_fail(ctx, GAS_ERR_INVALID, "gas_voice_state_export: bad voice slot");
		}
		GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_ids_mix, voices, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->s_mix));
		GAS_CUDA(ctx, launch_state_export(ctx, n, ctx->d_ids_mix, (gas_voice_state *)ctx->d_scratch_mix, ctx->s_mix));
		if (n > 0) {
			GAS_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_scratch_mix, n * sizeof(gas_voice_state), cudaMemcpyDeviceToHost, ctx->s_mix));
		}
	}
	GAS_CUDA(ctx, cudaStreamSynchronize(ctx->s_mix));
	return GAS_OK;
}

int gas_voice_state_import(gas_ctx *ctx, int32_t n, const int32_t *voices, const gas_voice_state *in) {
	ENTER(ctx);
	if (n < 0 || n > ctx->cfg.max_voices || (n > 0 && (!voices || !in)) || !ids_valid(voices, n, ctx->cfg.max_voices)) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_voice_state_import: bad voice slot");
	}
	GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_ids_mix, voices, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->s_mix));
	GAS_CUDA(ctx, cudaMemcpyAsync(ctx->d_scratch_mix, in, n * sizeof(gas_voice_state), cudaMemcpyHostToDevice, ctx->s_mix));
	GAS_CUDA(ctx, launch_state_import(ctx, n, ctx->d_ids_mix, (const gas_voice_state *)ctx->d_scratch_mix, ctx->s_mix));
	return GAS_OK;
}

// ---- CUDA-graph capture ---------------------------------------------------------------------------------
int gas_capture_begin(gas_ctx *ctx) {
	ENTER(ctx);
	if (ctx->capturing) {
		return gas_fail(ctx, GAS_ERR_STATE, "gas_capture_begin: already capturing");
	}
	GAS_CUDA(ctx, cudaStreamSynchronize(ctx->s_gain));
	GAS_CUDA(ctx, cudaStreamSynchronize(ctx->s_comm));
	GAS_CUDA(ctx, cudaStreamSynchronize(ctx->s_mix));
	GAS_CUDA(ctx, cudaStreamSynchronize(ctx->s_voice));
	ctx->gain_pending = ctx->prologue_pending = ctx->mix_pending = ctx->comm_pending = false;
	ctx->stream_started_pending = false;
	ctx->step_done_valid = false;
	for (int i = 0; i < GAS_PLAN_DEPTH; i++) {
		ctx->block_inflight[i] = false;
		ctx->comm_ring_valid[i] = false;
	}
	GAS_CUDA(ctx, cudaStreamBeginCapture(ctx->s_mix, cudaStreamCaptureModeThreadLocal));
	// fork: the gain, exchange and voice streams join the capture
	GAS_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->s_mix));
	GAS_CUDA(ctx, cudaStreamWaitEvent(ctx->s_gain, ctx->ev_fork, 0));
	GAS_CUDA(ctx, cudaStreamWaitEvent(ctx->s_comm, ctx->ev_fork, 0));
	GAS_CUDA(ctx, cudaStreamWaitEvent(ctx->s_voice, ctx->ev_fork, 0));
	ctx->capturing = true;
	ctx->capture_profiled = false;
	for (int k = 0; k < GAS_KERNEL_KINDS; k++) {
		ctx->gev_used[k] = false;
	}
	ctx->capture_launches0 = ctx->launches;
	return GAS_OK;
}

int gas_capture_end(gas_ctx *ctx, int32_t *out_graph) {
	ENTER(ctx);
	if (!ctx->capturing || !out_graph) {
		return gas_fail(ctx, GAS_ERR_STATE, "gas_capture_end: not capturing (or null out_graph)");
	}
	ctx->capturing = false;
	cudaGraph_t graph = nullptr;
	cudaError_t e = cudaEventRecord(ctx->ev_join, ctx->s_gain); // join the gain stream back
	if (e == cudaSuccess) {
		e = cudaStreamWaitEvent(ctx->s_mix, ctx->ev_join, 0);
	}
	if (e == cudaSuccess) {
		e = cudaEventRecord(ctx->ev_join2, ctx->s_comm); // and the exchange stream
	}
	if (e == cudaSuccess) {
		e = cudaStreamWaitEvent(ctx->s_mix, ctx->ev_join2, 0);
	}
	if (e == cudaSuccess) {
		e = cudaEventRecord(ctx->ev_voice_join, ctx->s_voice); // and the voice stream
	}
	if (e == cudaSuccess) {
		e = cudaStreamWaitEvent(ctx->s_mix, ctx->ev_voice_join, 0);
	}
	ctx->step_done_valid = false;
	for (int i = 0; i < GAS_PLAN_DEPTH; i++) {
		ctx->block_inflight[i] = false;
		ctx->comm_ring_valid[i] = false;
	}
	cudaError_t e2 = cudaStreamEndCapture(ctx->s_mix, &graph);
	const uint64_t kernels = ctx->launches - ctx->capture_launches0;
	ctx->launches = ctx->capture_launches0; // nothing ran yet
	ctx->gain_pending = ctx->prologue_pending = ctx->mix_pending = ctx->comm_pending = false; // events recorded while capturing are graph edges only
	ctx->stream_started_pending = false;
	if (e != cudaSuccess || e2 != cudaSuccess || !graph) {
		return gas_fail(ctx, GAS_ERR_CUDA, "gas_capture_end: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
	}
	gas_ctx::Graph g;
	e = cudaGraphInstantiate(&g.exec, graph, 0);
	cudaGraphDestroy(graph);
	if (e != cudaSuccess) {
		return gas_fail(ctx, GAS_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
	}
	g.kernels = kernels;
	g.profiled = ctx->capture_profiled;
	g.cfg_epoch = ctx->cfg_epoch;
	ctx->graphs.push_back(g);
	*out_graph = (int32_t)ctx->graphs.size() - 1;
	return GAS_OK;
}

int gas_graph_launch(gas_ctx *ctx, int32_t graph) {
	ENTER(ctx);
	if (ctx->capturing || graph < 0 || graph >= (int32_t)ctx->graphs.size() || !ctx->graphs[graph].exec) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_graph_launch: bad graph id or capture in progress");
	}
	if (ctx->graphs[graph].cfg_epoch != ctx->cfg_epoch) {
		return gas_fail(ctx, GAS_ERR_STATE, "gas_graph_launch: the graph was captured before gas_set_speaker_mode / gas_set_mix_rate / "
				"gas_set_global_panning_strength changed the globals it has baked in: capture it again");
	}
	if (ctx->gain_pending) {
		GAS_CUDA(ctx, cudaStreamWaitEvent(ctx->s_mix, ctx->ev_gain_done, 0));
	}
	GAS_CUDA(ctx, cudaGraphLaunch(ctx->graphs[graph].exec, ctx->s_mix));
	ctx->launches += ctx->graphs[graph].kernels;
	if (ctx->graphs[graph].profiled && ctx->profiling) { // a profiled graph is read back after every launch
		GAS_CUDA(ctx, cudaStreamSynchronize(ctx->s_mix));
		for (int k = 0; k < GAS_KERNEL_KINDS; k++) {
			if (ctx->gev[k][0]) {
				float ms = 0.f;
				if (cudaEventElapsedTime(&ms, ctx->gev[k][0], ctx->gev[k][1]) == cudaSuccess) {
					ctx->prof_ms[k] += ms;
					ctx->prof_n[k] += 1;
				}
			}
		}
	}
	GAS_CUDA(ctx, cudaEventRecord(ctx->ev_prologue_done, ctx->s_mix)); // later gain-side work waits for the whole graph
	ctx->prologue_pending = true;
	GAS_CUDA(ctx, cudaEventRecord(ctx->ev_mix_done, ctx->s_mix)); // ... and so does a later exchange
	ctx->mix_pending = true;
	// ... and the voice-parallel kernel of a pipelined step enqueued after this launch (its plan may come out of this graph)
	GAS_CUDA(ctx, cudaEventRecord(ctx->ev_step_done[(ctx->step_count + GAS_PLAN_DEPTH - 1) % GAS_PLAN_DEPTH], ctx->s_mix));
	ctx->step_done_valid = true;
	return GAS_OK;
}

int gas_graph_destroy(gas_ctx *ctx, int32_t graph) {
	ENTER(ctx);
	if (graph < 0 || graph >= (int32_t)ctx->graphs.size()) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_graph_destroy: bad graph id");
	}
	if (ctx->graphs[graph].exec) {
		GAS_CUDA(ctx, cudaStreamSynchronize(ctx->s_mix));
		cudaGraphExecDestroy(ctx->graphs[graph].exec);
		ctx->graphs[graph].exec = nullptr;
	}
	return GAS_OK;
}

// ---- per-kernel timing ------------------------------------------------------------------------------------
int gas_profile_enable(gas_ctx *ctx, int32_t on) {
	ENTER(ctx);
	int st = prof_drain(ctx);
	if (st) {
		return st;
	}
	ctx->profiling = on != 0;
	if (on) {
		for (int k = 0; k < GAS_KERNEL_KINDS; k++) {
			ctx->prof_ms[k] = 0.0;
			ctx->prof_n[k] = 0;
		}
	}
	return GAS_OK;
}

int gas_profile_read(gas_ctx *ctx, double ms_out[GAS_KERNEL_KINDS], uint64_t launches_out[GAS_KERNEL_KINDS]) {
	ENTER(ctx);
	if (!ms_out || !launches_out) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_profile_read: null output");
	}
	int st = prof_drain(ctx);
	if (st) {
		return st;
	}
	for (int k = 0; k < GAS_KERNEL_KINDS; k++) {
		ms_out[k] = ctx->prof_ms[k];
		launches_out[k] = ctx->prof_n[k];
	}
	return GAS_OK;
}

// ---- multi-GPU exchange: peer-memory sum of the partial bus buffers (gas_comm.cu) --------------------------------
static size_t comm_alloc_bytes(gas_ctx *ctx) {
	return ((size_t)2 * ctx->comm_stride_f4 + 16) * sizeof(float4); // two parity buffers + the arrival counter
}

int gas_comm_export(gas_ctx *ctx, void *handle_out, size_t handle_bytes) {
	ENTER(ctx);
	if (!handle_out || handle_bytes < sizeof(cudaIpcMemHandle_t)) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_comm_export: handle buffer must hold %zu bytes", sizeof(cudaIpcMemHandle_t));
	}
	if (!ctx->d_exchange) {
		ctx->comm_stride_f4 = ctx->cfg.num_buses * GAS_MAX_CHANNELS_PER_BUS * ctx->cfg.max_frames / 2;
		GAS_CUDA(ctx, cudaMalloc((void **)&ctx->d_exchange, comm_alloc_bytes(ctx)));
		GAS_CUDA(ctx, cudaMemset(ctx->d_exchange, 0, comm_alloc_bytes(ctx)));
		GAS_CUDA(ctx, cudaMalloc((void **)&ctx->d_comm_seq, 2 * sizeof(unsigned long long)));
		GAS_CUDA(ctx, cudaMemset(ctx->d_comm_seq, 0, 2 * sizeof(unsigned long long)));
		GAS_CUDA(ctx, cudaMalloc((void **)&ctx->d_comm_ticket, 2 * sizeof(int)));
		GAS_CUDA(ctx, cudaMemset(ctx->d_comm_ticket, 0, 2 * sizeof(int)));
	}
	cudaIpcMemHandle_t h;
	GAS_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->d_exchange));
	memcpy(handle_out, &h, sizeof(h));
	return GAS_OK;
}

int gas_comm_open(gas_ctx *ctx, int32_t rank, int32_t n_ranks, const void *handles, size_t handle_bytes) {
	ENTER(ctx);
	if (n_ranks < 1 || n_ranks > 8 || rank < 0 || rank >= n_ranks || !handles || handle_bytes < sizeof(cudaIpcMemHandle_t)) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_comm_open: 1..8 ranks, one %zu-byte handle per rank", sizeof(cudaIpcMemHandle_t));
	}
	if (!ctx->d_exchange) {
		return gas_fail(ctx, GAS_ERR_STATE, "gas_comm_open: call gas_comm_export first");
	}
	for (int r = 0; r < n_ranks; r++) {
		if (r == rank) {
			ctx->peer_exchange[r] = ctx->d_exchange;
			continue;
		}
		cudaIpcMemHandle_t h;
		memcpy(&h, (const unsigned char *)handles + (size_t)r * handle_bytes, sizeof(h));
		void *p = nullptr;
		GAS_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
		ctx->peer_exchange[r] = (gas_frame *)p;
	}
	ctx->comm_rank = rank;
	ctx->comm_ranks = n_ranks;
	return GAS_OK;
}

static int reduce_half(gas_ctx *ctx, gas_frame *d_bus, int32_t frames, bool begin, bool end, const char *who) {
	if (!d_bus || frames < 2 || (frames & 1) || frames > ctx->cfg.max_frames || ((uintptr_t)d_bus & 15u)) {
		return gas_fail(ctx, GAS_ERR_INVALID, "%s: bad buffer or frame count", who);
	}
	if (ctx->comm_ranks <= 1) {
		return GAS_OK; // one rank: the partial sum is the sum
	}
	if (!ctx->d_exchange || !ctx->peer_exchange[ctx->comm_ranks - 1]) {
		return gas_fail(ctx, GAS_ERR_STATE, "%s: gas_comm_open has not been called", who);
	}
	if (ctx->comm_pending) {
		GAS_CUDA(ctx, cudaStreamWaitEvent(ctx->s_mix, ctx->ev_comm_done, 0));
		ctx->comm_pending = false;
	}
	{
		int st = join_voice_stream(ctx); // voice-parallel kernels of pipelined steps may still be adding to the buffer
		if (st) {
			return st;
		}
	}
	if (begin && !end && ctx->reduce_open) {
		return gas_fail(ctx, GAS_ERR_STATE, "%s: the previous block's gas_reduce_bus_end_device has not been called", who);
	}
	if (begin) {
		GAS_CUDA(ctx, launch_comm_push(ctx, d_bus, frames, ctx->s_mix));
		ctx->reduce_open = !end;
	}
	if (end) {
		GAS_CUDA(ctx, launch_comm_finish(ctx, d_bus, frames, ctx->s_mix));
		ctx->reduce_open = false;
	}
	return GAS_OK;
}

int gas_reduce_bus_device(gas_ctx *ctx, gas_frame *d_bus, int32_t frames) {
	ENTER(ctx);
	return reduce_half(ctx, d_bus, frames, true, true, "gas_reduce_bus_device");
}
int gas_reduce_bus_begin_device(gas_ctx *ctx, const gas_frame *d_bus, int32_t frames) {
	ENTER(ctx);
	return reduce_half(ctx, const_cast<gas_frame *>(d_bus), frames, true, false, "gas_reduce_bus_begin_device");
}
int gas_reduce_bus_end_device(gas_ctx *ctx, gas_frame *d_bus, int32_t frames) {
	ENTER(ctx);
	return reduce_half(ctx, d_bus, frames, false, true, "gas_reduce_bus_end_device");
}

int gas_reduce_bus_exchange_device(gas_ctx *ctx, const gas_frame *d_partial, gas_frame *d_prev_sum, int32_t frames) {
	ENTER(ctx);
	if (!d_partial || frames < 2 || (frames & 1) || frames > ctx->cfg.max_frames || ((uintptr_t)d_partial & 15u) || ((uintptr_t)d_prev_sum & 15u)) {
		return gas_fail(ctx, GAS_ERR_INVALID, "gas_reduce_bus_exchange_device: bad buffer or frame count");
	}
	if (ctx->comm_ranks <= 1) {
		return gas_fail(ctx, GAS_ERR_STATE, "gas_reduce_bus_exchange_device: needs an opened exchange of at least 2 ranks");
	}
	bool waited = false;
	for (int i = 0; i < GAS_PLAN_DEPTH; i++) { // a block of a pipelined run: complete when its step kernel and its voice-parallel kernel are
		if (ctx->block_inflight[i] && ctx->inflight_bus[i] == d_partial) {
			GAS_CUDA(ctx, cudaStreamWaitEvent(ctx->s_comm, ctx->ev_block_done[i], 0));
			waited = true;
		}
	}
	if (!waited && ctx->mix_pending) { // the partial sums must be complete (also when they were mixed earlier in the same capture)
		GAS_CUDA(ctx, cudaStreamWaitEvent(ctx->s_comm, ctx->ev_mix_done, 0));
	}
	GAS_CUDA(ctx, launch_comm_exchange(ctx, d_partial, d_prev_sum, frames, ctx->s_comm));
	GAS_CUDA(ctx, cudaEventRecord(ctx->ev_comm_done, ctx->s_comm));
	ctx->comm_pending = true;
	{
		const int r = (int)(ctx->comm_count++ % GAS_PLAN_DEPTH);
		if (ctx->comm_ring_valid[r]) { // four exchanges old: long done, but never dropped without an edge
			GAS_CUDA(ctx, cudaStreamWaitEvent(ctx->s_mix, ctx->ev_comm_ring[r], 0));
		}
		GAS_CUDA(ctx, cudaEventRecord(ctx->ev_comm_ring[r], ctx->s_comm));
		ctx->comm_src[r] = d_partial;
		ctx->comm_ring_valid[r] = true;
	}
	return GAS_OK;
}

int gas_comm_close(gas_ctx *ctx) {
	if (!ctx) {
		return GAS_OK;
	}
	for (int r = 0; r < 8; r++) {
		if (ctx->peer_exchange[r] && ctx->peer_exchange[r] != ctx->d_exchange) {
			cudaIpcCloseMemHandle(ctx->peer_exchange[r]);
		}
		ctx->peer_exchange[r] = nullptr;
	}
	ctx->comm_ranks = 1;
	ctx->comm_rank = 0;
	return GAS_OK;
}

} // extern "C"
