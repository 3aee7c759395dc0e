// gas_internal.h — context, device tables and per-block records shared by the .cu files.
// Product code: nothing here depends on the CPU checker that lives outside this package.
#pragma once

#include "../../include/gas.h"

#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>
#include <vector>

// ---- limits of the per-block plan -----------------------------------------------------------------
#define GAS_MAX_SENDS 12      // union of current and previous bus details: 6 + 6
#define GAS_MAX_CLASSES 16    // distinct (path, mode, send-mask) classes per block
#define GAS_K2_MAX_ROWS 6     // weight rows per (pair, side) the streaming kernel holds in registers
#define GAS_K2_ROW_FLOATS (GAS_K2_MAX_ROWS * GAS_MAX_CHANNELS_PER_BUS * 2)

// upstream AudioStreamPlaybackBusDetails of an instance's proxy playbacks, folded over the proxies:
// vol[k][c] is what reaches pair c of bus[k] (reference audio_spatializer.cpp:274-324).
struct BusDetails {
	int32_t n;
	int32_t bus[GAS_MAX_BUSES_PER_PLAYBACK];
	float vol[GAS_MAX_BUSES_PER_PLAYBACK][GAS_MAX_CHANNELS_PER_BUS][2];
};

// Sends of one instance for the current block, ascending by bus index: ramp N from vp to vn.
struct InstSends {
	int32_t n;
	uint32_t mask; // bit b set <=> a send to bus b exists
	int32_t bus[GAS_MAX_SENDS];
	float vp[GAS_MAX_SENDS][GAS_MAX_CHANNELS_PER_BUS][2];
	float vn[GAS_MAX_SENDS][GAS_MAX_CHANNELS_PER_BUS][2];
};

// Paths a voice can take through the block.
enum : int32_t {
	PATH_NONE = 0,   // nothing to stream (inactive instance)
	PATH_STREAM = 1, // K2: frame-parallel streaming contraction (no filter, no peak)
	PATH_VOICE = 2   // K3: voice-parallel serial-in-time kernel (filters, effect chains, peaks)
};
enum : int32_t { MODE_A = 0, MODE_B = 1, MODE_E = 2 };

// class flags
#define CLS_LIN 1u    // every weight is linear in t (2 rows per send instead of 3)
#define CLS_SHARED 2u // all sends carry identical weights: one row group fanned out to every bus of the mask
#define CLS_FILT 4u   // attenuation filter active (linear_attenuation >= 0.001)

struct ClassInfo {
	unsigned long long key; // 0 = empty
	int32_t count;          // voices appended so far
	int32_t path;
	int32_t mode;
	uint32_t flags;
	uint32_t mask;   // bus mask of the sends
	int32_t n_send;  // popcount(mask)
	int32_t n_group; // row groups: 1 if CLS_SHARED else n_send
	int32_t n_rows;  // n_group * (LIN ? 2 : 3)
};

// What K3 needs about one voice besides its persistent state.
struct VoiceRec {
	int32_t voice;
	int32_t instance;
	int32_t src_row;
	uint32_t flags;       // GAS_VOICE_* | (clear-history bits << 8, one per pair)
	float m_prev[GAS_MAX_CHANNELS_PER_BUS][2];
	float m_new[GAS_MAX_CHANNELS_PER_BUS][2];
	float target[5];      // high-shelf target coefficients (b0,b1,b2,a1,a2) when CLS_FILT
	int32_t n_fx;         // MODE_E: effects in the chain
	int32_t fx_stages[GAS_MAX_EFFECTS];
	float fx_coef[GAS_MAX_EFFECTS][5];
};

struct BlockPlan {
	ClassInfo *cls;      // [2][GAS_MAX_CLASSES]: by block parity; block n uses [n & 1], clears [(n + 1) & 1]
	int32_t *n_cls;      // [1]
	int32_t *overflow;   // [1] set when more than GAS_MAX_CLASSES classes were needed
	int32_t *k2_src;     // [GAS_MAX_CLASSES][max_voices] source row per list position
	float *k2_rows;      // [GAS_MAX_CLASSES][max_voices][GAS_K2_ROW_FLOATS] (compact: n_rows*C*2 used)
	int32_t *k3_list;    // [GAS_MAX_CLASSES][max_voices] call-order index j
	VoiceRec *rec;       // [max_voices] by call-order index
};

struct DevTables {
	gas_spatializer *spat;
	int32_t *inst_spat;
	gas_params *inst_params;
	int32_t *inst_was_further;
	int32_t *inst_active;
	BusDetails *inst_cur;
	BusDetails *inst_prev;   // [2][max_instances]: double-buffered by block parity (read [p], write [1-p])
	int32_t *inst_mode;      // MODE_A/B/E | (effect_gain_binding + 1) << 8, latched at instantiate()
	int32_t *blk;            // [0] block counter (parity of inst_prev), [1] CTA ticket of the prologue
	int32_t max_instances;
	gas_effect_chain *inst_fx;
	InstSends *inst_sends;
	float *vs_prev;              // [max_voices][4][2]
	gas_processor_state *vs_proc; // [max_voices][8]
	float *vs_fx;                // [max_voices][GAS_MAX_EFFECTS][2][GAS_MAX_FILTER_STAGES][4]
};

struct GlobalCfg {
	int32_t speaker_mode;
	int32_t channels;
	int32_t num_buses;
	float mix_rate;
	float global_panning;
	int32_t max_instances;
	int32_t max_voices;
	int32_t max_spatializers;
};

struct gas_ctx {
	gas_config cfg;
	GlobalCfg g;
	int device = 0;
	int num_sms = 0;
	int l2_bytes = 0;
	cudaStream_t s_mix = nullptr, s_gain = nullptr;
	cudaEvent_t ev_gain_done = nullptr, ev_prologue_done = nullptr, ev_fork = nullptr, ev_join = nullptr;
	bool gain_pending = false, prologue_pending = false;
	DevTables t{};
	BlockPlan plan{};
	// staging (device)
	gas_voice *d_voices = nullptr;
	gas_frame *d_src = nullptr;
	gas_frame *d_bus = nullptr;
	gas_frame *d_peaks = nullptr;
	gas_emitter *d_emitters = nullptr;
	gas_listener *d_listeners = nullptr;
	gas_area *d_areas = nullptr;
	int32_t max_areas = 0;
	gas_params *d_params_out = nullptr;
	int32_t *d_ids = nullptr;  // scratch id list [max(max_voices,max_instances)]
	int32_t *d_ids2 = nullptr;
	void *d_scratch = nullptr; // generic scratch for set/get/import/export payloads
	size_t scratch_bytes = 0;
	int32_t inst_hwm = 0; // instances [0, inst_hwm) have been initialised at least once
	// multi-GPU exchange
	gas_frame *d_exchange = nullptr;
	int32_t comm_rank = 0, comm_ranks = 1;
	gas_frame *peer_exchange[8] = {};
	uint64_t launches = 0;
	bool k2_smem_attr_set = false;
	int32_t n_listeners_res = 0, n_areas_res = 0; // resident listeners / areas (gas_listeners_set / gas_areas_set)
	// CUDA-graph capture
	bool capturing = false;
	uint64_t capture_launches0 = 0;
	struct Graph {
		cudaGraphExec_t exec = nullptr;
		uint64_t kernels = 0;
	};
	std::vector<Graph> graphs;
	// per-kernel timing
	bool profiling = false;
	struct ProfPair {
		cudaEvent_t a, b;
		int kind;
	};
	std::vector<ProfPair> prof_pairs;
	size_t prof_used = 0;
	double prof_ms[GAS_KERNEL_KINDS] = {};
	uint64_t prof_n[GAS_KERNEL_KINDS] = {};
	std::mutex mu;
	std::string err;
};

// ---- error helpers (gas_api.cu) ---------------------------------------------------------------------
int gas_fail(gas_ctx *ctx, int status, const char *fmt, ...);
#define GAS_CUDA(ctx, expr)                                                                    \
	do {                                                                                       \
		cudaError_t _e = (expr);                                                               \
		if (_e != cudaSuccess) {                                                               \
			return gas_fail((ctx), GAS_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e));     \
		}                                                                                      \
	} while (0)

// ---- kernel launchers (each returns cudaError_t from the launch) -----------------------------------
// gas_gain.cu
cudaError_t launch_gain(gas_ctx *ctx, int n, const gas_emitter *d_em, int n_listeners, const gas_listener *d_l,
		const gas_area *d_areas, gas_params *d_out, cudaStream_t st);
cudaError_t launch_params_set(gas_ctx *ctx, int n, const int32_t *d_ids, const gas_params *d_params, cudaStream_t st);
cudaError_t launch_instance_start(gas_ctx *ctx, int n, const int32_t *d_ids, cudaStream_t st);
// gas_prologue.cu
cudaError_t launch_prologue(gas_ctx *ctx, int n_voices, const gas_voice *d_voices, int src_rows, int frames, gas_frame *d_bus,
		gas_frame *d_peaks, cudaStream_t st);
// gas_mix_stream.cu (K2) / gas_mix_voice.cu (K3)
cudaError_t launch_mix_stream(gas_ctx *ctx, const gas_frame *d_src, int src_stride, int frames, gas_frame *d_bus, cudaStream_t st);
cudaError_t launch_mix_voice(gas_ctx *ctx, const gas_frame *d_src, int src_stride, int frames, gas_frame *d_bus,
		gas_frame *d_peaks, cudaStream_t st);
// gas_state.cu
cudaError_t launch_instance_init(gas_ctx *ctx, int n, const int32_t *d_ids, const int32_t *d_spat, cudaStream_t st);
cudaError_t launch_instance_stop(gas_ctx *ctx, int n, const int32_t *d_ids, cudaStream_t st);
cudaError_t launch_voice_init(gas_ctx *ctx, int n, const int32_t *d_ids, cudaStream_t st);
cudaError_t launch_state_export(gas_ctx *ctx, int n, const int32_t *d_ids, gas_voice_state *d_out, cudaStream_t st);
cudaError_t launch_state_import(gas_ctx *ctx, int n, const int32_t *d_ids, const gas_voice_state *d_in, cudaStream_t st);
cudaError_t launch_params_get(gas_ctx *ctx, int n, const int32_t *d_ids, gas_params *d_out, cudaStream_t st);
cudaError_t launch_fx_set(gas_ctx *ctx, int n, const int32_t *d_ids, const gas_effect_chain *d_in, cudaStream_t st);
cudaError_t launch_defaults(gas_ctx *ctx, cudaStream_t st);
