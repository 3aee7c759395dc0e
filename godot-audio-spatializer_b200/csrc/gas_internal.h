// gas_internal.h — context, device tables and per-block records shared by the .cu files.
// Product code: nothing here depends on the CPU checker that lives outside this package.
#pragma once

#include "../../include/gas.h"

#include <cuda_runtime.h>
#include <stdint.h>

#include "gas_ptx.cuh"

#include <mutex>
#include <string>
#include <vector>

// ---- limits of the per-block plan -----------------------------------------------------------------
#define GAS_MAX_SENDS 12      // union of current and previous bus details: 6 + 6
#define GAS_MAX_CLASSES 128   // class slots: distinct (path, mode, flags, send-mask, degree) combinations seen since the last reset
#define GAS_K2_MAX_ROWS 6     // polynomial rows per (pair, side) of a streamed voice: 2 (A, B) or 3 (A, B, C) per row group
#define GAS_K2_ROW_FLOATS (GAS_K2_MAX_ROWS * GAS_MAX_CHANNELS_PER_BUS * 2)
#define GAS_PLAN_DEPTH 4      // plans in flight: block b lives in slot b % 4 (the step kernel of block b plans block b + 1 while the
                              // voice-parallel kernels of blocks b - 1 and b may still be reading theirs)
#define GAS_PLAN_ROW_DEPTH 2  // weight rows are only read by the streaming part of the step kernel: block parity is enough

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
#define CLS_SHARED 2u // all sends carry identical weights: one row group fanned out to every bus of the mask
#define CLS_FILT 4u   // attenuation filter active (linear_attenuation >= 0.001)
#define CLS_SCALED 8u // every send is send 0 times one scalar per send (the class's aux word): one row group, scaled at the flush
#define CLS_AUX_NONE 0xffffffffffffffffULL // aux word of a free slot / of a slot whose claimer has not published it yet
#define CLS_GENERIC 1u // voice-parallel path only: the class of voices whose own class found no slot; one voice per unit
#define GAS_CLS_GENERIC_SLOTS 6 // (mode A / B / E) x (filter off / on), the last slots of the table, set up at creation
#define GAS_CLS_DYNAMIC (GAS_MAX_CLASSES - GAS_CLS_GENERIC_SLOTS) // slots claimed and recycled at run time
#define GAS_CLS_IDLE_BLOCKS 8 // a class slot that stayed empty for this many blocks is recycled

// A class is identified by a 64-bit key; everything a kernel needs to know about it is decoded from the key:
//   path [0,2)  mode [2,4)  flags [4,8)  n_send [8,12)  bus mask [16,32)  quad [32,33)
// CLS_SCALED classes are told apart by a second word as well (aux: the float bits of the scale of send 1 in the low
// half, of send 2 in the high half): two voices share a class only if their sends are the same multiples of send 0.
struct ClassInfo {
	unsigned long long key; // 0 = empty
	float scale[2];         // CLS_SCALED: send 1 = scale[0] * send 0, send 2 = scale[1] * send 0
	int32_t count;          // voices of the class in this block
	int32_t path;
	int32_t mode;
	uint32_t flags;
	uint32_t mask;   // bus mask of the sends
	uint32_t quad;   // 1 <=> every row group carries a t^2 row (some ramp product of the class is not linear in t)
	int32_t n_send;  // popcount(mask)
	int32_t n_group; // row groups: 1 if CLS_SHARED / CLS_SCALED else n_send
	int32_t n_rows;  // n_group * (2 + quad); 0 on the voice-parallel path
	int32_t slot;    // slot of the class in the global table (addresses its list)
	int32_t pad[2];  // 64 bytes: tables of these are copied 16 bytes at a time
};
static_assert(sizeof(ClassInfo) == 64, "ClassInfo is copied as int4 words");

static __host__ __device__ __forceinline__ unsigned long long cls_key(int path, int mode, uint32_t flags, int n_send, uint32_t mask, uint32_t quad) {
	return (unsigned long long)path | ((unsigned long long)mode << 2) | ((unsigned long long)flags << 4) | ((unsigned long long)n_send << 8) |
			((unsigned long long)mask << 16) | ((unsigned long long)quad << 32);
}
static __host__ __device__ __forceinline__ ClassInfo cls_decode(unsigned long long key, int count) {
	ClassInfo ci;
	ci.key = key;
	ci.count = count;
	ci.slot = 0;
	ci.path = (int32_t)(key & 3u);
	ci.mode = (int32_t)((key >> 2) & 3u);
	ci.flags = (uint32_t)((key >> 4) & 15u);
	ci.n_send = (int32_t)((key >> 8) & 15u);
	ci.mask = (uint32_t)((key >> 16) & 0xffffu);
	ci.quad = (uint32_t)((key >> 32) & 1u);
	ci.scale[0] = ci.scale[1] = 0.f;
	ci.n_group = ci.path == PATH_STREAM ? ((ci.flags & (CLS_SHARED | CLS_SCALED)) ? 1 : ci.n_send) : ci.n_send;
	ci.n_rows = ci.path == PATH_STREAM ? ci.n_group * (2 + (int)ci.quad) : 0;
	ci.pad[0] = ci.pad[1] = 0;
	return ci;
}

// Weight record of one streamed voice (floats): per row group g and pair c a 16-byte element {A_L, A_R, B_L, B_R}, then,
// for classes with a t^2 row, per (g, c) an 8-byte element {C_L, C_R}; padded to a multiple of 16 bytes.
static __host__ __device__ __forceinline__ int cls_row_floats(int n_group, int quad, int C) { return (n_group * C * (quad ? 6 : 4) + 3) & ~3; }

// Complete description of one planned block, written by the planner's last CTA and published through `seq`.
struct PlanHdr {
	int32_t seq;    // block index + 1 once the plan is complete (release store; readers acquire)
	int32_t n_cls;  // streaming classes with voices in this block (compact table below, slot order)
	int32_t n_vcls; // voice-parallel classes with voices in this block (0: the voice-parallel kernel has nothing to do)
	int32_t pad;
	ClassInfo cls[GAS_MAX_CLASSES];
	ClassInfo vcls[GAS_MAX_CLASSES];
};

// device-side counters (DevTables::blk), one per 32-byte sector
enum : int32_t {
	BLK_P = 0,       // plans produced so far = index of the next block to plan
	BLK_P_TICKET = 8,
	BLK_S = 16,      // step-kernel launches so far = index of the next block to stream
	BLK_S_TICKET = 24,
	BLK_Q = 32,      // voice-parallel launches so far = index of the next block it mixes
	BLK_Q_TICKET = 40,
	BLK_GAIN_DONE = 48, // control warps of the current step launch that have finished their gain tasks
	BLK_BAR_GEN = 56,
	BLK_WORDS = 64
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
	unsigned long long *cls_key; // [GAS_MAX_CLASSES] slot -> class key (0 = free); slots are stable across blocks
	unsigned long long *cls_aux; // [GAS_MAX_CLASSES] second word of the class identity (CLS_SCALED: the scales), CLS_AUX_NONE when unset
	int32_t *cls_idle;   // [GAS_MAX_CLASSES] consecutive blocks the slot stayed empty (recycled at GAS_CLS_IDLE_BLOCKS)
	int32_t *cls_count;  // [GAS_PLAN_DEPTH][GAS_MAX_CLASSES] by plan slot; the planner of block b clears the counts of slot (b + 1) % depth when it is done
	int32_t *overflow;   // [1] set when more than GAS_MAX_CLASSES classes were needed
	int2 *list;          // [GAS_PLAN_DEPTH][GAS_MAX_CLASSES][max_voices] {call-order index j, source row} per list position
	float *k2_rows;      // [GAS_PLAN_ROW_DEPTH][GAS_MAX_CLASSES][max_voices][GAS_K2_ROW_FLOATS] by list position (compact: cls_row_floats per voice)
	VoiceRec *rec;       // [GAS_PLAN_DEPTH][max_voices] by call-order index
	InstSends *sends;    // [GAS_PLAN_DEPTH][max_voices] by call-order index: resolved sends of the voice's instance (K3 voices only)
	PlanHdr *hdr;        // [GAS_PLAN_DEPTH]
};
static __host__ __device__ __forceinline__ int2 *plan_list(const BlockPlan &p, int slot, int cid, int maxv) {
	return p.list + ((size_t)slot * GAS_MAX_CLASSES + cid) * maxv;
}
static __host__ __device__ __forceinline__ float *plan_rows(const BlockPlan &p, int block, int cid, int maxv) {
	return p.k2_rows + ((size_t)(block & (GAS_PLAN_ROW_DEPTH - 1)) * GAS_MAX_CLASSES + cid) * maxv * GAS_K2_ROW_FLOATS;
}

struct DevTables {
	gas_spatializer *spat;
	int32_t *inst_spat;
	gas_params *inst_params;
	int32_t *inst_was_further;
	int32_t *inst_active;
	BusDetails *inst_cur;
	BusDetails *inst_prev;   // [2][max_instances]: double-buffered by block parity (read [p], write [1-p])
	int32_t *inst_mode;      // MODE_A/B/E | (effect_gain_binding + 1) << 8, latched at instantiate()
	int32_t *inst_seq;       // [max_instances] block index + 1 of the last in-kernel gain computation (release store): lets the planner of
	                         // that block start on a voice as soon as ITS instance's parameters are in place
	int32_t *blk;            // [BLK_WORDS] device-side block counters and tickets (BLK_*)
	int32_t max_instances;
	gas_effect_chain *inst_fx;
	float *vs_prev;              // [max_voices][4][2]
	gas_processor_state *vs_proc; // [max_voices][8]
	float *vs_fx;                // [max_voices][GAS_MAX_EFFECTS][2][GAS_MAX_FILTER_STAGES][4]
	// voice lifecycle (stream form of the mix)
	gas_frame *vs_look;          // [max_voices][64] lookahead
	uint32_t *vs_life;           // [max_voices] GAS_VOICE_ACTIVE | GAS_VOICE_HAS_FRAMES
	int32_t *vs_src;             // [max_voices] source slot the voice plays (-1: none), gas_voice_play
	long long *vs_start;         // [max_voices] first source frame of the playback
	unsigned long long *vs_pos;  // [max_voices] 16.16 fixed-point position since begin_resample
	float *inst_threshold;       // [max_instances] db_to_linear(playback_disable_threshold_db)
	float threshold_default;     // db_to_linear(-80 dB), evaluated on the host like the reference does (audio_spatializer.cpp:465)
	int32_t max_voices;
};

// one device-resident PCM source (gas_source_set)
struct SourceDesc {
	const gas_frame *pcm;
	int32_t n_frames;
	int32_t loop;
	float sample_rate;
	int32_t pad;
};

// what calculate_spatialization derives from a listener alone (computed once per listener upload)
struct ListenerPre {
	float inv[12];  // orthonormalised, then affine-inverted transform (basis rows + origin)
	float inv2[12]; // plain affine inverse (reverb-area position)
	float on[12];   // orthonormalised transform (doppler)
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
	// SPCAP constants of the current speaker mode (reference audio_spatializer_3d.cpp:47-55, :903-916), computed
	// on the host with the reference's float/double sequence: normalised speaker directions and the
	// "effective number of speakers" per speaker
	float spk_dir[7][3];
	float spk_eff[7];
};

struct gas_ctx {
	gas_config cfg;
	GlobalCfg g;
	int device = 0;
	int num_sms = 0;
	int l2_bytes = 0;
	cudaStream_t s_mix = nullptr, s_gain = nullptr, s_comm = nullptr; // s_comm: the multi-GPU exchange, beside the next block's mix
	cudaStream_t s_voice = nullptr; // the voice-parallel kernel of a block, beside its streaming kernel (par_voice)
	cudaEvent_t ev_voice_fork = nullptr, ev_voice_join = nullptr;
	bool par_voice = false;
	bool gain_after_stream = true; // GAS_K1_GATE=0: gain-side work no longer waits for the streaming kernel's CTAs to be resident
	cudaEvent_t ev_stream_started = nullptr; // programmatic event of the last streaming kernel launch
	bool stream_started_pending = false;
	bool scaled_classes = true; // GAS_K2_SCALED=0 turns the scaled-send classes off (experiments)
	bool k3_legacy = false;     // GAS_K3_LEGACY=1: the voice-parallel kernel keeps every class on its per-warp path (A/B against the filter-tile path)
	cudaEvent_t ev_gain_done = nullptr, ev_prologue_done = nullptr, ev_fork = nullptr, ev_join = nullptr, ev_mix_done = nullptr, ev_comm_done = nullptr, ev_join2 = nullptr;
	bool mix_pending = false, comm_pending = false;
	bool reduce_open = false; // gas_reduce_bus_begin_device without its _end yet
	bool gain_pending = false, prologue_pending = false;
	// pipelined form (gas_step_device): the block that has been planned and not streamed yet, and the voice-parallel kernels
	// still in flight on the side stream
	struct PlannedBlock {
		bool valid = false;
		int n_voices = 0, frames = 0, src_rows = 0;
		gas_frame *bus = nullptr, *peaks = nullptr;
	} planned;
	uint64_t step_count = 0;                        // pipelined steps launched (or captured) so far
	cudaEvent_t ev_step_done[GAS_PLAN_DEPTH] = {};  // recorded on the mix stream behind the step kernel of step % depth (and behind a priming plan)
	cudaEvent_t ev_block_done[GAS_PLAN_DEPTH] = {}; // recorded on the voice stream: step kernel AND voice-parallel kernel of that block are complete
	bool block_inflight[GAS_PLAN_DEPTH] = {};
	gas_frame *inflight_bus[GAS_PLAN_DEPTH] = {}, *inflight_peaks[GAS_PLAN_DEPTH] = {};
	bool step_done_valid = false; // ev_step_done[(step_count - 1) % depth] has been recorded
	// exchanges of a pipelined run still in flight on the exchange stream: a step only waits for the one that reads the buffer it is
	// about to clear (two steps of slack when the caller rotates four bus buffers), not for the latest
	cudaEvent_t ev_comm_ring[GAS_PLAN_DEPTH] = {};
	const gas_frame *comm_src[GAS_PLAN_DEPTH] = {};
	bool comm_ring_valid[GAS_PLAN_DEPTH] = {};
	uint64_t comm_count = 0;
	DevTables t{};
	BlockPlan plan{};
	// staging (device)
	gas_voice *d_voices = nullptr;
	gas_frame *d_src = nullptr;
	gas_frame *d_bus = nullptr;
	gas_frame *d_peaks = nullptr;
	gas_frame *d_rep = nullptr; // [replicas][num_buses][channels][frames]: K2 partial sums, combined by the K3 launch
	gas_frame *d_stage = nullptr;    // [max_voices][max_frames]: post-lookahead blocks of the stream form
	gas_voice *d_voices_stage = nullptr; // [max_voices] voice list rewritten by the lifecycle stage
	int32_t *d_mixed = nullptr, *d_status = nullptr; // [max_voices] staging of the host-pointer stream form
	int replicas = 8;           // GAS_K2_REPLICAS (1 = K2 adds straight into the bus buffers)
	gas_emitter *d_emitters = nullptr;
	gas_listener *d_listeners = nullptr;
	ListenerPre *d_listener_pre = nullptr;
	float bus_volume_lin[GAS_MAX_BUSES];    // bus graph (gas_bus_layout_set): linear volume per bus after mute / solo, 1 by default
	int32_t bus_send[GAS_MAX_BUSES];        // send target per bus (0 = Master by default)
	SourceDesc *d_sources = nullptr;        // [max_sources] device-resident PCM sources
	std::vector<SourceDesc> h_sources;      // host copy (device pointers owned by the context)
	int32_t max_sources = 0;
	gas_frame *d_rs_rows = nullptr;         // [max_voices][max_frames] rows the resampler hands to the stream form (gas_mix_block_resident)
	int32_t *d_rs_mixed = nullptr;          // [max_voices] // [GAS_MAX_LISTENERS] refreshed on the gain stream behind every listener upload
	gas_area *d_areas = nullptr;
	int32_t max_areas = 0;
	gas_params *d_params_out = nullptr;
	int32_t *d_ids = nullptr;  // scratch id list [max(max_voices,max_instances)]
	int32_t *d_ids2 = nullptr;
	void *d_scratch = nullptr; // gain-stream scratch for set/get payloads
	int32_t *d_ids_mix = nullptr; // mix-stream twins of d_ids / d_scratch (the two streams are not ordered against each other)
	void *d_scratch_mix = nullptr;
	size_t scratch_bytes = 0;
	int32_t inst_hwm = 0; // instances [0, inst_hwm) have been initialised at least once
	// multi-GPU exchange
	gas_frame *d_exchange = nullptr;
	int32_t comm_rank = 0, comm_ranks = 1;
	gas_frame *peer_exchange[8] = {};
	int comm_stride_f4 = 0;                  // 16-byte elements between the two parity buffers of an exchange allocation
	unsigned long long *d_comm_seq = nullptr; // [2] blocks pushed / finished so far (device-side, so that captured graphs stay valid)
	int *d_comm_ticket = nullptr;             // [2] CTA tickets of the two exchange kernels
	uint64_t launches = 0;
	bool k2_smem_attr_set = false, k3_smem_attr_set = false;
	int skip = 0;     // GAS_SKIP bits (experiments only)
	unsigned long long *d_timeline = nullptr; // GAS_K2_DEBUG & 8: per-CTA globaltimer stamps of the last K2 launch
	int pdl = 0; // GAS_PDL bit mask (experiments): programmatic dependent launch of 1 = prologue, 2 = streaming kernel, 4 = voice-parallel kernel
	int32_t n_listeners_res = 0, n_areas_res = 0; // resident listeners / areas (gas_listeners_set / gas_areas_set)
	// CUDA-graph capture
	bool capturing = false;
	uint64_t capture_launches0 = 0;
	struct Graph {
		cudaGraphExec_t exec = nullptr;
		uint64_t kernels = 0;
		bool profiled = false; // captured while per-kernel timing was on: carries event-record nodes around the kernels
		int prof_blocks = 0;   // mix blocks in the graph (only the last block's events survive a launch)
		uint64_t cfg_epoch = 0; // AudioServer globals the graph was captured under (speaker mode, mix rate, panning strength are baked in)
	};
	uint64_t cfg_epoch = 0; // bumped by gas_set_speaker_mode / gas_set_mix_rate / gas_set_global_panning_strength
	std::vector<Graph> graphs;
	// per-kernel timing
	bool profiling = false;
	struct ProfPair {
		cudaEvent_t a, b;
		int kind;
		cudaStream_t st;
	};
	std::vector<ProfPair> prof_pairs;
	ProfPair graph_pair[GAS_KERNEL_KINDS] = {}; // the event pairs of a profiled capture
	cudaEvent_t gev[GAS_KERNEL_KINDS][2] = {}; // event-record nodes of profiled graphs
	bool gev_used[GAS_KERNEL_KINDS] = {};
	bool capture_profiled = false;
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

// Launch with (optionally) the programmatic-dependent-launch attribute: the kernel may become resident
// while its stream predecessor drains; it must execute griddepcontrol.wait before touching global memory.
// `started` (optional): a timing-disabled event that fires once every CTA of the grid has started (programmatic
// event, triggered at block start): work on another stream that waits for it runs BESIDE this kernel instead of
// racing it for the SMs at launch.
template <typename... KArgs, typename... Args>
static inline cudaError_t gas_launch_ev(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, cudaEvent_t started,
		Args... args) {
	cudaLaunchConfig_t lc{};
	lc.gridDim = grid;
	lc.blockDim = block;
	lc.dynamicSmemBytes = smem;
	lc.stream = st;
	cudaLaunchAttribute attr[2];
	int na = 0;
	if (pdl) {
		attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		attr[na].val.programmaticStreamSerializationAllowed = 1;
		na++;
	}
	if (started) {
		attr[na].id = cudaLaunchAttributeProgrammaticEvent;
		attr[na].val.programmaticEvent.event = started;
		attr[na].val.programmaticEvent.flags = 0;
		attr[na].val.programmaticEvent.triggerAtBlockStart = 1;
		na++;
	}
	lc.attrs = attr;
	lc.numAttrs = na;
	cudaError_t e = cudaLaunchKernelEx(&lc, kernel, static_cast<KArgs>(args)...);
	return e != cudaSuccess ? e : cudaGetLastError();
}
template <typename... KArgs, typename... Args>
static inline cudaError_t gas_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
	return gas_launch_ev(kernel, grid, block, smem, st, pdl, (cudaEvent_t) nullptr, args...);
}
// (GAS_GRID_DEP_WAIT / GAS_GRID_DEP_LAUNCH: gas_ptx.cuh)

// ---- kernel launchers (each returns cudaError_t from the launch) -----------------------------------
// gas_gain.cu
cudaError_t launch_gain(gas_ctx *ctx, int n, const gas_emitter *d_em, int n_listeners, const gas_listener *d_l,
		const gas_area *d_areas, gas_params *d_out, cudaStream_t st);
cudaError_t launch_listener_pre(gas_ctx *ctx, int n_listeners, cudaStream_t st);
cudaError_t launch_params_set(gas_ctx *ctx, int n, const int32_t *d_ids, const gas_params *d_params, cudaStream_t st);
cudaError_t launch_instance_start(gas_ctx *ctx, int n, const int32_t *d_ids, cudaStream_t st);
// gas_prologue.cu
cudaError_t launch_plan(gas_ctx *ctx, int n_voices, const gas_voice *d_voices, int src_rows, int frames, gas_frame *d_bus,
		gas_frame *d_peaks, cudaStream_t st);
// frames of one replica of the partial-sum buffers for a block of `frames` frames (16-byte units)
static inline int gas_bus_f4(const gas_ctx *ctx, int frames) { return ctx->g.num_buses * ctx->g.channels * frames / 2; }
// gas_mix_stream.cu (the step kernel) / gas_mix_voice.cu (K3)
// what the control warps of a step launch prepare: gains (n_emitters > 0, resident listeners / areas) and plan of the next block
struct StepNext {
	int n_emitters;
	const gas_emitter *d_emitters;
	int n_voices;
	const gas_voice *d_voices;
	int src_rows;
	int frames;
	gas_frame *d_bus;
	gas_frame *d_peaks;
};
cudaError_t launch_step(gas_ctx *ctx, const gas_frame *d_src, int src_stride, int frames, gas_frame *d_bus, const StepNext *next, cudaStream_t st,
		bool pdl);
// after_stream: launched right behind the streaming kernel on the same stream (its class-table look may then precede the dependency wait)
cudaError_t launch_mix_voice(gas_ctx *ctx, const gas_frame *d_src, int src_stride, int frames, gas_frame *d_bus,
		gas_frame *d_peaks, cudaStream_t st, bool after_stream);
// gas_comm.cu
cudaError_t launch_comm_push(gas_ctx *ctx, const gas_frame *d_bus, int frames, cudaStream_t st);
cudaError_t launch_comm_finish(gas_ctx *ctx, gas_frame *d_bus, int frames, cudaStream_t st);
cudaError_t launch_comm_exchange(gas_ctx *ctx, const gas_frame *d_partial, gas_frame *d_prev_sum, int frames, cudaStream_t st);
// gas_life.cu
cudaError_t launch_life_stage(gas_ctx *ctx, int n_voices, const gas_voice *d_voices, const int32_t *d_mixed, const gas_frame *d_src, int src_rows,
		int src_stride, int frames, gas_voice *d_out_voices, gas_frame *d_stage, int stage_stride, cudaStream_t st);
cudaError_t launch_life_post(gas_ctx *ctx, int n_voices, const gas_voice *d_voices, const gas_frame *d_peaks, int32_t *d_status, cudaStream_t st);
cudaError_t launch_threshold_set(gas_ctx *ctx, int n, const int32_t *d_ids, const float *d_lin, cudaStream_t st);
cudaError_t launch_life_export(gas_ctx *ctx, int n, const int32_t *d_ids, gas_voice_life *d_out, cudaStream_t st);
cudaError_t launch_life_import(gas_ctx *ctx, int n, const int32_t *d_ids, const gas_voice_life *d_in, cudaStream_t st);
// gas_bus.cu
cudaError_t launch_bus_graph(gas_ctx *ctx, gas_frame *d_bus, int frames, cudaStream_t st);
// gas_resample.cu
cudaError_t launch_resample(gas_ctx *ctx, int n_voices, const gas_voice *d_voices, int frames, gas_frame *d_rows, int row_stride, int src_rows,
		int32_t *d_mixed, cudaStream_t st);
cudaError_t launch_voice_play(gas_ctx *ctx, int n, const int32_t *d_voices, const int32_t *d_sources, const int32_t *d_starts, cudaStream_t st);
// gas_single.cu: channel < 0 = process_frames, else mix_channel of that pair
cudaError_t launch_single_voice(gas_ctx *ctx, int instance, int voice, int channel, gas_frame *d_out, const gas_frame *d_src, int frames, cudaStream_t st);
// gas_state.cu
cudaError_t launch_instance_init(gas_ctx *ctx, int n, const int32_t *d_ids, const int32_t *d_spat, cudaStream_t st);
cudaError_t launch_instance_stop(gas_ctx *ctx, int n, const int32_t *d_ids, cudaStream_t st);
cudaError_t launch_voice_init(gas_ctx *ctx, int n, const int32_t *d_ids, cudaStream_t st);
cudaError_t launch_state_export(gas_ctx *ctx, int n, const int32_t *d_ids, gas_voice_state *d_out, cudaStream_t st);
cudaError_t launch_state_import(gas_ctx *ctx, int n, const int32_t *d_ids, const gas_voice_state *d_in, cudaStream_t st);
cudaError_t launch_params_get(gas_ctx *ctx, int n, const int32_t *d_ids, gas_params *d_out, cudaStream_t st);
cudaError_t launch_fx_set(gas_ctx *ctx, int n, const int32_t *d_ids, const gas_effect_chain *d_in, cudaStream_t st);
cudaError_t launch_defaults(gas_ctx *ctx, cudaStream_t st);
