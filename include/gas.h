/*
 * gas.h — C ABI of the B200-native batched spatial-audio mixer ("gas" = godot audio spatializer).
 *
 * This is the drop-in boundary for the data-parallel hot path of BuzzLord/godot-audio-spatializer:
 * per-instance gain computation + per-voice volume-ramped / filtered mixing of AudioFrame buffers into
 * bus channel buffers.  The reference has no C ABI; the closest thing is the raw-pointer GDExtension
 * virtuals on AudioSpatializerInstance (reference audio_spatializer.h:103-112, :144-150).  Every entry
 * point below names the reference interface it replaces.  Plain C types only: no torch, no C++.
 *
 * Vocabulary (same as the reference):
 *   instance  = one AudioSpatializerInstance (one per AudioStreamPlayerSpatial).  Owns the
 *               SpatializerParameters of the current physics tick and the AudioServer-side bus
 *               volume state of its proxy playbacks.
 *   voice     = one SpatialPlaybackListNode (audio_spatializer.h:55-66): one live playback of an
 *               instance.  Owns SpatializerPlaybackData (prev mix volumes, filter processors).
 *   pair      = one stereo channel pair of a bus (0 front L/R, 1 centre/LFE, 2 rear L/R, 3 side L/R).
 *   bus       = AudioServer bus, addressed by index (0 = Master).  Unknown index => Master, like
 *               AudioStreamPlayerSpatial::get_bus (audio_stream_player_spatial.cpp:405-413).
 *
 * Error convention: every call returns a gas_status (0 = OK).  Nothing throws, nothing aborts; on an
 * invalid argument the context is left untouched (reference: ERR_FAIL_* macros, e.g.
 * audio_spatializer_3d.cpp:670-672, spatializer_parameters.cpp:35-47).  gas_last_error() returns the
 * message of the last failing call.  There is NO CPU fallback: without a CUDA device gas_create fails.
 *
 * Threading (reference audio_spatializer.h:135-139, README.md:26,32,51): gas_gain_compute /
 * gas_params_set / gas_effect_params_set may be called from a "physics" thread while one "audio"
 * thread calls gas_mix_block*; parameter hand-off is double-buffered inside the context and swapped
 * under a mutex (reference audio_spatializer.cpp:558-574).  gas_mix_block* is single-caller.
 */
#ifndef GAS_H
#define GAS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define GAS_API __declspec(dllexport)
#else
#define GAS_API __attribute__((visibility("default")))
#endif

/* ---- limits (reference audio_spatializer.h:47-52) ---------------------------------------------- */
#define GAS_MAX_CHANNELS_PER_BUS 4
#define GAS_LOOKAHEAD_BUFFER_SIZE 64
#define GAS_MAX_BUSES_PER_PLAYBACK 6
#define GAS_MAX_LISTENERS 8
#define GAS_MAX_EFFECTS 4        /* AudioEffectFilter instances per AudioSpatializerEffect chain */
#define GAS_MAX_FILTER_STAGES 4  /* AudioEffectFilter FILTER_6DB..FILTER_24DB */
#define GAS_MAX_BUSES 16         /* bus indices a context can mix into */
#define GAS_ABI_VERSION 3

typedef enum gas_status {
	GAS_OK = 0,
	GAS_ERR_INVALID = 1, /* bad argument; context untouched */
	GAS_ERR_CUDA = 2,    /* CUDA runtime/driver error; message has the CUDA string */
	GAS_ERR_NOMEM = 3,
	GAS_ERR_STATE = 4,   /* call not valid in the current state (e.g. comm not initialised) */
	GAS_ERR_NCCL = 5,
	GAS_ERR_NO_DEVICE = 6 /* no usable sm_100 device: there is no CPU fallback */
} gas_status;

/* AudioFrame (upstream servers/audio/audio_frame.h): 2 x fp32, 8 bytes. */
typedef struct gas_frame {
	float l, r;
} gas_frame;

/* AudioServer::SpeakerMode; channel pair count = mode + 1 (AudioServer::get_channel_count). */
typedef enum gas_speaker_mode {
	GAS_SPEAKER_MODE_STEREO = 0,
	GAS_SPEAKER_SURROUND_31 = 1,
	GAS_SPEAKER_SURROUND_51 = 2,
	GAS_SPEAKER_SURROUND_71 = 3
} gas_speaker_mode;

/* AudioSpatializer3D::AttenuationModel (audio_spatializer_3d.h:157-162). */
typedef enum gas_attenuation_model {
	GAS_ATTENUATION_INVERSE_DISTANCE = 0,
	GAS_ATTENUATION_INVERSE_SQUARE_DISTANCE = 1,
	GAS_ATTENUATION_LOGARITHMIC = 2,
	GAS_ATTENUATION_DISABLED = 3
} gas_attenuation_model;

/* AudioSpatializer3D::DopplerTracking (audio_spatializer_3d.h:164-168). */
typedef enum gas_doppler_tracking {
	GAS_DOPPLER_TRACKING_DISABLED = 0,
	GAS_DOPPLER_TRACKING_IDLE_STEP = 1,
	GAS_DOPPLER_TRACKING_PHYSICS_STEP = 2
} gas_doppler_tracking;

typedef enum gas_spatializer_kind {
	GAS_SPATIALIZER_3D = 0,    /* AudioSpatializer3D (audio_spatializer_3d.h:153-241) */
	GAS_SPATIALIZER_EFFECT = 1 /* AudioSpatializerEffect (audio_spatializer_effect.h:83-96), filter-chain subset */
} gas_spatializer_kind;

/* upstream AudioFilterSW::Mode (servers/audio/audio_filter_sw.h). */
typedef enum gas_filter_mode {
	GAS_FILTER_BANDPASS = 0,
	GAS_FILTER_HIGHPASS = 1,
	GAS_FILTER_LOWPASS = 2,
	GAS_FILTER_NOTCH = 3,
	GAS_FILTER_PEAK = 4,
	GAS_FILTER_BANDLIMIT = 5,
	GAS_FILTER_LOWSHELF = 6,
	GAS_FILTER_HIGHSHELF = 7
} gas_filter_mode;

/* One AudioEffectFilter of an AudioSpatializerEffect chain (upstream AudioEffectFilter properties;
 * example gd_spatializer.gd:14-19).  `stages` = int(db)+1, 1..4. */
typedef struct gas_effect {
	int32_t mode; /* gas_filter_mode */
	float cutoff_hz;
	float resonance;
	float gain;
	int32_t stages;
} gas_effect;

typedef struct gas_effect_chain {
	int32_t n_effects; /* 0..GAS_MAX_EFFECTS; 0 => process_frames is a copy (audio_spatializer_effect.cpp:41-46) */
	gas_effect effects[GAS_MAX_EFFECTS];
} gas_effect_chain;

/* AudioSpatializer3D resource properties (audio_spatializer_3d.h:171-188), plus the effect chain of an
 * AudioSpatializerEffect.  Defaults: see gas_spatializer_defaults(). */
typedef struct gas_spatializer {
	int32_t kind;              /* gas_spatializer_kind */
	int32_t attenuation_model; /* gas_attenuation_model */
	float unit_size;
	float max_distance;
	float panning_strength;
	uint32_t area_mask; /* carried, unused on device (physics query is the caller's) */
	int32_t emission_angle_enabled;
	float emission_angle; /* degrees, [0, 90] */
	float emission_angle_filter_attenuation_db;
	float attenuation_filter_cutoff_hz;
	float attenuation_filter_db;
	int32_t doppler_tracking; /* gas_doppler_tracking */
	float doppler_speed_of_sound;
	int32_t mix_channel_mode; /* 0 = Mode A (_process_frames), 1 = Mode B (_mix_channel); EFFECT is always 0 */
	/* EFFECT kind only: effect whose `gain` is bound to SpatializerParameters3D::linear_attenuation each
	 * block, as the example's _process_effects does (gd_spatializer_instance.gd:125-127); -1 = none. */
	int32_t effect_gain_binding;
	gas_effect_chain chain;
} gas_spatializer;

/* Listener = Camera3D / AudioListener3D global transform + doppler velocity
 * (audio_spatializer_3d.cpp:335-344, :408-412).  basis is row-major rows[3][3] like Godot's Basis. */
typedef struct gas_listener {
	float basis[9];
	float origin[3];
	float velocity[3];
} gas_listener;

/* Result of the physics-side Area3D query for one emitter (audio_spatializer_3d.cpp:206-245, :346-354).
 * closest_point[l] = PhysicsDirectSpaceState3D::get_closest_point_to_object_volume(area, listener l origin). */
typedef struct gas_area {
	int32_t override_bus; /* Area3D::is_overriding_audio_bus */
	int32_t bus;          /* Area3D::get_audio_bus_name, as bus index */
	int32_t use_reverb;   /* Area3D::is_using_reverb_bus */
	int32_t reverb_bus;   /* Area3D::get_reverb_bus_name, as bus index */
	float reverb_amount;
	float reverb_uniformity;
	float closest_point[GAS_MAX_LISTENERS][3];
} gas_area;

/* Per-instance inputs of calculate_spatialization (audio_spatializer_3d.cpp:277-489): what the
 * reference reads from get_audio_player() and the scene. */
typedef struct gas_emitter {
	int32_t instance;    /* instance slot the result is stored for */
	int32_t spatializer; /* gas_spatializer slot (the shared resource) */
	int32_t area;        /* index into the areas array, -1 = none */
	int32_t bus;         /* AudioStreamPlayerSpatial::get_bus as index (out of range => 0 = Master) */
	float origin[3];     /* get_global_transform().origin */
	float basis_z[3];    /* get_global_transform().basis.get_column(2) (emission direction is -Z) */
	float velocity[3];   /* VelocityTracker3D::get_tracked_linear_velocity */
	float volume_db;
	float max_db;
	float pitch_scale;
} gas_emitter;

/* SpatializerParameters + SpatializerParameters3D (spatializer_parameters.h:39-67,
 * audio_spatializer_3d.h:61-83) as one POD record.  bus/bus_volumes keep Dictionary insertion order. */
typedef struct gas_params {
	float mix_volumes[GAS_MAX_CHANNELS_PER_BUS][2];
	float pitch_scale;
	float linear_attenuation;            /* aka high-shelf gain */
	float attenuation_filter_cutoff_hz;
	int32_t update_parameters;
	int32_t n_bus;                       /* 0..GAS_MAX_BUSES_PER_PLAYBACK */
	int32_t bus[GAS_MAX_BUSES_PER_PLAYBACK];
	float bus_volumes[GAS_MAX_BUSES_PER_PLAYBACK][GAS_MAX_CHANNELS_PER_BUS][2];
} gas_params;

/* One voice of a mix block. */
#define GAS_VOICE_WANT_PEAK 1u /* compute this voice's block peak (reference computes it always,
                                  audio_spatializer.cpp:419-461, but only reads it when !has_frames, :464) */
typedef struct gas_voice {
	int32_t voice;    /* playback-data slot */
	int32_t instance; /* owning instance slot */
	int32_t src_row;  /* row of `src` holding this voice's frames; -1 = silence (the reference's
	                     zero-filled playback_buffer once has_frames is clear, audio_spatializer.cpp:405-408) */
	uint32_t flags;
} gas_voice;

/* upstream AudioFilterSW::Processor persistent state. */
typedef struct gas_processor_state {
	float b0, b1, b2, a1, a2; /* current (interpolating) coefficients; a1/a2 stored negated */
	float ha1, ha2, hb1, hb2; /* history */
} gas_processor_state;

/* SpatializerPlaybackData3D (audio_spatializer_3d.h:85-99) + effect-chain histories
 * (SpatializerPlaybackDataEffect's AudioEffectFilterInstance::filter_process[2][4]). */
typedef struct gas_voice_state {
	float prev_mix_volumes[GAS_MAX_CHANNELS_PER_BUS][2];
	gas_processor_state filter_processors[2 * GAS_MAX_CHANNELS_PER_BUS]; /* index pair*2 + (left?0:1), :887-894 */
	float effect_history[GAS_MAX_EFFECTS][2][GAS_MAX_FILTER_STAGES][4];  /* [effect][l/r][stage]{ha1,ha2,hb1,hb2} */
} gas_voice_state;

/* The rest of SpatialPlaybackListNode (audio_spatializer.h:55-66) kept per voice slot for the stream form of the mix:
 * flags bit 0 = active, bit 1 = has_frames. */
#define GAS_VOICE_ACTIVE 1u
#define GAS_VOICE_HAS_FRAMES 2u
typedef struct gas_voice_life {
	gas_frame lookahead[GAS_LOOKAHEAD_BUFFER_SIZE];
	uint32_t flags;
} gas_voice_life;

typedef struct gas_config {
	int32_t device; /* CUDA ordinal */
	int32_t max_instances;
	int32_t max_voices;
	int32_t max_frames;       /* largest block size (frames) */
	int32_t max_spatializers; /* gas_spatializer slots */
	int32_t num_buses;        /* 1..GAS_MAX_BUSES */
	int32_t speaker_mode;     /* gas_speaker_mode */
	float mix_rate;           /* AudioServer::get_mix_rate */
	float global_panning_strength; /* project setting audio/general/3d_panning_strength (audio_spatializer_3d.cpp:633) */
} gas_config;

typedef struct gas_ctx gas_ctx;

/* ---- lifetime ---------------------------------------------------------------------------------- */
GAS_API int gas_abi_version(void);
/* sizeof() of the records above as compiled into the library, so a foreign-language binding can check
 * its own layout before the first call.  Returns 0 for an unknown id. */
typedef enum gas_struct_id {
	GAS_STRUCT_FRAME = 0,
	GAS_STRUCT_EFFECT = 1,
	GAS_STRUCT_EFFECT_CHAIN = 2,
	GAS_STRUCT_SPATIALIZER = 3,
	GAS_STRUCT_LISTENER = 4,
	GAS_STRUCT_AREA = 5,
	GAS_STRUCT_EMITTER = 6,
	GAS_STRUCT_PARAMS = 7,
	GAS_STRUCT_VOICE = 8,
	GAS_STRUCT_PROCESSOR_STATE = 9,
	GAS_STRUCT_VOICE_STATE = 10,
	GAS_STRUCT_CONFIG = 11,
	GAS_STRUCT_VOICE_LIFE = 12,
	GAS_STRUCT_BUS_DESC = 13,
	GAS_STRUCT_STEP_NEXT = 14
} gas_struct_id;
GAS_API size_t gas_abi_sizeof(int32_t struct_id);
GAS_API void gas_config_defaults(gas_config *cfg);
/* Replaces module registration + AudioServer globals the path reads (register_types.cpp:40-60). */
GAS_API int gas_create(const gas_config *cfg, gas_ctx **out_ctx);
GAS_API void gas_destroy(gas_ctx *ctx);
/* Message of the last failing call on ctx (or of the last failing gas_create when ctx == NULL). */
GAS_API const char *gas_last_error(const gas_ctx *ctx);
/* AudioServer::get_speaker_mode / get_mix_rate / project setting, changed at run time. */
GAS_API int gas_set_speaker_mode(gas_ctx *ctx, int32_t speaker_mode);
GAS_API int gas_set_mix_rate(gas_ctx *ctx, float mix_rate);
GAS_API int gas_set_global_panning_strength(gas_ctx *ctx, float strength);
GAS_API int gas_get_channel_count(const gas_ctx *ctx);

/* ---- resources / slots -------------------------------------------------------------------------- */
/* AudioSpatializer3D defaults (audio_spatializer_3d.h:171-188). */
GAS_API void gas_spatializer_defaults(gas_spatializer *s);
/* AudioSpatializer3D setters with their validation (audio_spatializer_3d.cpp:654-765): max_distance>=0,
 * emission_angle in [0,90], attenuation_model<4, panning_strength>=0, doppler_speed_of_sound>0. */
GAS_API int gas_spatializer_set(gas_ctx *ctx, int32_t slot, const gas_spatializer *s);
/* AudioSpatializer::instantiate (audio_spatializer_3d.cpp:645-652): binds instances to a spatializer
 * slot and resets all per-instance state (parameters, was_further_than_max_distance_last_frame,
 * AudioServer-side bus details). */
GAS_API int gas_instance_init(gas_ctx *ctx, int32_t n, const int32_t *instances, const int32_t *spatializers);
/* First voice of an inactive instance: the proxy playbacks are (re)registered with AudioServer using
 * get_bus_map(current parameters); previous bus details start empty => fade-in
 * (audio_spatializer.cpp:75-95; upstream AudioServer::start_playback_stream). */
GAS_API int gas_instance_start(gas_ctx *ctx, int32_t n, const int32_t *instances);
/* AudioSpatializerInstance::_manage_playback_state stopping the proxies (audio_spatializer.cpp:484-491). */
GAS_API int gas_instance_stop(gas_ctx *ctx, int32_t n, const int32_t *instances);
/* instantiate_playback_data (audio_spatializer_3d.cpp:200-204, audio_spatializer_effect.cpp:79-88): zero state. */
GAS_API int gas_voice_init(gas_ctx *ctx, int32_t n, const int32_t *voices);

/* ---- gain side (physics thread) ----------------------------------------------------------------- */
/* Batched AudioSpatializerInstance3D::calculate_spatialization + update_spatializer_parameters
 * (audio_spatializer_3d.cpp:277-489, audio_spatializer.cpp:258-272): for every emitter, computes the
 * SpatializerParameters3D of its instance on the GPU, stores them as the instance's current
 * parameters and, when update_parameters is set, pushes get_bus_map() to the AudioServer-side state
 * (audio_spatializer.cpp:274-324).  `areas` may be NULL when no emitter references one.
 * out_params (optional, host, n entries) receives a copy of the computed parameters. */
GAS_API int gas_gain_compute(gas_ctx *ctx, int32_t n, const gas_emitter *emitters,
		int32_t n_listeners, const gas_listener *listeners,
		int32_t n_areas, const gas_area *areas, gas_params *out_params);
/* Same with device-resident emitters (listeners/areas are small and stay host pointers); asynchronous
 * on the context's gain stream. */
GAS_API int gas_gain_compute_device(gas_ctx *ctx, int32_t n, const gas_emitter *d_emitters,
		int32_t n_listeners, const gas_listener *listeners,
		int32_t n_areas, const gas_area *areas, gas_params *d_out_params);
/* Device-resident copies of the listeners / areas, used by gas_gain_compute_device when its
 * `listeners` (resp. `areas`) argument is NULL.  Needed inside a captured graph, where host-to-device
 * copies from pageable memory are not allowed. */
GAS_API int gas_listeners_set(gas_ctx *ctx, int32_t n_listeners, const gas_listener *listeners);
GAS_API int gas_areas_set(gas_ctx *ctx, int32_t n_areas, const gas_area *areas);
/* Hand-off of parameters computed elsewhere (a custom _calculate_spatialization): the batched
 * set_spatializer_parameters + bus-map push (audio_spatializer.cpp:258-272, :558-564).  Validation:
 * n_bus <= 6 (spatializer_parameters.cpp:35-47 sizes are fixed by the struct). */
GAS_API int gas_params_set(gas_ctx *ctx, int32_t n, const int32_t *instances, const gas_params *params);
GAS_API int gas_params_get(gas_ctx *ctx, int32_t n, const int32_t *instances, gas_params *out_params);
/* Per-instance effect parameters for the next blocks: what a _process_effects override would write
 * into its AudioEffectFilter resources (audio_spatializer_effect.cpp:39, :90-92). */
GAS_API int gas_effect_params_set(gas_ctx *ctx, int32_t n, const int32_t *instances, const gas_effect_chain *chains);

/* ---- mix side (audio thread) -------------------------------------------------------------------- */
/* One mix block for all voices of all instances: batched
 * AudioSpatializerInstance::_mix_from_playback_list (audio_spatializer.cpp:326-471) calling
 * process_frames (audio_spatializer_3d.cpp:491-552, audio_spatializer_effect.cpp:33-77) or mix_channel
 * (audio_spatializer_3d.cpp:554-609) per voice, followed by the AudioServer bus accumulate of the
 * proxy playbacks (upstream AudioServer::_mix_step_for_channel, fed by audio_spatializer.cpp:93,269).
 *   voices   n_voices descriptors (host); a voice slot may appear at most once per block.
 *   src      src_rows x frames AudioFrames (host): the post-lookahead playback buffers
 *            (audio_spatializer.cpp:367-408), row r used by voices with src_row == r.
 *   frames   block size, even, <= max_frames (the reference accepts any count; frames travel as 16-byte pairs here, and
 *            AudioServer's block is 512).
 *   bus_out  host, [num_buses][channel_count][frames] AudioFrames, fully overwritten.
 *   peaks    optional host, n_voices entries: per-voice block peak (max |sample| over pairs, l/r
 *            separately), valid for voices flagged GAS_VOICE_WANT_PEAK, else {0,0}.
 * Synchronous: returns after bus_out/peaks are written. */
GAS_API int gas_mix_block(gas_ctx *ctx, int32_t n_voices, const gas_voice *voices,
		const gas_frame *src, int32_t src_rows, int32_t frames,
		gas_frame *bus_out, gas_frame *peaks);
/* Same with everything device-resident; asynchronous on the context's mix stream.  d_bus_out is
 * overwritten; d_peaks may be NULL.  src_row_stride in frames (>= frames, even). */
GAS_API int gas_mix_block_device(gas_ctx *ctx, int32_t n_voices, const gas_voice *d_voices,
		const gas_frame *d_src, int32_t src_rows, int32_t src_row_stride, int32_t frames,
		gas_frame *d_bus_out, gas_frame *d_peaks);
/* ---- pipelined form: one kernel launch per block, for callers that render blocks back to back -------------------------------
 * No reference analogue in structure (the reference mixes when the audio driver asks), same results: gas_step_device streams the
 * block that the PREVIOUS call planned into that block's bus buffers and, in the same launch and on the same SMs, does for the
 * NEXT block what gas_gain_compute_device (calculate_spatialization + the bus-map push, with the resident listeners / areas of
 * gas_listeners_set / gas_areas_set) and the planning part of gas_mix_block_device do: the latency-bound per-instance / per-voice
 * work of block k + 1 hides behind the bandwidth-bound streaming of block k, and consecutive launches overlap (programmatic
 * dependent launch) because block indices and plans live on the device.
 *   d_src / src_row_stride  source rows of the block being streamed (ignored by the first call of a run, which only plans)
 *   next                    the block to prepare: its emitters (n_emitters = 0 keeps the current parameters), voice list, source
 *                           row count, frame count and output buffers (zeroed by this call, complete after the NEXT one);
 *                           NULL ends the run (the planned block is streamed, nothing is prepared)
 * next->d_bus_out / d_peaks must differ from the buffers of the block being streamed; reusing the buffers of the block before
 * that is allowed but costs overlap (rotate over three or more).  Voices that need the voice-parallel kernel (filters, effect
 * chains, peaks) are mixed by a launch on a side stream: a block's buffers are complete for work on the mix stream after
 * gas_step_join_device (gas_sync, gas_mix_block* and the reduce calls join by themselves).  gas_mix_block* is refused
 * (GAS_ERR_STATE) while a planned block is waiting.  Captured into graphs like the other device-resident calls: a replayed graph
 * must then meet the device in the state it was captured in (a planned block waiting, or none).
 * Residency: a step launch takes one whole SM per CTA.  Its control warps wait for the gains of an instance only where a voice or an
 * instance is planned by another lane pair than the one that computed its gains: with emitter i, instance i and voice i at the same
 * index (the layout bench.py uses) no CTA waits for another one, otherwise CTAs of a launch wait for each other and the launch needs
 * all of them resident.  So: while a pipelined run is in flight do not enqueue, from another stream of this GPU, a kernel that
 * spin-waits on other GPUs (an NCCL collective, a hand-written barrier) — it can take an SM the launch needs and wait for peers whose
 * own SMs are held the same way.  Synchronise (gas_sync) before such a call; the library's own exchange
 * (gas_reduce_bus_exchange_device) is ordered so that it cannot close that cycle. */
typedef struct gas_step_next {
	int32_t n_emitters;
	const gas_emitter *d_emitters;
	int32_t n_voices;
	const gas_voice *d_voices;
	int32_t src_rows;
	int32_t frames;
	gas_frame *d_bus_out;
	gas_frame *d_peaks; /* optional */
} gas_step_next;
GAS_API int gas_step_device(gas_ctx *ctx, const gas_frame *d_src, int32_t src_row_stride, const gas_step_next *next);
GAS_API int gas_step_join_device(gas_ctx *ctx);
/* ---- the per-call virtuals, one voice per call (reference audio_spatializer.h:146,148; GDVIRTUAL mirror :103-112) --------
 * Same argument meaning as the reference: the instance's current SpatializerParameters and the voice's SpatializerPlaybackData
 * are the two Ref arguments (addressed by slot), out/src are caller-owned host arrays of `frames` AudioFrames, out is fully
 * overwritten (Q13), the playback data advance.
 *   gas_process_frames  AudioSpatializerInstance3D::process_frames (audio_spatializer_3d.cpp:491-552) for a 3D instance,
 *                       AudioSpatializerInstanceEffect::process_frames (audio_spatializer_effect.cpp:33-77) for an EFFECT one
 *   gas_mix_channel     AudioSpatializerInstance3D::mix_channel (audio_spatializer_3d.cpp:554-609) for pair `channel` (0..3)
 * The batched mix never calls these: they are what a subclass that overrides one of the two virtuals calls for the built-in
 * behaviour, and how a single playback is processed outside a mix step.  Synchronous; any frame count in [1, max_frames]. */
GAS_API int gas_process_frames(gas_ctx *ctx, int32_t instance, int32_t voice, gas_frame *out, const gas_frame *src, int32_t frames);
GAS_API int gas_mix_channel(gas_ctx *ctx, int32_t instance, int32_t voice, int32_t channel, gas_frame *out, const gas_frame *src,
		int32_t frames);

/* ---- stream form: the voice lifecycle of _mix_from_playback_list on the device (audio_spatializer.cpp:353-408, :464-492) ----
 * Row r of `src` holds what AudioStreamPlayback::mix returned for the voice this block (audio_spatializer.cpp:378):
 * mixed_frames[i] <= frames new frames, NOT yet spliced behind the lookahead.  Per voice slot the context keeps the node's
 * 64-frame lookahead, `has_frames` and `active` (set / zeroed by gas_voice_init, like start_playback_stream :57-72) and does
 * what the reference does around the per-voice call: inactive voices are skipped (:355); the lookahead goes in front of the
 * new frames and the last 64 frames become the next lookahead (:369-373, :401-403); a short count ends the stream — the last
 * 64 valid frames are faded with 0.96^k * (64 - k) / 64, the rest is zeroed, has_frames is cleared (:380-398); a voice without
 * frames is mixed with silence (filter tails, :405-408) and deactivated once its block peak is at or below its instance's
 * playback_disable_threshold_db (:464-469).  status_out[i] (optional): GAS_VOICE_ACTIVE | GAS_VOICE_HAS_FRAMES after the block;
 * a voice that comes back without GAS_VOICE_ACTIVE should be dropped from the list (_manage_playback_state, :473-492), and an
 * instance whose voices are all gone stopped with gas_instance_stop.
 * Deliberate difference: in the block in which the LAST voice of a Mode-B instance is deactivated the reference delivers that
 * block (at or below the threshold) only on the pair whose proxy AudioServer happens to mix first, because playback_active is
 * cleared in the middle of the mix step (:484-491, :683-690); this library mixes it on every pair
 * (tests/test_lifecycle.py::test_reference_drops_the_last_block_of_the_other_pairs). */
GAS_API int gas_mix_block_stream(gas_ctx *ctx, int32_t n_voices, const gas_voice *voices, const gas_frame *src, int32_t src_rows,
		int32_t frames, const int32_t *mixed_frames, gas_frame *bus_out, int32_t *status_out);
GAS_API int gas_mix_block_stream_device(gas_ctx *ctx, int32_t n_voices, const gas_voice *d_voices, const gas_frame *d_src, int32_t src_rows,
		int32_t src_row_stride, int32_t frames, const int32_t *d_mixed_frames, gas_frame *d_bus_out, int32_t *d_status_out);
/* ---- the bus graph after the mix -------------------------------------------------------------------------
 * What upstream AudioServer::_mix_step does with the bus buffers after every playback has been mixed into them (reference
 * README.md:98-100 steps 3-5; the demo's default_bus_layout.tres sends its Reverb bus to Master): from the last bus to the first,
 * each bus is scaled by its volume — 0 when muted, or, if any bus is soloed, when it is neither soloed nor on a soloed bus's send
 * chain — and added to its send bus; bus 0 (Master) ends up holding what the audio driver gets.  Bus effects are not part of the
 * library (a caller with effects on a bus runs them between two calls); restated from Godot 4.x as recalled.
 *   gas_bus_layout_set    one descriptor per bus of the context (AudioBusLayout: volume_db, mute, solo, send as bus index; a send
 *                         that is not to a bus on the left goes to Master, like upstream)
 *   gas_bus_graph_device  d_bus [num_buses][channel_count][frames] in place, asynchronous on the mix stream; with sharded voices call
 *                         it on the reduced sums */
typedef struct gas_bus_desc {
	float volume_db;
	int32_t mute;
	int32_t solo;
	int32_t send;
} gas_bus_desc;
GAS_API int gas_bus_layout_set(gas_ctx *ctx, int32_t n_buses, const gas_bus_desc *buses);
GAS_API int gas_bus_graph_device(gas_ctx *ctx, gas_frame *d_bus, int32_t frames);
/* The same pass over host bus buffers (up, graph, down; synchronous). */
GAS_API int gas_bus_graph(gas_ctx *ctx, gas_frame *bus_inout, int32_t frames);
/* ---- device-resident sources + the resampler in front of the path --------------------------------------
 * What `playback->stream_playback->mix(&buf[LOOKAHEAD_BUFFER_SIZE], pitch_scale, p_buffer_size)` (audio_spatializer.cpp:375-378)
 * does for a resampled PCM stream, on the device: upstream AudioStreamPlaybackResampled::mix (16.16 fixed-point offset, 4-tap cubic
 * interpolation, 128-frame internal buffer; restated from Godot 4.x as recalled — the engine is not part of the reference tree) at
 * `sample_rate * pitch_scale / mix_rate`, pitch_scale being the instance's current SpatializerParameters::pitch_scale (Doppler
 * included).  With resident sources a block needs no frames from the host: only the emitters travel.
 *   gas_source_set     copies a PCM clip (AudioFrames at sample_rate, >= 128 frames; loop = wrap over the whole clip) into HBM as
 *                      source `slot` (0..4095); n_frames = 0 frees the slot.  Synchronous.
 *   gas_voice_play     AudioStreamPlaybackResampled::begin_resample for n voices: voice i plays sources[i] (-1 = none) from
 *                      start_frames[i] (NULL = 0) with a cleared interpolation history.  Call it next to gas_voice_init.
 *   gas_resample_block_device  row voices[j].src_row of d_rows receives `frames` new frames of voice j and d_mixed_frames[j] the
 *                      number of valid ones (< frames once the clip has ended, upstream's end rule): exactly the inputs of
 *                      gas_mix_block_stream_device.  Asynchronous on the mix stream.
 *   gas_mix_block_resident     resample + gas_mix_block_stream in one synchronous call with host voice descriptors: src_row of a
 *                      voice only names its row in the context's internal row buffer (use 0..n_voices-1). */
GAS_API int gas_source_set(gas_ctx *ctx, int32_t slot, const gas_frame *frames, int32_t n_frames, float sample_rate, int32_t loop);
GAS_API int gas_voice_play(gas_ctx *ctx, int32_t n, const int32_t *voices, const int32_t *sources, const int32_t *start_frames);
GAS_API int gas_resample_block_device(gas_ctx *ctx, int32_t n_voices, const gas_voice *d_voices, int32_t frames, gas_frame *d_rows,
		int32_t row_stride, int32_t src_rows, int32_t *d_mixed_frames);
GAS_API int gas_mix_block_resident(gas_ctx *ctx, int32_t n_voices, const gas_voice *voices, int32_t frames, gas_frame *bus_out,
		int32_t *status_out);
/* AudioSpatializerInstance::set_playback_disable_threshold_db (audio_spatializer.cpp:576-582; default -80, reset by
 * gas_instance_init). */
GAS_API int gas_set_playback_disable_threshold_db(gas_ctx *ctx, int32_t n, const int32_t *instances, const float *db);
GAS_API int gas_sync(gas_ctx *ctx);
/* cudaStream_t of the mix / gain streams as void*, for callers that enqueue their own work
 * (collectives, copies) in order with the mixer. */
GAS_API void *gas_mix_stream(gas_ctx *ctx);
GAS_API void *gas_gain_stream(gas_ctx *ctx);
/* Number of kernels this library has launched on ctx since creation (evidence for bench.py); graph
 * replays count the kernels of the replayed graph. */
GAS_API uint64_t gas_kernel_launches(const gas_ctx *ctx);

/* ---- CUDA-graph capture of the device-resident path ------------------------------------------------
 * Everything enqueued by gas_gain_compute_device (with resident listeners/areas) and
 * gas_mix_block_device between gas_capture_begin and gas_capture_end is recorded into a CUDA graph
 * instead of being executed; gas_graph_launch replays it on the mix stream.  One block is a handful of
 * ~10 us kernels, so a caller that renders many blocks back to back (offline render, benchmarks)
 * removes the per-launch host cost this way.  Other gas_* calls are not allowed while capturing. */
GAS_API int gas_capture_begin(gas_ctx *ctx);
GAS_API int gas_capture_end(gas_ctx *ctx, int32_t *out_graph);
GAS_API int gas_graph_launch(gas_ctx *ctx, int32_t graph);
GAS_API int gas_graph_destroy(gas_ctx *ctx, int32_t graph);

/* ---- per-kernel timing (CUDA events on the launching stream) ----------------------------------------
 * While enabled, every kernel launch is bracketed by timing events.  Inside a capture they become
 * event-record nodes of the graph: a graph captured with timing on is synchronised and read back after
 * every gas_graph_launch, so it reports each kernel's duration as it runs inside the replayed step.
 * gas_profile_read synchronises and returns, per kernel kind, the summed duration in milliseconds and
 * the number of launches since gas_profile_enable(ctx, 1). */
typedef enum gas_kernel_kind {
	GAS_KERNEL_PROLOGUE = 0, /* k_prologue_inst + k_prologue_voice */
	GAS_KERNEL_MIX_STREAM = 1, /* K2 */
	GAS_KERNEL_MIX_VOICE = 2,  /* K3 */
	GAS_KERNEL_GAIN = 3,       /* K1 (timed on the gain stream) */
	GAS_KERNEL_NONE = 4,       /* an event pair around nothing, recorded once per mix block right after the prologue's:
	                              what the timer itself reads inside the same graph / stream (to be subtracted by the reader) */
	GAS_KERNEL_KINDS = 5
} gas_kernel_kind;
GAS_API int gas_profile_enable(gas_ctx *ctx, int32_t on);
GAS_API int gas_profile_read(gas_ctx *ctx, double ms_out[GAS_KERNEL_KINDS], uint64_t launches_out[GAS_KERNEL_KINDS]);

/* ---- persistent state (checkpoint / re-sharding; no reference analogue) ------------- */
GAS_API int gas_voice_state_export(gas_ctx *ctx, int32_t n, const int32_t *voices, gas_voice_state *out);
GAS_API int gas_voice_state_import(gas_ctx *ctx, int32_t n, const int32_t *voices, const gas_voice_state *in);
GAS_API int gas_voice_life_export(gas_ctx *ctx, int32_t n, const int32_t *voices, gas_voice_life *out);
GAS_API int gas_voice_life_import(gas_ctx *ctx, int32_t n, const int32_t *voices, const gas_voice_life *in);
/* Sticky status bits since the last call (then cleared).  Bit 0: more distinct routing classes were in use than the plan has
 * dynamic slots; the voices concerned were mixed through the generic class of the voice-parallel kernel (correct, slower). */
#define GAS_STATUS_CLASS_OVERFLOW 1u
GAS_API int gas_status_flags(gas_ctx *ctx, uint32_t *out_flags);

/* ---- multi-GPU: voices sharded over ranks, partial bus buffers summed over peer memory ----------------
 * No reference analogue (one audio thread).  Every rank exports an IPC handle of its exchange allocation and
 * opens the others'; gas_reduce_bus_device then turns every rank's partial bus buffer into the sum over all
 * ranks: each rank adds its partial sums straight into every rank's exchange buffer with vector reductions on
 * peer pointers (NVLink / NVSwitch), one arrival counter round per block is the only synchronisation.
 * handle_out: 64 bytes (cudaIpcMemHandle_t).  All ranks must call gas_reduce_bus_device once per block, in the
 * same order; it is asynchronous on the mix stream and can be captured into a step graph. */
GAS_API int gas_comm_export(gas_ctx *ctx, void *handle_out, size_t handle_bytes);
GAS_API int gas_comm_open(gas_ctx *ctx, int32_t rank, int32_t n_ranks, const void *handles, size_t handle_bytes);
GAS_API int gas_reduce_bus_device(gas_ctx *ctx, gas_frame *d_bus, int32_t frames);
/* The two halves of gas_reduce_bus_device, for callers that overlap the other ranks' skew with their next block:
 * begin pushes this rank's partial sums of block n to every rank, end waits for every rank's push of the oldest
 * unfinished block and writes its complete sum to d_bus.  Per rank the order must be begin(n), end(n), begin(n + 1), ...:
 * begin(n + 1) zeroes the exchange buffer block n is still accumulating in, so it must follow end(n) (enforced:
 * GAS_ERR_STATE otherwise).  What overlaps the other ranks' skew is the caller's work between begin(n) and end(n); the
 * one-block-in-flight form is gas_reduce_bus_exchange_device. */
GAS_API int gas_reduce_bus_begin_device(gas_ctx *ctx, const gas_frame *d_bus, int32_t frames);
GAS_API int gas_reduce_bus_end_device(gas_ctx *ctx, gas_frame *d_bus, int32_t frames);
/* One block in flight, off the critical path: on the context's exchange stream (beside whatever the mix stream
 * does next) write the complete sum of the previously pushed block to d_prev_sum (skipped when no block is
 * outstanding; may be NULL then) and push d_partial as the next block.  d_partial must stay untouched until the
 * call after next has been enqueued (the mix stream waits for the exchange before it reuses bus buffers, and the exchange
 * waits for the mix that produced d_partial — as stream events outside a capture, as graph edges inside one; an exchange
 * captured without the mix of its block in the same graph relies on graph launches being ordered on the mix stream).
 * Drain with gas_reduce_bus_end_device. */
GAS_API int gas_reduce_bus_exchange_device(gas_ctx *ctx, const gas_frame *d_partial, gas_frame *d_prev_sum, int32_t frames);
GAS_API int gas_comm_close(gas_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* GAS_H */
