// audio_spatializer_gpu.h — C++ host mirror of the reference's spatializer API for the batched GPU path.
//
// Same class names, property names, defaults, validation and threading contract as the Godot module
// (reference audio_spatializer.h, audio_spatializer_3d.h, audio_spatializer_effect.h,
// spatializer_parameters.h), minus the engine: no Object/ClassDB/Ref machinery, a handful of POD math
// types instead of core/math.  Everything that touches a sample or a gain goes through the C ABI
// (include/gas.h) to the CUDA kernels; this layer only keeps the bookkeeping the reference keeps in its
// C++ objects (which instances exist, which playbacks are alive, who owns which slot) and turns the
// per-object virtual calls into one batched call per physics tick / mix step.
//
//   reference (per object, per tick)                              this layer (batched)
//   AudioStreamPlayerSpatial::_notification(PHYSICS_PROCESS)
//     -> AudioSpatializerInstance::update_spatializer_parameters   BatchMixer::update_spatializer_parameters()
//        -> calculate_spatialization()            [virtual]           built-in 3D math: gas_gain_compute (K1)
//                                                                     overridden:       gas_params_set
//   AudioServer::_mix_step -> AudioStreamPlaybackSpatial::mix
//     -> get_mixed_frames -> _mix_from_playback_list               BatchMixer::mix()  (gas_mix_block: prologue, K2, K3)
//        -> process_frames() / mix_channel()      [virtual]           built-in semantics on the device
//
// Error convention (reference: ERR_FAIL_* print and return, never throw): setters with invalid arguments
// leave the object untouched, record the message in last_error() and return false.
#pragma once

#include "../../include/gas.h"

#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

namespace gas {

// ---- godot-lite math / audio types ------------------------------------------------------------------
struct Vector2 {
	float x = 0.f, y = 0.f;
};
struct Vector3 {
	float x = 0.f, y = 0.f, z = 0.f;
};
struct Basis { // rows, like Godot's Basis
	float rows[3][3] = { { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 } };
	Vector3 get_column(int c) const { return Vector3{ rows[0][c], rows[1][c], rows[2][c] }; }
};
struct Transform3D {
	Basis basis;
	Vector3 origin;
};
using AudioFrame = gas_frame;
template <class T>
using Ref = std::shared_ptr<T>;

const std::string &last_error();
void set_last_error(const std::string &msg);

// ---- SpatializerParameters (reference spatializer_parameters.h:39-67) --------------------------------
class SpatializerParameters {
public:
	virtual ~SpatializerParameters() = default;
	// bus -> 4 x Vector2, Dictionary insertion order; size must be 4 (spatializer_parameters.cpp:35-38)
	bool add_bus_volume(int p_bus, const std::vector<Vector2> &p_volumes);
	const std::vector<std::pair<int, std::vector<Vector2>>> &get_bus_volumes() const { return bus_volumes; }
	bool set_mix_volumes(const std::vector<Vector2> &p_volumes); // size must be 4 (:44-47)
	const std::vector<Vector2> &get_mix_volumes() const { return mix_volumes; }
	void set_pitch_scale(float p) { pitch_scale = p; }
	float get_pitch_scale() const { return pitch_scale; }
	void set_update_parameters(bool p) { update_parameters = p; }
	bool should_update_parameters() const { return update_parameters; }
	virtual void to_pod(gas_params &out) const;
	virtual void from_pod(const gas_params &in);

protected:
	std::vector<std::pair<int, std::vector<Vector2>>> bus_volumes;
	std::vector<Vector2> mix_volumes = std::vector<Vector2>(GAS_MAX_CHANNELS_PER_BUS);
	float pitch_scale = 1.0f;
	bool update_parameters = false;
};

// SpatializerParameters3D (reference audio_spatializer_3d.h:61-83)
class SpatializerParameters3D : public SpatializerParameters {
public:
	void set_linear_attenuation(float v) { linear_attenuation = v; }
	float get_linear_attenuation() const { return linear_attenuation; }
	void set_attenuation_filter_cutoff_hz(float v) { attenuation_filter_cutoff_hz = v; }
	float get_attenuation_filter_cutoff_hz() const { return attenuation_filter_cutoff_hz; }
	void to_pod(gas_params &out) const override;
	void from_pod(const gas_params &in) override;

private:
	float linear_attenuation = 0.0f;
	float attenuation_filter_cutoff_hz = 5000.0f;
};

class BatchMixer;

// SpatializerPlaybackData / SpatializerPlaybackData3D (reference audio_spatializer.h, audio_spatializer_3d.h:85-99):
// the persistent per-playback state lives in HBM; the host object is the handle of its slot.
class SpatializerPlaybackData {
public:
	virtual ~SpatializerPlaybackData() = default;
	int voice_slot = -1;
	BatchMixer *mixer = nullptr;
};
class SpatializerPlaybackData3D : public SpatializerPlaybackData {
public:
	// reference audio_spatializer_3d.cpp:880-885: missing entries read as (0, 0); reads the device state
	Vector2 get_prev_mix_volume(int p_channel) const;
};

class AudioSpatializerInstance;

// ---- AudioSpatializer (reference audio_spatializer.h:153-163) -------------------------------------------------
class AudioSpatializer : public std::enable_shared_from_this<AudioSpatializer> {
public:
	virtual ~AudioSpatializer() = default;
	virtual Ref<AudioSpatializerInstance> instantiate() = 0;
	virtual void to_pod(gas_spatializer &out) const = 0;
	int slot = -1; // gas_spatializer slot once registered with a BatchMixer
	bool dirty = true;
};

// AudioSpatializer3D (reference audio_spatializer_3d.h:153-241, setters audio_spatializer_3d.cpp:654-765)
class AudioSpatializer3D : public AudioSpatializer {
public:
	enum AttenuationModel { ATTENUATION_INVERSE_DISTANCE, ATTENUATION_INVERSE_SQUARE_DISTANCE, ATTENUATION_LOGARITHMIC, ATTENUATION_DISABLED };
	enum DopplerTracking { DOPPLER_TRACKING_DISABLED, DOPPLER_TRACKING_IDLE_STEP, DOPPLER_TRACKING_PHYSICS_STEP };

	AudioSpatializer3D() { gas_spatializer_defaults(&p); }
	Ref<AudioSpatializerInstance> instantiate() override;
	void to_pod(gas_spatializer &out) const override { out = p; }

	void set_mix_channel_mode(bool v) { p.mix_channel_mode = v; dirty = true; }
	bool get_mix_channel_mode() const { return p.mix_channel_mode != 0; }
	void set_unit_size(float v) { p.unit_size = v; dirty = true; }
	float get_unit_size() const { return p.unit_size; }
	bool set_max_distance(float p_metres); // >= 0 (:671)
	float get_max_distance() const { return p.max_distance; }
	void set_area_mask(uint32_t v) { p.area_mask = v; dirty = true; }
	uint32_t get_area_mask() const { return p.area_mask; }
	void set_emission_angle_enabled(bool v) { p.emission_angle_enabled = v; dirty = true; }
	bool is_emission_angle_enabled() const { return p.emission_angle_enabled != 0; }
	bool set_emission_angle(float p_angle); // [0, 90] (:696)
	float get_emission_angle() const { return p.emission_angle; }
	void set_emission_angle_filter_attenuation_db(float v) { p.emission_angle_filter_attenuation_db = v; dirty = true; }
	float get_emission_angle_filter_attenuation_db() const { return p.emission_angle_filter_attenuation_db; }
	void set_attenuation_filter_cutoff_hz(float v) { p.attenuation_filter_cutoff_hz = v; dirty = true; }
	float get_attenuation_filter_cutoff_hz() const { return p.attenuation_filter_cutoff_hz; }
	void set_attenuation_filter_db(float v) { p.attenuation_filter_db = v; dirty = true; }
	float get_attenuation_filter_db() const { return p.attenuation_filter_db; }
	bool set_attenuation_model(int p_model); // index < 4 (:729)
	AttenuationModel get_attenuation_model() const { return (AttenuationModel)p.attenuation_model; }
	bool set_panning_strength(float v); // >= 0 (:738)
	float get_panning_strength() const { return p.panning_strength; }
	void set_doppler_tracking(DopplerTracking v) { p.doppler_tracking = (int)v; dirty = true; }
	DopplerTracking get_doppler_tracking() const { return (DopplerTracking)p.doppler_tracking; }
	bool set_doppler_speed_of_sound(float v); // > 0 (:759)
	float get_doppler_speed_of_sound() const { return p.doppler_speed_of_sound; }

protected:
	gas_spatializer p;
};

// AudioSpatializerEffect (reference audio_spatializer_effect.h:83-96), filter-chain subset: the effects are
// AudioEffectFilter-family biquads (what the example project uses, gd_spatializer.gd:11-20).
class AudioSpatializerEffect : public AudioSpatializer3D {
public:
	AudioSpatializerEffect() { p.kind = GAS_SPATIALIZER_EFFECT; p.mix_channel_mode = 0; }
	Ref<AudioSpatializerInstance> instantiate() override;
	bool add_effect(const gas_effect &e);       // at most GAS_MAX_EFFECTS
	bool set_effect(int index, const gas_effect &e);
	int get_effect_count() const { return p.chain.n_effects; }
	// effect whose gain follows SpatializerParameters3D::linear_attenuation every block, like the example's
	// _process_effects (gd_spatializer_instance.gd:125-127); -1 = none
	void set_effect_gain_binding(int index) { p.effect_gain_binding = index; dirty = true; }
};

// ---- AudioSpatializerInstance (reference audio_spatializer.h:41-151) -------------------------------------------------
class AudioSpatializerInstance {
public:
	enum { MAX_CHANNELS_PER_BUS = GAS_MAX_CHANNELS_PER_BUS, LOOKAHEAD_BUFFER_SIZE = GAS_LOOKAHEAD_BUFFER_SIZE, MAX_BUSES_PER_PLAYBACK = GAS_MAX_BUSES_PER_PLAYBACK };
	virtual ~AudioSpatializerInstance() = default;

	// --- the plugin virtuals (audio_spatializer.h:144-150) ---
	// Return nullptr to have the parameters computed by the built-in batched kernel; return an object to hand
	// your own SpatializerParameters over (what a custom _calculate_spatialization does).  Physics thread.
	virtual Ref<SpatializerParameters> calculate_spatialization() { return nullptr; }
	virtual bool should_process_frames() const { return true; }
	virtual bool should_mix_channels() const { return false; }
	virtual Ref<SpatializerPlaybackData> instantiate_playback_data() { return std::make_shared<SpatializerPlaybackData>(); }
	virtual void initialize_audio_player() {}
	// The per-call virtuals with the reference's signature (audio_spatializer.h:146,148; raw-pointer GDVIRTUAL mirror
	// :103-112).  The base implementations ARE the built-in behaviour (AudioSpatializerInstance3D::process_frames /
	// mix_channel, AudioSpatializerInstanceEffect::process_frames), run on the device for this one playback through
	// gas_process_frames / gas_mix_channel: p_parameters must be the instance's current parameters (nullptr = current)
	// and p_playback_data one of its playbacks.  The batched mix does the same work for all voices at once and does NOT
	// call them, so a subclass that overrides one must say so with uses_builtin_dsp() = false; BatchMixer then refuses to
	// batch that instance (mix() fails with a message naming it) and the caller mixes it with its own loop, where the
	// override may still call the base implementation for the built-in part.
	virtual void process_frames(const Ref<SpatializerParameters> &p_parameters, const Ref<SpatializerPlaybackData> &p_playback_data,
			AudioFrame *p_output_buf, const AudioFrame *p_source_buf, int p_frame_count);
	virtual void mix_channel(const Ref<SpatializerParameters> &p_parameters, const Ref<SpatializerPlaybackData> &p_playback_data, int p_channel,
			AudioFrame *p_output_buf, const AudioFrame *p_source_buf, int p_frame_count);
	virtual bool uses_builtin_dsp() const { return true; }

	// --- what the reference reads from get_audio_player() / the scene (audio_stream_player_spatial.h:60,101-105) ---
	void set_global_transform(const Transform3D &t) { transform = t; }
	void set_linear_velocity(const Vector3 &v) { velocity = v; }
	bool set_volume_db(float db); // NaN rejected (audio_stream_player_spatial.cpp:193)
	void set_max_db(float db) { max_db = db; }
	void set_pitch_scale(float p) { pitch_scale = p; }
	void set_bus(int bus) { bus_index = bus; }
	void set_area(const gas_area *a); // result of the Area3D query; nullptr = none

	// --- playback lifecycle (audio_spatializer.h:121-126) ---
	Ref<SpatializerPlaybackData> start_playback_stream(); // returns the playback's data handle
	void stop_playback_stream(const Ref<SpatializerPlaybackData> &p);
	bool is_playback_active() const { return !playbacks.empty(); }
	float get_playback_disable_threshold_db() const { return playback_disable_threshold_db; }
	void set_playback_disable_threshold_db(float v); // audio_spatializer.cpp:580-582; reaches the device's lifecycle table

	Ref<SpatializerParameters> get_spatializer_parameters() const; // last parameters handed to the mixer

	int slot = -1;
	BatchMixer *mixer = nullptr;
	Ref<AudioSpatializer> base;
	std::vector<Ref<SpatializerPlaybackData>> playbacks;

protected:
	friend class BatchMixer;
	Transform3D transform;
	Vector3 velocity;
	float volume_db = 0.f, max_db = 3.f, pitch_scale = 1.f;
	int bus_index = 0;
	bool has_area = false;
	gas_area area{};
	float playback_disable_threshold_db = -80.0f;
};

class AudioSpatializerInstance3D : public AudioSpatializerInstance {
public:
	bool should_process_frames() const override { return !mix_channel_mode; }
	bool should_mix_channels() const override { return mix_channel_mode; }
	Ref<SpatializerPlaybackData> instantiate_playback_data() override { return std::make_shared<SpatializerPlaybackData3D>(); }
	bool mix_channel_mode = false; // latched from the resource at instantiate() (audio_spatializer_3d.cpp:649)
};

class AudioSpatializerInstanceEffect : public AudioSpatializerInstance3D {
public:
	// what a _process_effects override would write into the effects for the next blocks (audio_spatializer_effect.cpp:39)
	bool set_effect_parameters(const gas_effect_chain &chain);
};

// ---- BatchMixer: the AudioServer-side owner of the device context --------------------------------------------------------
struct BatchMixerConfig {
	int device = 0;
	int max_instances = 1024, max_voices = 1024, max_frames = 512, max_spatializers = 16, num_buses = 2;
	int speaker_mode = GAS_SPEAKER_MODE_STEREO;
	float mix_rate = 44100.f, global_panning_strength = 0.5f;
};

class BatchMixer {
public:
	explicit BatchMixer(const BatchMixerConfig &cfg);
	~BatchMixer();
	bool ok() const { return ctx != nullptr; }
	gas_ctx *context() { return ctx; }
	int get_channel_count() const { return ctx ? gas_get_channel_count(ctx) : 0; }

	// AudioStreamPlayerSpatial::set_spatializer -> AudioSpatializer::instantiate (audio_spatializer_3d.cpp:645-652)
	Ref<AudioSpatializerInstance> instantiate(const Ref<AudioSpatializer> &spatializer);
	void free_instance(const Ref<AudioSpatializerInstance> &inst);

	void set_listeners(const std::vector<gas_listener> &l) { std::lock_guard<std::mutex> lk(mu); listeners = l; }

	// Physics thread: update_spatializer_parameters of every instance (audio_spatializer.cpp:258-272), batched.
	bool update_spatializer_parameters();

	// Audio thread: one AudioServer mix step for every live playback.  sources[k] points at the F frames of
	// playback k in playback_order() (nullptr = silent tail).  bus_out: [num_buses][channels][frames].
	// peaks (optional): one AudioFrame per playback in playback_order().
	bool mix(int frames, const std::vector<const AudioFrame *> &sources, AudioFrame *bus_out, AudioFrame *peaks = nullptr);
	// The same step with the voice lifecycle of _mix_from_playback_list on the device (audio_spatializer.cpp:353-408,
	// :464-492): sources[k] points at the counts[k] <= frames frames AudioStreamPlayback::mix returned for playback k this
	// block (not spliced behind the lookahead; nullptr / 0 once the stream has ended).  Lookahead, end-of-stream fade,
	// silent tails and the deactivation below playback_disable_threshold_db happen in gas_mix_block_stream; playbacks that
	// come back inactive are dropped (_manage_playback_state) and an instance whose last playback is gone is stopped.
	// finished (optional) receives the playbacks dropped by this step.
	bool mix_streams(int frames, const std::vector<const AudioFrame *> &sources, const std::vector<int> &counts, AudioFrame *bus_out,
			std::vector<Ref<SpatializerPlaybackData>> *finished = nullptr);
	// Device-resident PCM streams (gas_source_set / gas_voice_play / gas_mix_block_resident): an AudioStreamWAV-like clip is uploaded
	// once; a playback started on it is resampled on the device (AudioStreamPlaybackResampled::mix at sample_rate x pitch_scale /
	// mix_rate) and goes through the same lifecycle as mix_streams — no frames travel per block.  Every live playback must have
	// been started with play_source before mix_resident is called.
	bool set_source(int slot, const AudioFrame *frames, int n_frames, float sample_rate, bool loop);
	bool play_source(const Ref<SpatializerPlaybackData> &playback, int source_slot, int from_frame = 0);
	bool mix_resident(int frames, AudioFrame *bus_out, std::vector<Ref<SpatializerPlaybackData>> *finished = nullptr);
	// AudioBusLayout: volume / mute / solo / send per bus (upstream AudioServer::_mix_step after the playbacks, README.md:98-100);
	// apply_bus_graph runs that pass over host bus buffers [num_buses][channels][frames] in place.
	bool set_bus_layout(const std::vector<gas_bus_desc> &layout);
	bool apply_bus_graph(int frames, AudioFrame *bus_inout);
	// the playbacks a mix() call expects sources for, in order (instances by slot, playbacks by start order)
	std::vector<Ref<SpatializerPlaybackData>> playback_order() const;

	bool voice_state(int voice_slot, gas_voice_state &out);

private:
	friend class AudioSpatializerInstance;
	int alloc_voice();
	void sync_spatializer(const Ref<AudioSpatializer> &s);
	gas_ctx *ctx = nullptr;
	BatchMixerConfig cfg;
	mutable std::mutex mu; // the reference's parameter hand-off mutex (audio_spatializer.cpp:558-574)
	std::vector<Ref<AudioSpatializerInstance>> instances; // by slot
	std::vector<Ref<AudioSpatializer>> spatializers;       // by slot
	std::vector<int> free_voices;
	int next_voice = 0;
	std::vector<int> started; // instances whose first playback has been registered (gas_instance_start)
	std::vector<gas_listener> listeners;
	std::vector<gas_frame> staging;
	bool refuse_custom_dsp() const; // sets last_error and returns true if an instance with uses_builtin_dsp() == false has playbacks
	std::map<int, Ref<SpatializerParameters>> last_params;
};

} // namespace gas
