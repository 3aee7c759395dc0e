// audio_spatializer_gpu.cpp — see audio_spatializer_gpu.h.  Host bookkeeping only: every gain and every sample is
// computed by the CUDA kernels behind the C ABI; there is no CPU implementation of the path in this file.
#include "audio_spatializer_gpu.h"

#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>

namespace gas {

static thread_local std::string g_err;
const std::string &last_error() { return g_err; }
void set_last_error(const std::string &msg) {
	g_err = msg;
	fprintf(stderr, "ERROR: %s\n", msg.c_str()); // ERR_FAIL_* prints and returns
}
#define GAS_FAIL_COND_V(cond, msg, ret) \
	do {                                \
		if (cond) {                     \
			set_last_error(msg);        \
			return ret;                 \
		}                               \
	} while (0)

// ---- SpatializerParameters ---------------------------------------------------------------------------------
bool SpatializerParameters::add_bus_volume(int p_bus, const std::vector<Vector2> &p_volumes) {
	GAS_FAIL_COND_V(p_volumes.size() != GAS_MAX_CHANNELS_PER_BUS, "Condition \"p_volumes.size() != 4\" is true.", false); // spatializer_parameters.cpp:36
	for (auto &e : bus_volumes) {
		if (e.first == p_bus) {
			e.second = p_volumes; // Dictionary assignment keeps the key's position
			return true;
		}
	}
	bus_volumes.emplace_back(p_bus, p_volumes);
	return true;
}

bool SpatializerParameters::set_mix_volumes(const std::vector<Vector2> &p_volumes) {
	GAS_FAIL_COND_V(p_volumes.size() != GAS_MAX_CHANNELS_PER_BUS, "Condition \"p_volumes.size() != 4\" is true.", false); // :45
	mix_volumes = p_volumes;
	return true;
}

void SpatializerParameters::to_pod(gas_params &out) const {
	memset(&out, 0, sizeof(out));
	for (int c = 0; c < GAS_MAX_CHANNELS_PER_BUS; c++) {
		out.mix_volumes[c][0] = mix_volumes[c].x;
		out.mix_volumes[c][1] = mix_volumes[c].y;
	}
	out.pitch_scale = pitch_scale;
	out.attenuation_filter_cutoff_hz = 5000.0f;
	out.update_parameters = update_parameters ? 1 : 0;
	out.n_bus = (int)std::min<size_t>(bus_volumes.size(), GAS_MAX_BUSES_PER_PLAYBACK); // audio_spatializer.cpp:284-322 keeps 6
	for (int k = 0; k < out.n_bus; k++) {
		out.bus[k] = bus_volumes[k].first;
		for (int c = 0; c < GAS_MAX_CHANNELS_PER_BUS; c++) {
			out.bus_volumes[k][c][0] = bus_volumes[k].second[c].x;
			out.bus_volumes[k][c][1] = bus_volumes[k].second[c].y;
		}
	}
}

void SpatializerParameters::from_pod(const gas_params &in) {
	bus_volumes.clear();
	for (int c = 0; c < GAS_MAX_CHANNELS_PER_BUS; c++) {
		mix_volumes[c] = Vector2{ in.mix_volumes[c][0], in.mix_volumes[c][1] };
	}
	pitch_scale = in.pitch_scale;
	update_parameters = in.update_parameters != 0;
	for (int k = 0; k < in.n_bus && k < GAS_MAX_BUSES_PER_PLAYBACK; k++) {
		std::vector<Vector2> v(GAS_MAX_CHANNELS_PER_BUS);
		for (int c = 0; c < GAS_MAX_CHANNELS_PER_BUS; c++) {
			v[c] = Vector2{ in.bus_volumes[k][c][0], in.bus_volumes[k][c][1] };
		}
		bus_volumes.emplace_back(in.bus[k], v);
	}
}

void SpatializerParameters3D::to_pod(gas_params &out) const {
	SpatializerParameters::to_pod(out);
	out.linear_attenuation = linear_attenuation;
	out.attenuation_filter_cutoff_hz = attenuation_filter_cutoff_hz;
}
void SpatializerParameters3D::from_pod(const gas_params &in) {
	SpatializerParameters::from_pod(in);
	linear_attenuation = in.linear_attenuation;
	attenuation_filter_cutoff_hz = in.attenuation_filter_cutoff_hz;
}

Vector2 SpatializerPlaybackData3D::get_prev_mix_volume(int p_channel) const {
	gas_voice_state st;
	if (!mixer || voice_slot < 0 || p_channel < 0 || p_channel >= GAS_MAX_CHANNELS_PER_BUS || !mixer->voice_state(voice_slot, st)) {
		return Vector2{};
	}
	return Vector2{ st.prev_mix_volumes[p_channel][0], st.prev_mix_volumes[p_channel][1] };
}

// ---- AudioSpatializer3D setters with the reference's validation ---------------------------------------------------------
bool AudioSpatializer3D::set_max_distance(float p_metres) {
	GAS_FAIL_COND_V(!(p_metres >= 0.0f), "Condition \"p_metres < 0.0\" is true.", false);
	p.max_distance = p_metres;
	dirty = true;
	return true;
}
bool AudioSpatializer3D::set_emission_angle(float p_angle) {
	GAS_FAIL_COND_V(!(p_angle >= 0.f && p_angle <= 90.f), "Condition \"p_angle < 0 || p_angle > 90\" is true.", false);
	p.emission_angle = p_angle;
	dirty = true;
	return true;
}
bool AudioSpatializer3D::set_attenuation_model(int p_model) {
	GAS_FAIL_COND_V(p_model < 0 || p_model >= 4, "Index (int)p_model is out of bounds (4).", false);
	p.attenuation_model = p_model;
	dirty = true;
	return true;
}
bool AudioSpatializer3D::set_panning_strength(float v) {
	GAS_FAIL_COND_V(!(v >= 0.f), "Panning strength must be a positive number.", false);
	p.panning_strength = v;
	dirty = true;
	return true;
}
bool AudioSpatializer3D::set_doppler_speed_of_sound(float v) {
	GAS_FAIL_COND_V(!(v > 0.f), "Speed of sound must be a positive number.", false);
	p.doppler_speed_of_sound = v;
	dirty = true;
	return true;
}
Ref<AudioSpatializerInstance> AudioSpatializer3D::instantiate() {
	auto ins = std::make_shared<AudioSpatializerInstance3D>();
	ins->base = shared_from_this();
	ins->mix_channel_mode = p.mix_channel_mode != 0; // audio_spatializer_3d.cpp:649
	return ins;
}

bool AudioSpatializerEffect::add_effect(const gas_effect &e) {
	GAS_FAIL_COND_V(p.chain.n_effects >= GAS_MAX_EFFECTS, "too many effects for the batched filter chain", false);
	p.chain.effects[p.chain.n_effects++] = e;
	dirty = true;
	return true;
}
bool AudioSpatializerEffect::set_effect(int index, const gas_effect &e) {
	GAS_FAIL_COND_V(index < 0 || index >= p.chain.n_effects, "Index p_index is out of bounds (effects.size()).", false);
	p.chain.effects[index] = e;
	dirty = true;
	return true;
}
Ref<AudioSpatializerInstance> AudioSpatializerEffect::instantiate() {
	auto ins = std::make_shared<AudioSpatializerInstanceEffect>();
	ins->base = shared_from_this();
	ins->mix_channel_mode = false; // the effect spatializer only implements process_frames (audio_spatializer_effect.h:39-62)
	return ins;
}

// ---- AudioSpatializerInstance -------------------------------------------------------------------------------------------
bool AudioSpatializerInstance::set_volume_db(float db) {
	GAS_FAIL_COND_V(isnan(db), "Volume can't be set to NaN.", false); // audio_stream_player_spatial.cpp:193
	volume_db = db;
	return true;
}
void AudioSpatializerInstance::set_area(const gas_area *a) {
	has_area = a != nullptr;
	if (a) {
		area = *a;
	}
}
Ref<SpatializerPlaybackData> AudioSpatializerInstance::start_playback_stream() {
	GAS_FAIL_COND_V(!mixer, "instance is not registered with a BatchMixer", nullptr);
	Ref<SpatializerPlaybackData> d = instantiate_playback_data();
	d->mixer = mixer;
	d->voice_slot = mixer->alloc_voice();
	GAS_FAIL_COND_V(d->voice_slot < 0, "out of voice slots", nullptr);
	std::lock_guard<std::mutex> lk(mixer->mu);
	int32_t v = d->voice_slot;
	if (gas_voice_init(mixer->ctx, 1, &v) != GAS_OK) {
		set_last_error(gas_last_error(mixer->ctx));
		return nullptr;
	}
	playbacks.push_back(d);
	return d;
}
void AudioSpatializerInstance::stop_playback_stream(const Ref<SpatializerPlaybackData> &p) {
	if (!mixer) {
		return;
	}
	std::lock_guard<std::mutex> lk(mixer->mu);
	auto it = std::find(playbacks.begin(), playbacks.end(), p);
	if (it == playbacks.end()) {
		return;
	}
	mixer->free_voices.push_back((*it)->voice_slot);
	playbacks.erase(it);
	if (playbacks.empty()) { // _manage_playback_state stops the proxies (audio_spatializer.cpp:484-491)
		int32_t q = slot;
		gas_instance_stop(mixer->ctx, 1, &q);
		mixer->started[slot] = 0;
	}
}
void AudioSpatializerInstance::set_playback_disable_threshold_db(float v) {
	playback_disable_threshold_db = v;
	if (mixer && slot >= 0) {
		int32_t q = slot;
		if (gas_set_playback_disable_threshold_db(mixer->context(), 1, &q, &v) != GAS_OK) {
			set_last_error(gas_last_error(mixer->context()));
		}
	}
}

// reference ERR_FAIL_COND_MSG(!Object::cast_to<...>) at the top of the virtuals (audio_spatializer_3d.cpp:493-494, :556-557):
// wrong or foreign playback data => the call is a no-op with an error
static bool per_call_args_ok(const AudioSpatializerInstance *self, const Ref<SpatializerPlaybackData> &pd, const char *what) {
	if (!self->mixer || self->slot < 0) {
		set_last_error(std::string(what) + ": instance is not registered with a BatchMixer");
		return false;
	}
	if (!pd || pd->mixer != self->mixer || pd->voice_slot < 0 ||
			std::find(self->playbacks.begin(), self->playbacks.end(), pd) == self->playbacks.end()) {
		set_last_error(std::string(what) + ": Unexpected SpatializerPlaybackData; expected one of this instance's playbacks");
		return false;
	}
	return true;
}
void AudioSpatializerInstance::process_frames(const Ref<SpatializerParameters> &, const Ref<SpatializerPlaybackData> &p_playback_data,
		AudioFrame *p_output_buf, const AudioFrame *p_source_buf, int p_frame_count) {
	if (!per_call_args_ok(this, p_playback_data, "process_frames")) {
		return;
	}
	if (gas_process_frames(mixer->context(), slot, p_playback_data->voice_slot, p_output_buf, p_source_buf, p_frame_count) != GAS_OK) {
		set_last_error(gas_last_error(mixer->context()));
	}
}
void AudioSpatializerInstance::mix_channel(const Ref<SpatializerParameters> &, const Ref<SpatializerPlaybackData> &p_playback_data, int p_channel,
		AudioFrame *p_output_buf, const AudioFrame *p_source_buf, int p_frame_count) {
	if (!per_call_args_ok(this, p_playback_data, "mix_channel")) {
		return;
	}
	if (gas_mix_channel(mixer->context(), slot, p_playback_data->voice_slot, p_channel, p_output_buf, p_source_buf, p_frame_count) != GAS_OK) {
		set_last_error(gas_last_error(mixer->context()));
	}
}

Ref<SpatializerParameters> AudioSpatializerInstance::get_spatializer_parameters() const {
	if (!mixer) {
		return nullptr;
	}
	gas_params pod;
	int32_t q = slot;
	if (gas_params_get(mixer->ctx, 1, &q, &pod) != GAS_OK) {
		return nullptr;
	}
	auto p = std::make_shared<SpatializerParameters3D>();
	p->from_pod(pod);
	return p;
}
bool AudioSpatializerInstanceEffect::set_effect_parameters(const gas_effect_chain &chain) {
	GAS_FAIL_COND_V(!mixer, "instance is not registered with a BatchMixer", false);
	int32_t q = slot;
	if (gas_effect_params_set(mixer->context(), 1, &q, &chain) != GAS_OK) {
		set_last_error(gas_last_error(mixer->context()));
		return false;
	}
	return true;
}

// ---- BatchMixer -------------------------------------------------------------------------------------------------------------
BatchMixer::BatchMixer(const BatchMixerConfig &c) : cfg(c) {
	gas_config gc;
	gas_config_defaults(&gc);
	gc.device = c.device;
	gc.max_instances = c.max_instances;
	gc.max_voices = c.max_voices;
	gc.max_frames = c.max_frames;
	gc.max_spatializers = c.max_spatializers;
	gc.num_buses = c.num_buses;
	gc.speaker_mode = c.speaker_mode;
	gc.mix_rate = c.mix_rate;
	gc.global_panning_strength = c.global_panning_strength;
	if (gas_create(&gc, &ctx) != GAS_OK) {
		set_last_error(gas_last_error(nullptr)); // no CUDA device => no mixer: there is no CPU fallback
		ctx = nullptr;
	}
	instances.resize(c.max_instances);
	spatializers.resize(c.max_spatializers);
	started.assign(c.max_instances, 0);
}
BatchMixer::~BatchMixer() {
	if (ctx) {
		gas_destroy(ctx);
	}
}
int BatchMixer::alloc_voice() {
	std::lock_guard<std::mutex> lk(mu);
	if (!free_voices.empty()) {
		int v = free_voices.back();
		free_voices.pop_back();
		return v;
	}
	return next_voice < cfg.max_voices ? next_voice++ : -1;
}
void BatchMixer::sync_spatializer(const Ref<AudioSpatializer> &s) {
	if (s->slot < 0) {
		for (int i = 0; i < (int)spatializers.size(); i++) {
			if (!spatializers[i]) {
				spatializers[i] = s;
				s->slot = i;
				s->dirty = true;
				break;
			}
		}
	}
	if (s->slot >= 0 && s->dirty) {
		gas_spatializer pod;
		s->to_pod(pod);
		if (gas_spatializer_set(ctx, s->slot, &pod) != GAS_OK) {
			set_last_error(gas_last_error(ctx));
		}
		s->dirty = false;
	}
}
Ref<AudioSpatializerInstance> BatchMixer::instantiate(const Ref<AudioSpatializer> &spatializer) {
	GAS_FAIL_COND_V(!ctx, "no device context", nullptr);
	GAS_FAIL_COND_V(!spatializer, "Parameter \"spatializer\" is null.", nullptr);
	std::lock_guard<std::mutex> lk(mu);
	sync_spatializer(spatializer);
	GAS_FAIL_COND_V(spatializer->slot < 0, "out of spatializer slots", nullptr);
	int q = -1;
	for (int i = 0; i < (int)instances.size(); i++) {
		if (!instances[i]) {
			q = i;
			break;
		}
	}
	GAS_FAIL_COND_V(q < 0, "out of instance slots", nullptr);
	Ref<AudioSpatializerInstance> ins = spatializer->instantiate();
	ins->slot = q;
	ins->mixer = this;
	int32_t qi = q, si = spatializer->slot;
	if (gas_instance_init(ctx, 1, &qi, &si) != GAS_OK) {
		set_last_error(gas_last_error(ctx));
		return nullptr;
	}
	instances[q] = ins;
	ins->initialize_audio_player();
	return ins;
}
void BatchMixer::free_instance(const Ref<AudioSpatializerInstance> &inst) {
	if (!inst || inst->mixer != this) {
		return;
	}
	while (!inst->playbacks.empty()) {
		inst->stop_playback_stream(inst->playbacks.back());
	}
	std::lock_guard<std::mutex> lk(mu);
	instances[inst->slot].reset();
	inst->mixer = nullptr;
}

bool BatchMixer::update_spatializer_parameters() {
	GAS_FAIL_COND_V(!ctx, "no device context", false);
	std::lock_guard<std::mutex> lk(mu);
	std::vector<gas_emitter> em;
	std::vector<gas_area> areas;
	std::vector<int32_t> custom_ids;
	std::vector<gas_params> custom;
	std::vector<int32_t> to_start;
	for (auto &ins : instances) {
		if (!ins) {
			continue;
		}
		sync_spatializer(ins->base); // property edits reach the device before the next tick
		Ref<SpatializerParameters> own = ins->calculate_spatialization();
		if (own) { // a custom _calculate_spatialization: hand its parameters over as they are
			gas_params pod;
			own->to_pod(pod);
			custom_ids.push_back(ins->slot);
			custom.push_back(pod);
		} else {
			gas_emitter e;
			memset(&e, 0, sizeof(e));
			e.instance = ins->slot;
			e.spatializer = ins->base->slot;
			e.area = -1;
			if (ins->has_area) {
				e.area = (int)areas.size();
				areas.push_back(ins->area);
			}
			e.bus = ins->bus_index;
			const Vector3 z = ins->transform.basis.get_column(2);
			e.origin[0] = ins->transform.origin.x, e.origin[1] = ins->transform.origin.y, e.origin[2] = ins->transform.origin.z;
			e.basis_z[0] = z.x, e.basis_z[1] = z.y, e.basis_z[2] = z.z;
			e.velocity[0] = ins->velocity.x, e.velocity[1] = ins->velocity.y, e.velocity[2] = ins->velocity.z;
			e.volume_db = ins->volume_db;
			e.max_db = ins->max_db;
			e.pitch_scale = ins->pitch_scale;
			em.push_back(e);
		}
		if (!ins->playbacks.empty() && !started[ins->slot]) {
			to_start.push_back(ins->slot);
		}
	}
	if (!em.empty() && gas_gain_compute(ctx, (int)em.size(), em.data(), (int)listeners.size(), listeners.data(), (int)areas.size(),
							   areas.empty() ? nullptr : areas.data(), nullptr) != GAS_OK) {
		set_last_error(gas_last_error(ctx));
		return false;
	}
	if (!custom.empty() && gas_params_set(ctx, (int)custom.size(), custom_ids.data(), custom.data()) != GAS_OK) {
		set_last_error(gas_last_error(ctx));
		return false;
	}
	// first playback of an instance: its proxies are registered with the current parameters (audio_spatializer.cpp:75-95)
	if (!to_start.empty()) {
		if (gas_instance_start(ctx, (int)to_start.size(), to_start.data()) != GAS_OK) {
			set_last_error(gas_last_error(ctx));
			return false;
		}
		for (int q : to_start) {
			started[q] = 1;
		}
	}
	return true;
}

std::vector<Ref<SpatializerPlaybackData>> BatchMixer::playback_order() const {
	std::lock_guard<std::mutex> lk(mu);
	std::vector<Ref<SpatializerPlaybackData>> out;
	for (auto &ins : instances) {
		if (ins) {
			out.insert(out.end(), ins->playbacks.begin(), ins->playbacks.end());
		}
	}
	return out;
}

bool BatchMixer::refuse_custom_dsp() const {
	for (auto &ins : instances) {
		if (ins && !ins->playbacks.empty() && !ins->uses_builtin_dsp()) {
			set_last_error("instance " + std::to_string(ins->slot) + " overrides process_frames / mix_channel (uses_builtin_dsp() == false): the batched mix "
					"cannot run a per-voice override; mix this instance's playbacks with your own loop (the base-class virtuals run the built-in "
					"behaviour per call) and free it from the BatchMixer step");
			return true;
		}
	}
	return false;
}

bool BatchMixer::mix_streams(int frames, const std::vector<const AudioFrame *> &sources, const std::vector<int> &counts, AudioFrame *bus_out,
		std::vector<Ref<SpatializerPlaybackData>> *finished) {
	GAS_FAIL_COND_V(!ctx, "no device context", false);
	std::vector<std::pair<Ref<AudioSpatializerInstance>, Ref<SpatializerPlaybackData>>> order;
	std::vector<gas_voice> voices;
	std::vector<int32_t> mixed;
	{
		std::lock_guard<std::mutex> lk(mu);
		if (refuse_custom_dsp()) {
			return false;
		}
		size_t k = 0;
		for (auto &ins : instances) {
			if (!ins) {
				continue;
			}
			for (auto &pb : ins->playbacks) {
				gas_voice v;
				v.voice = pb->voice_slot;
				v.instance = ins->slot;
				const bool has = k < sources.size() && sources[k] && k < counts.size() && counts[k] > 0;
				v.src_row = has ? (int)k : -1;
				v.flags = 0u;
				voices.push_back(v);
				mixed.push_back(has ? counts[k] : 0);
				order.emplace_back(ins, pb);
				k++;
			}
		}
		GAS_FAIL_COND_V(sources.size() != voices.size() || counts.size() != voices.size(), "one source pointer and one frame count per live playback (see playback_order())", false);
		GAS_FAIL_COND_V(frames <= 0 || frames > cfg.max_frames || (frames & 1), "Condition \"p_frame_count != mix_buffer[ch].size()\" is true.", false);
		staging.resize(voices.size() * (size_t)frames);
		for (size_t i = 0; i < voices.size(); i++) {
			GAS_FAIL_COND_V(mixed[i] > frames, "a playback cannot return more frames than were asked for", false);
			memset(&staging[i * frames], 0, (size_t)frames * sizeof(AudioFrame));
			if (voices[i].src_row >= 0) {
				memcpy(&staging[i * frames], sources[i], (size_t)mixed[i] * sizeof(AudioFrame));
			}
		}
		std::vector<int32_t> status(voices.size() ? voices.size() : 1);
		if (gas_mix_block_stream(ctx, (int)voices.size(), voices.data(), staging.data(), (int)voices.size(), frames, mixed.data(), bus_out,
					status.data()) != GAS_OK) {
			set_last_error(gas_last_error(ctx));
			return false;
		}
		// _manage_playback_state (audio_spatializer.cpp:473-492): inactive playbacks leave the list
		for (size_t i = 0; i < order.size(); i++) {
			if (!(status[i] & GAS_VOICE_ACTIVE)) {
				if (finished) {
					finished->push_back(order[i].second);
				}
			} else {
				order[i].second.reset();
			}
		}
	}
	for (auto &e : order) {
		if (e.second) {
			e.first->stop_playback_stream(e.second); // frees the slot; stops the instance's proxies when it was the last one
		}
	}
	return true;
}

bool BatchMixer::mix(int frames, const std::vector<const AudioFrame *> &sources, AudioFrame *bus_out, AudioFrame *peaks) {
	GAS_FAIL_COND_V(!ctx, "no device context", false);
	std::lock_guard<std::mutex> lk(mu);
	if (refuse_custom_dsp()) {
		return false;
	}
	std::vector<gas_voice> voices;
	size_t k = 0;
	for (auto &ins : instances) {
		if (!ins) {
			continue;
		}
		for (auto &pb : ins->playbacks) {
			gas_voice v;
			v.voice = pb->voice_slot;
			v.instance = ins->slot;
			v.src_row = (k < sources.size() && sources[k]) ? (int)k : -1;
			v.flags = peaks ? GAS_VOICE_WANT_PEAK : 0u;
			voices.push_back(v);
			k++;
		}
	}
	GAS_FAIL_COND_V(sources.size() != voices.size(), "one source pointer per live playback (see playback_order())", false);
	// the reference errors out when the frame count changes under it (audio_spatializer.cpp:336-338, :521-522)
	GAS_FAIL_COND_V(frames <= 0 || frames > cfg.max_frames || (frames & 1), "Condition \"p_frame_count != mix_buffer[ch].size()\" is true.", false);
	staging.resize(voices.size() * (size_t)frames);
	for (size_t i = 0; i < voices.size(); i++) {
		if (sources[i]) {
			memcpy(&staging[i * frames], sources[i], (size_t)frames * sizeof(AudioFrame));
		} else {
			memset(&staging[i * frames], 0, (size_t)frames * sizeof(AudioFrame));
		}
	}
	if (gas_mix_block(ctx, (int)voices.size(), voices.data(), staging.data(), (int)voices.size(), frames, bus_out, peaks) != GAS_OK) {
		set_last_error(gas_last_error(ctx));
		return false;
	}
	return true;
}

bool BatchMixer::set_source(int slot, const AudioFrame *frames, int n_frames, float sample_rate, bool loop) {
	GAS_FAIL_COND_V(!ctx, "no device context", false);
	if (gas_source_set(ctx, slot, reinterpret_cast<const gas_frame *>(frames), n_frames, sample_rate, loop ? 1 : 0) != GAS_OK) {
		set_last_error(gas_last_error(ctx));
		return false;
	}
	return true;
}

bool BatchMixer::play_source(const Ref<SpatializerPlaybackData> &playback, int source_slot, int from_frame) {
	GAS_FAIL_COND_V(!ctx, "no device context", false);
	GAS_FAIL_COND_V(!playback || playback->voice_slot < 0, "Parameter \"p_playback\" is null.", false);
	const int32_t v = playback->voice_slot, s = source_slot, f = from_frame;
	if (gas_voice_play(ctx, 1, &v, &s, &f) != GAS_OK) {
		set_last_error(gas_last_error(ctx));
		return false;
	}
	return true;
}

bool BatchMixer::mix_resident(int frames, AudioFrame *bus_out, std::vector<Ref<SpatializerPlaybackData>> *finished) {
	GAS_FAIL_COND_V(!ctx, "no device context", false);
	std::vector<std::pair<Ref<AudioSpatializerInstance>, Ref<SpatializerPlaybackData>>> order;
	std::vector<gas_voice> voices;
	{
		std::lock_guard<std::mutex> lk(mu);
		if (refuse_custom_dsp()) {
			return false;
		}
		for (auto &ins : instances) {
			if (!ins) {
				continue;
			}
			for (auto &pb : ins->playbacks) {
				gas_voice v;
				v.voice = pb->voice_slot;
				v.instance = ins->slot;
				v.src_row = (int)voices.size(); // names the playback's row in the context's internal row buffer
				v.flags = 0u;
				voices.push_back(v);
				order.emplace_back(ins, pb);
			}
		}
		GAS_FAIL_COND_V(frames <= 0 || frames > cfg.max_frames || (frames & 1), "Condition \"p_frame_count != mix_buffer[ch].size()\" is true.", false);
		std::vector<int32_t> status(voices.size() ? voices.size() : 1);
		if (gas_mix_block_resident(ctx, (int)voices.size(), voices.data(), frames, reinterpret_cast<gas_frame *>(bus_out), status.data()) != GAS_OK) {
			set_last_error(gas_last_error(ctx));
			return false;
		}
		// _manage_playback_state (audio_spatializer.cpp:473-492): inactive playbacks leave the list
		for (size_t i = 0; i < order.size(); i++) {
			if (!(status[i] & GAS_VOICE_ACTIVE)) {
				if (finished) {
					finished->push_back(order[i].second);
				}
			} else {
				order[i].second.reset();
			}
		}
	}
	for (auto &e : order) {
		if (e.second) {
			e.first->stop_playback_stream(e.second);
		}
	}
	return true;
}

bool BatchMixer::set_bus_layout(const std::vector<gas_bus_desc> &layout) {
	GAS_FAIL_COND_V(!ctx, "no device context", false);
	if (gas_bus_layout_set(ctx, (int)layout.size(), layout.data()) != GAS_OK) {
		set_last_error(gas_last_error(ctx));
		return false;
	}
	return true;
}

bool BatchMixer::apply_bus_graph(int frames, AudioFrame *bus_inout) {
	GAS_FAIL_COND_V(!ctx, "no device context", false);
	if (gas_bus_graph(ctx, reinterpret_cast<gas_frame *>(bus_inout), frames) != GAS_OK) {
		set_last_error(gas_last_error(ctx));
		return false;
	}
	return true;
}

bool BatchMixer::voice_state(int voice_slot, gas_voice_state &out) {
	int32_t v = voice_slot;
	return ctx && gas_voice_state_export(ctx, 1, &v, &out) == GAS_OK;
}

} // namespace gas
