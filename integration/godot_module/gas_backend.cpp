// gas_backend.cpp — see gas_backend.h.  Host glue only: every gain and every sample is computed behind the C ABI.
#include "gas_backend.h"

#include "core/config/project_settings.h"

#include <string.h>

static GasBackend *singleton = nullptr;

GasBackend *GasBackend::get() {
	if (!singleton) {
		GasBackend *b = new GasBackend();
		if (!b->init()) {
			delete b;
			return nullptr;
		}
		singleton = b;
	}
	return singleton;
}

void GasBackend::shutdown() {
	if (singleton) {
		for (int i = 0; i < singleton->feeders.size(); i++) {
			AudioServer::get_singleton()->stop_playback_stream(singleton->feeders[i]);
		}
		if (singleton->ctx) {
			gas_destroy(singleton->ctx);
		}
		delete singleton;
		singleton = nullptr;
	}
}

bool GasBackend::init() {
	AudioServer *as = AudioServer::get_singleton();
	ERR_FAIL_NULL_V(as, false);
	gas_config c;
	gas_config_defaults(&c);
	max_slots = 65536;
	c.max_instances = c.max_voices = max_slots;
	c.max_frames = max_frames; // AudioServer's mix block
	num_buses = as->get_bus_count() < GAS_MAX_BUSES ? as->get_bus_count() : GAS_MAX_BUSES;
	c.num_buses = num_buses;
	c.speaker_mode = (int)as->get_speaker_mode();            // reference audio_spatializer_3d.cpp:59, :80, :113
	c.mix_rate = as->get_mix_rate();                         // :506, :571
	c.global_panning_strength = GLOBAL_GET_CACHED(float, "audio/general/3d_panning_strength"); // :633
	if (gas_create(&c, &ctx) != GAS_OK) {
		// no B200: the stock CPU classes of the reference module stay available, this class does not instantiate
		ERR_PRINT(gas_last_error(nullptr));
		ctx = nullptr;
		return false;
	}
	channels = gas_get_channel_count(ctx);
	pitch_of_instance.assign((size_t)max_slots, 1.0f);
	instance_started.assign((size_t)max_slots, 0);
	spat_owner.assign(16, nullptr);
	bus_out.assign((size_t)num_buses * channels * max_frames, gas_frame{ 0.f, 0.f });
	// one feeder per (bus, pair): the bus map sends pair `p` of the feeder's own bus at unit volume, nothing else
	for (int b = 0; b < num_buses; b++) {
		for (int p = 0; p < channels; p++) {
			Ref<GasFeederPlayback> f;
			f.instantiate();
			f->bus = b;
			f->pair = p;
			HashMap<StringName, Vector<AudioFrame>> map;
			Vector<AudioFrame> vol;
			vol.resize(AudioServer::MAX_CHANNELS_PER_BUS);
			for (int k = 0; k < AudioServer::MAX_CHANNELS_PER_BUS; k++) {
				vol.write[k] = k == p ? AudioFrame(1, 1) : AudioFrame(0, 0);
			}
			map[StringName(as->get_bus_name(b))] = vol;
			as->start_playback_stream(f, map);
			feeders.push_back(f);
		}
	}
	return true;
}

int GasBackend::alloc_instance() {
	mutex.lock();
	int s = -1;
	if (!free_instances.empty()) {
		s = free_instances.back();
		free_instances.pop_back();
	} else if (next_instance < max_slots) {
		s = next_instance++;
	}
	mutex.unlock();
	return s;
}
void GasBackend::free_instance(int p_slot) {
	mutex.lock();
	if (p_slot >= 0) {
		int32_t q = p_slot;
		gas_instance_stop(ctx, 1, &q);
		instance_started[(size_t)p_slot] = 0;
		free_instances.push_back(p_slot);
	}
	mutex.unlock();
}
int GasBackend::alloc_voice() {
	mutex.lock();
	int s = -1;
	if (!free_voices.empty()) {
		s = free_voices.back();
		free_voices.pop_back();
	} else if (next_voice < max_slots) {
		s = next_voice++;
	}
	if (s >= 0) {
		int32_t v = s;
		gas_voice_init(ctx, 1, &v); // instantiate_playback_data: zero state (reference audio_spatializer_3d.cpp:200-204)
	}
	mutex.unlock();
	return s;
}
void GasBackend::free_voice(int p_slot) {
	mutex.lock();
	if (p_slot >= 0) {
		free_voices.push_back(p_slot);
	}
	mutex.unlock();
}

int GasBackend::spatializer_slot(const void *p_resource, const gas_spatializer &p_pod) {
	mutex.lock();
	int slot = -1;
	for (size_t i = 0; i < spat_owner.size(); i++) {
		if (spat_owner[i] == p_resource) {
			slot = (int)i;
			break;
		}
	}
	for (size_t i = 0; slot < 0 && i < spat_owner.size(); i++) {
		if (!spat_owner[i]) {
			spat_owner[i] = p_resource;
			slot = (int)i;
		}
	}
	if (slot >= 0 && gas_spatializer_set(ctx, slot, &p_pod) != GAS_OK) { // the setters' validation ran on the resource already
		ERR_PRINT(gas_last_error(ctx));
	}
	mutex.unlock();
	return slot;
}

void GasBackend::queue_emitter(const gas_emitter &p_emitter, const gas_area *p_area) {
	mutex.lock();
	gas_emitter e = p_emitter;
	e.area = -1;
	if (p_area) {
		e.area = (int32_t)pending_areas.size();
		pending_areas.push_back(*p_area);
	}
	// the same instance may tick twice between two audio steps: the later tick wins
	bool replaced = false;
	for (size_t i = 0; i < pending_emitters.size(); i++) {
		if (pending_emitters[i].instance == e.instance) {
			pending_emitters[i] = e;
			replaced = true;
			break;
		}
	}
	if (!replaced) {
		pending_emitters.push_back(e);
	}
	mutex.unlock();
}

void GasBackend::set_listeners(const gas_listener *p_listeners, int p_count) {
	mutex.lock();
	listeners.assign(p_listeners, p_listeners + (p_count < GAS_MAX_LISTENERS ? p_count : GAS_MAX_LISTENERS));
	mutex.unlock();
}

float GasBackend::last_pitch_scale(int p_instance) const {
	return (p_instance >= 0 && p_instance < max_slots) ? pitch_of_instance[(size_t)p_instance] : 1.0f;
}

void GasBackend::capture(int p_voice, int p_instance, const AudioFrame *p_src, int p_frames, bool p_tail) {
	if (p_frames > max_frames || (captured_frames != 0 && captured_frames != p_frames)) {
		return; // reference audio_spatializer.cpp:521-522: the block size may not change under a running mixer
	}
	captured_frames = p_frames;
	gas_voice v;
	v.voice = p_voice;
	v.instance = p_instance;
	v.src_row = (int32_t)voices.size();
	v.flags = p_tail ? GAS_VOICE_WANT_PEAK : 0u;
	voices.push_back(v);
	const size_t at = staging.size();
	staging.resize(at + (size_t)p_frames);
	for (int i = 0; i < p_frames; i++) {
		staging[at + (size_t)i] = gas_frame{ p_src[i].left, p_src[i].right };
	}
}

void GasBackend::run_mix(int p_frames) {
	// 1. the gains queued by the physics thread since the last step, as one batch (hand-off under the mutex)
	mutex.lock();
	batch_emitters.swap(pending_emitters);
	batch_areas.swap(pending_areas);
	pending_emitters.clear();
	pending_areas.clear();
	std::vector<gas_listener> l = listeners;
	mutex.unlock();
	if (!batch_emitters.empty()) {
		batch_params.resize(batch_emitters.size());
		if (gas_gain_compute(ctx, (int32_t)batch_emitters.size(), batch_emitters.data(), (int32_t)l.size(), l.data(), (int32_t)batch_areas.size(),
					batch_areas.empty() ? nullptr : batch_areas.data(), batch_params.data()) != GAS_OK) {
			ERR_PRINT(gas_last_error(ctx));
		}
		std::vector<int32_t> to_start;
		for (size_t i = 0; i < batch_emitters.size(); i++) {
			const int q = batch_emitters[i].instance;
			pitch_of_instance[(size_t)q] = batch_params[i].pitch_scale;
			if (!instance_started[(size_t)q]) { // first parameters exist: register the proxies (reference audio_spatializer.cpp:75-95)
				instance_started[(size_t)q] = 1;
				to_start.push_back(q);
			}
		}
		if (!to_start.empty()) {
			gas_instance_start(ctx, (int32_t)to_start.size(), to_start.data());
		}
	}
	// 2. one mix block over every playback that delivered frames in this step
	const int frames = captured_frames > 0 ? captured_frames : p_frames;
	peaks.resize(voices.size() ? voices.size() : 1);
	if (gas_mix_block(ctx, (int32_t)voices.size(), voices.data(), staging.data(), (int32_t)voices.size(), frames, bus_out.data(), peaks.data()) != GAS_OK) {
		ERR_PRINT(gas_last_error(ctx));
		memset(bus_out.data(), 0, bus_out.size() * sizeof(gas_frame));
	}
	voices.clear();
	staging.clear();
	captured_frames = 0;
}

void GasBackend::serve(int p_bus, int p_pair, AudioFrame *p_out, int p_frames) {
	if (!mixed_this_step) {
		run_mix(p_frames);
		mixed_this_step = true;
	}
	const gas_frame *row = bus_out.data() + ((size_t)p_bus * channels + p_pair) * (size_t)p_frames;
	for (int i = 0; i < p_frames; i++) {
		p_out[i] = AudioFrame(row[i].l, row[i].r);
	}
	if (++served >= feeders.size()) { // every feeder has been asked: the step is over
		served = 0;
		mixed_this_step = false;
	}
}

int GasFeederPlayback::mix(AudioFrame *p_buffer, float p_rate_scale, int p_frames) {
	GasBackend *b = GasBackend::get();
	if (!b) {
		return 0;
	}
	b->serve(bus, pair, p_buffer, p_frames);
	return p_frames;
}
