// gas_backend.h — the one device context behind every AudioSpatializerInstance3DGPU, and the two places where the
// per-object calls of the plugin API become one batched call each:
//   physics thread   calculate_spatialization() of every instance appends a gas_emitter (+ gas_area) to the pending
//                    batch; the batch is computed by ONE gas_gain_compute at the top of the next audio mix step
//   audio thread     process_frames() of every playback hands its post-lookahead block to capture(); the first feeder
//                    playback AudioServer asks in a mix step runs ONE gas_mix_block over everything captured and every
//                    feeder then serves its (bus, channel pair) row of the result
// Feeders are ordinary AudioStreamPlaybacks registered with AudioServer on one bus each at full volume; they are
// registered before any spatializer proxy, and AudioServer's playback list is newest-first, so within a mix step they
// are asked after every proxy has delivered (= captured) its frames.
#pragma once

#include "gas.h"

#include "servers/audio/audio_server.h"
#include "servers/audio/audio_stream.h"

#include <vector>

class GasBackend;

class GasFeederPlayback : public AudioStreamPlayback {
	GDCLASS(GasFeederPlayback, AudioStreamPlayback);
	friend class GasBackend;
	int bus = 0;
	int pair = 0;

public:
	virtual void start(double p_from_pos = 0.0) override {}
	virtual void stop() override {}
	virtual bool is_playing() const override { return true; }
	virtual int get_loop_count() const override { return 0; }
	virtual double get_playback_position() const override { return 0; }
	virtual void tag_used_streams() override {}
	virtual int mix(AudioFrame *p_buffer, float p_rate_scale, int p_frames) override;
};

class GasBackend {
public:
	static GasBackend *get();  // created on first use (AudioServer must exist); nullptr without a usable device
	static void shutdown();

	gas_ctx *context() { return ctx; }
	int alloc_instance();
	void free_instance(int p_slot);
	int alloc_voice();
	void free_voice(int p_slot);
	int spatializer_slot(const void *p_resource, const gas_spatializer &p_pod); // registers / refreshes a resource

	// physics thread
	void queue_emitter(const gas_emitter &p_emitter, const gas_area *p_area);
	void set_listeners(const gas_listener *p_listeners, int p_count);
	float last_pitch_scale(int p_instance) const; // Doppler pitch of the previous tick (what the playback is resampled with)

	// audio thread
	void capture(int p_voice, int p_instance, const AudioFrame *p_src, int p_frames, bool p_tail);
	void serve(int p_bus, int p_pair, AudioFrame *p_out, int p_frames);

private:
	GasBackend() {}
	bool init();
	void run_mix(int p_frames);

	gas_ctx *ctx = nullptr;
	Mutex mutex; // the parameter hand-off of the reference (audio_spatializer.cpp:558-574), for the whole batch
	int channels = 1, num_buses = 1, max_frames = 512;
	std::vector<int> free_instances, free_voices;
	int next_instance = 0, next_voice = 0, max_slots = 0;
	std::vector<const void *> spat_owner;
	// pending gains (physics thread) / their double buffer (audio thread)
	std::vector<gas_emitter> pending_emitters, batch_emitters;
	std::vector<gas_area> pending_areas, batch_areas;
	std::vector<gas_listener> listeners;
	std::vector<gas_params> batch_params;
	std::vector<float> pitch_of_instance;
	std::vector<char> instance_started;
	// captured voices of the current mix step
	std::vector<gas_voice> voices;
	std::vector<gas_frame> staging, bus_out, peaks;
	int captured_frames = 0;
	// feeders
	Vector<Ref<GasFeederPlayback>> feeders;
	int served = 0;
	bool mixed_this_step = false;
};
