// audio_spatializer_3d_gpu.h — AudioSpatializer3D with its arithmetic on a B200.
//
// AudioSpatializer3DGPU is a drop-in for AudioSpatializer3D in a scene: same properties (it IS an AudioSpatializer3D,
// reference audio_spatializer_3d.h:153-241), assigned to AudioStreamPlayerSpatial::spatializer the same way.  Its
// instance keeps the reference's scene queries (cameras and listeners, the overriding Area3D, closest points) and the
// reference's playback plumbing (AudioSpatializerInstance::_mix_from_playback_list pulls the streams, splices the
// lookahead, fades stream ends) and replaces the two hot loops through the plugin virtuals only:
//   calculate_spatialization()   queues a gas_emitter for the batched gain kernel instead of evaluating
//                                reference audio_spatializer_3d.cpp:342-463 per object
//   process_frames()             hands the playback's block to the batched mixer instead of filtering it here; the
//                                instance's own proxy playback is given an empty bus map, so AudioServer mixes nothing
//                                from it — the batched result reaches the buses through GasBackend's feeder playbacks
// Nothing in the reference module is modified.
#pragma once

#include "audio_spatializer_3d.h"

#include "gas.h"

class AudioSpatializer3DGPU;

class SpatializerPlaybackDataGPU : public SpatializerPlaybackData {
	GDCLASS(SpatializerPlaybackDataGPU, SpatializerPlaybackData);

public:
	int voice_slot = -1;
	~SpatializerPlaybackDataGPU();
};

// Derives from the plugin base, not from AudioSpatializerInstance3D: that class keeps its resource in a private member
// only AudioSpatializer3D::instantiate() can set (reference audio_spatializer_3d.h:106-108), so its scene helpers are
// re-done here (same queries, reference audio_spatializer_3d.cpp:206-245, :611-641).
class AudioSpatializerInstance3DGPU : public AudioSpatializerInstance {
	GDCLASS(AudioSpatializerInstance3DGPU, AudioSpatializerInstance);
	friend class AudioSpatializer3DGPU;

	Ref<AudioSpatializer3DGPU> gpu_base;
	Ref<VelocityTracker3D> velocity_tracker;
	int instance_slot = -1;

	static void _transform_changed_cb(void *self) { reinterpret_cast<AudioSpatializerInstance3DGPU *>(self)->update_doppler_tracked_velocity(); }
#ifndef PHYSICS_3D_DISABLED
	Area3D *_get_overriding_area();
#endif

public:
	AudioSpatializerInstance3DGPU();
	~AudioSpatializerInstance3DGPU();
	virtual void initialize_audio_player() override;
	void update_doppler_tracked_velocity();

	virtual Ref<SpatializerParameters> calculate_spatialization() override;
	virtual void process_frames(Ref<SpatializerParameters> p_parameters, Ref<SpatializerPlaybackData> p_playback_data, AudioFrame *p_output_buf,
			const AudioFrame *p_source_buf, int p_frame_count) override;
	virtual Ref<SpatializerPlaybackData> instantiate_playback_data() override;
	// towards the reference's mix driver the instance always looks like Mode A (one proxy, one process_frames call per
	// playback and block); mix_channel_mode is honoured on the device
	virtual bool should_process_frames() const override { return true; }
	virtual bool should_mix_channels() const override { return false; }
};

class AudioSpatializer3DGPU : public AudioSpatializer3D {
	GDCLASS(AudioSpatializer3DGPU, AudioSpatializer3D);

protected:
	static void _bind_methods() {}

public:
	virtual Ref<AudioSpatializerInstance> instantiate() override;
	void to_pod(gas_spatializer &r_pod) const;
};
