// audio_spatializer_3d_gpu.cpp — see audio_spatializer_3d_gpu.h.
#include "audio_spatializer_3d_gpu.h"

#include "gas_backend.h"

#include "audio_stream_player_spatial.h"
#include "scene/3d/audio_listener_3d.h"
#include "scene/3d/camera_3d.h"
#include "scene/3d/velocity_tracker_3d.h"
#include "scene/main/viewport.h"
#ifndef PHYSICS_3D_DISABLED
#include "scene/3d/physics/area_3d.h"
#endif

#include <string.h>

SpatializerPlaybackDataGPU::~SpatializerPlaybackDataGPU() {
	if (GasBackend *b = GasBackend::get()) {
		b->free_voice(voice_slot);
	}
}

AudioSpatializerInstance3DGPU::AudioSpatializerInstance3DGPU() {
	velocity_tracker.instantiate();
}

AudioSpatializerInstance3DGPU::~AudioSpatializerInstance3DGPU() {
	if (get_audio_player() != nullptr) {
		get_audio_player()->remove_transform_changed_callback(_transform_changed_cb, this);
	}
	if (GasBackend *b = GasBackend::get()) {
		b->free_instance(instance_slot);
	}
}

// Doppler bookkeeping of the player node, like the stock instance does (reference audio_spatializer_3d.cpp:611-629)
void AudioSpatializerInstance3DGPU::initialize_audio_player() {
	if (get_audio_player() == nullptr || gpu_base.is_null()) {
		return;
	}
	if ((int)gpu_base->get_doppler_tracking() != 0 /* DOPPLER_TRACKING_DISABLED: the enum is private to the resource */) {
		get_audio_player()->add_transform_changed_callback(_transform_changed_cb, this);
		velocity_tracker->set_track_physics_step((int)gpu_base->get_doppler_tracking() == 2 /* DOPPLER_TRACKING_PHYSICS_STEP */);
		if (get_audio_player()->is_inside_tree()) {
			velocity_tracker->reset(get_audio_player()->get_global_transform().origin);
		}
	}
}

void AudioSpatializerInstance3DGPU::update_doppler_tracked_velocity() {
	if (gpu_base.is_valid() && (int)gpu_base->get_doppler_tracking() != 0 /* DOPPLER_TRACKING_DISABLED: the enum is private to the resource */) {
		velocity_tracker->update_position(get_audio_player()->get_global_transform().origin);
	}
}

#ifndef PHYSICS_3D_DISABLED
// which Area3D diverts this player's sound (same query as reference audio_spatializer_3d.cpp:206-245)
Area3D *AudioSpatializerInstance3DGPU::_get_overriding_area() {
	Ref<World3D> world_3d = get_audio_player()->get_world_3d();
	ERR_FAIL_COND_V(world_3d.is_null(), nullptr);
	PhysicsDirectSpaceState3D *space_state = PhysicsServer3D::get_singleton()->space_get_direct_state(world_3d->get_space());
	PhysicsDirectSpaceState3D::ShapeResult sr[MAX_INTERSECT_AREAS];
	PhysicsDirectSpaceState3D::PointParameters point_params;
	point_params.position = get_audio_player()->get_global_transform().origin;
	point_params.collision_mask = gpu_base->get_area_mask();
	point_params.collide_with_bodies = false;
	point_params.collide_with_areas = true;
	const int areas = space_state->intersect_point(point_params, sr, MAX_INTERSECT_AREAS);
	for (int i = 0; i < areas; i++) {
		Area3D *tarea = sr[i].collider ? Object::cast_to<Area3D>(sr[i].collider) : nullptr;
		if (tarea && (tarea->is_overriding_audio_bus() || tarea->is_using_reverb_bus())) {
			return tarea;
		}
	}
	return nullptr;
}
#endif

void AudioSpatializer3DGPU::to_pod(gas_spatializer &r_pod) const {
	gas_spatializer_defaults(&r_pod);
	r_pod.kind = GAS_SPATIALIZER_3D;
	r_pod.attenuation_model = (int32_t)get_attenuation_model();
	r_pod.unit_size = get_unit_size();
	r_pod.max_distance = get_max_distance();
	r_pod.panning_strength = get_panning_strength();
	r_pod.area_mask = get_area_mask();
	r_pod.emission_angle_enabled = is_emission_angle_enabled() ? 1 : 0;
	r_pod.emission_angle = get_emission_angle();
	r_pod.emission_angle_filter_attenuation_db = get_emission_angle_filter_attenuation_db();
	r_pod.attenuation_filter_cutoff_hz = get_attenuation_filter_cutoff_hz();
	r_pod.attenuation_filter_db = get_attenuation_filter_db();
	r_pod.doppler_tracking = (int32_t)get_doppler_tracking();
	r_pod.doppler_speed_of_sound = get_doppler_speed_of_sound();
	r_pod.mix_channel_mode = get_mix_channel_mode() ? 1 : 0;
}

Ref<AudioSpatializerInstance> AudioSpatializer3DGPU::instantiate() {
	GasBackend *b = GasBackend::get();
	if (!b) { // no usable device: behave like the stock resource (CPU path of the reference module)
		return AudioSpatializer3D::instantiate();
	}
	Ref<AudioSpatializerInstance3DGPU> ins;
	ins.instantiate();
	ins->gpu_base = Ref<AudioSpatializer3DGPU>(this);
	gas_spatializer pod;
	to_pod(pod);
	const int spat = b->spatializer_slot(this, pod);
	ins->instance_slot = b->alloc_instance();
	ERR_FAIL_COND_V_MSG(spat < 0 || ins->instance_slot < 0, Ref<AudioSpatializerInstance>(), "out of spatializer / instance slots");
	int32_t q = ins->instance_slot, s = spat;
	if (gas_instance_init(b->context(), 1, &q, &s) != GAS_OK) {
		ERR_PRINT(gas_last_error(b->context()));
	}
	return ins;
}

Ref<SpatializerPlaybackData> AudioSpatializerInstance3DGPU::instantiate_playback_data() {
	Ref<SpatializerPlaybackDataGPU> d;
	d.instantiate();
	if (GasBackend *b = GasBackend::get()) {
		d->voice_slot = b->alloc_voice();
	}
	return d;
}

// The scene side of reference audio_spatializer_3d.cpp:277-354 (which cameras listen, where they are, which Area3D the
// player is in, the area's closest point per listener) stays here; everything from :356 on is the gain kernel's.
Ref<SpatializerParameters> AudioSpatializerInstance3DGPU::calculate_spatialization() {
	Ref<SpatializerParameters3D> parameters;
	GasBackend *b = GasBackend::get();
	AudioStreamPlayerSpatial *player = get_audio_player();
	ERR_FAIL_NULL_V(player, parameters);
	ERR_FAIL_NULL_V(b, parameters);
	Ref<World3D> world_3d = player->get_world_3d();
	ERR_FAIL_COND_V(world_3d.is_null(), parameters);
	parameters.instantiate();

	gas_spatializer pod; // property edits since the last tick reach the device with this tick
	gpu_base->to_pod(pod);
	const int spat = b->spatializer_slot(gpu_base.ptr(), pod);

	gas_listener listeners[GAS_MAX_LISTENERS];
	Node3D *listener_nodes[GAS_MAX_LISTENERS];
	int n_listeners = 0;
	HashSet<Camera3D *> cameras = world_3d->get_cameras();
	cameras.insert(player->get_viewport()->get_camera_3d());
	for (Camera3D *camera : cameras) {
		if (!camera || n_listeners >= GAS_MAX_LISTENERS) {
			continue;
		}
		Viewport *vp = camera->get_viewport();
		if (!vp || !vp->is_audio_listener_3d()) {
			continue;
		}
		Node3D *node = camera;
		Vector3 velocity = camera->get_doppler_tracked_velocity();
		if (AudioListener3D *listener = vp->get_audio_listener_3d()) {
			node = listener;
			velocity = listener->get_doppler_tracked_velocity();
		}
		const Transform3D t = node->get_global_transform();
		gas_listener &l = listeners[n_listeners];
		for (int r = 0; r < 3; r++) {
			for (int c = 0; c < 3; c++) {
				l.basis[r * 3 + c] = t.basis[r][c];
			}
		}
		l.origin[0] = t.origin.x, l.origin[1] = t.origin.y, l.origin[2] = t.origin.z;
		l.velocity[0] = velocity.x, l.velocity[1] = velocity.y, l.velocity[2] = velocity.z;
		listener_nodes[n_listeners++] = node;
	}
	b->set_listeners(listeners, n_listeners);

	gas_emitter e;
	memset(&e, 0, sizeof(e));
	e.instance = instance_slot;
	e.spatializer = spat;
	const Transform3D gt = player->get_global_transform();
	const Vector3 bz = gt.basis.get_column(2);
	e.origin[0] = gt.origin.x, e.origin[1] = gt.origin.y, e.origin[2] = gt.origin.z;
	e.basis_z[0] = bz.x, e.basis_z[1] = bz.y, e.basis_z[2] = bz.z;
	e.volume_db = player->get_volume_db();
	e.max_db = player->get_max_db();
	e.pitch_scale = player->get_pitch_scale();
	if ((int)gpu_base->get_doppler_tracking() != 0 /* DOPPLER_TRACKING_DISABLED: the enum is private to the resource */) { // reference :296-299
		const Vector3 v = velocity_tracker->get_tracked_linear_velocity();
		e.velocity[0] = v.x, e.velocity[1] = v.y, e.velocity[2] = v.z;
	}
	e.bus = AudioServer::get_singleton()->get_bus_index(player->get_bus()); // unknown => Master inside the library
	gas_area area;
	const gas_area *area_ptr = nullptr;
#ifndef PHYSICS_3D_DISABLED
	if (Area3D *a = _get_overriding_area()) { // reference :206-245
		memset(&area, 0, sizeof(area));
		area.override_bus = a->is_overriding_audio_bus() ? 1 : 0;
		area.bus = AudioServer::get_singleton()->get_bus_index(a->get_audio_bus_name());
		area.use_reverb = a->is_using_reverb_bus() ? 1 : 0;
		area.reverb_bus = AudioServer::get_singleton()->get_bus_index(a->get_reverb_bus_name());
		area.reverb_amount = a->get_reverb_amount();
		area.reverb_uniformity = a->get_reverb_uniformity();
		if (area.use_reverb && area.reverb_uniformity > 0) { // reference :350-353
			PhysicsDirectSpaceState3D *space_state = PhysicsServer3D::get_singleton()->space_get_direct_state(world_3d->get_space());
			for (int i = 0; i < n_listeners; i++) {
				const Vector3 p = space_state->get_closest_point_to_object_volume(a->get_rid(), listener_nodes[i]->get_global_transform().origin);
				area.closest_point[i][0] = p.x, area.closest_point[i][1] = p.y, area.closest_point[i][2] = p.z;
			}
		}
		area_ptr = &area;
	}
#endif
	b->queue_emitter(e, area_ptr);

	// What the reference's own plumbing still reads from the parameters: the pitch the playbacks are resampled with
	// (audio_spatializer.cpp:375) — the Doppler pitch the gain kernel computed on the previous tick — and a bus map for the
	// instance's proxy playback: none, so AudioServer mixes nothing from it.
	Vector<Vector2> silent;
	silent.resize(AudioServer::MAX_CHANNELS_PER_BUS);
	silent.fill(Vector2(0, 0));
	parameters->set_mix_volumes(silent);
	parameters->set_pitch_scale(b->last_pitch_scale(instance_slot));
	parameters->set_update_parameters(true);
	return parameters;
}

void AudioSpatializerInstance3DGPU::process_frames(Ref<SpatializerParameters> p_parameters, Ref<SpatializerPlaybackData> p_playback_data,
		AudioFrame *p_output_buf, const AudioFrame *p_source_buf, int p_frame_count) {
	ERR_FAIL_COND_MSG(!Object::cast_to<SpatializerPlaybackDataGPU>(*p_playback_data), "Unexpected SpatializerPlaybackData type; expected SpatializerPlaybackDataGPU");
	SpatializerPlaybackDataGPU *data = Object::cast_to<SpatializerPlaybackDataGPU>(*p_playback_data);
	GasBackend *b = GasBackend::get();
	if (b && data->voice_slot >= 0) {
		b->capture(data->voice_slot, instance_slot, p_source_buf, p_frame_count, /* tail */ false);
	}
	// The output only feeds the reference's peak detection (audio_spatializer.cpp:449-461): handing the source back keeps
	// "still audible" meaning what it means there for the unfiltered signal.
	for (int i = 0; i < p_frame_count; i++) {
		p_output_buf[i] = p_source_buf[i];
	}
}
