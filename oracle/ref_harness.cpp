/*
 * ref_harness.cpp — drives the UNMODIFIED reference module (/root/reference/*.cpp, compiled against the
 * godot-lite stand-in headers in oracle/godot_lite/) through the same call sequence as the oracle and the
 * C ABI, so that tests can compare  oracle == reference code  on identical inputs.
 *
 * TEST INFRASTRUCTURE ONLY.  Built by `make -C oracle ref` into oracle/_ref/libgas_ref.so (git-ignored; only
 * where /root/reference exists).  Nothing under godot-audio-spatializer_b200/ may reference it.
 *
 * What runs here is the reference's own code: AudioSpatializer3D / AudioSpatializerInstance3D
 * (calculate_spatialization, calc_output_vol*, get_attenuation_db, calc_reverb_vol, process_frames,
 * mix_channel, SpeakerPlacementConfiguration), AudioSpatializerInstance (update_spatializer_parameters,
 * get_bus_map, start_playback_stream, _mix_from_playback_list incl. lookahead / end fade / deactivation,
 * get_mixed_frames), AudioStreamPlaybackSpatial::mix, SpatializerParameters, AudioSpatializer(Instance)Effect
 * (process_frames ping-pong chain, instantiate), AudioStreamPlayerSpatial (getters, get_bus fallback,
 * ENTER_TREE instantiation).  What does NOT come from the reference is upstream Godot (AudioServer mix step,
 * AudioFilterSW, Math, containers): restated in oracle/godot_lite/ from Godot 4.x as recalled.
 *
 * The harness plays the scene: it stores the query results the module asks the engine for (listener
 * transforms, the overlapping Area3D, closest points, velocities) and reads the module's outputs back.
 * This file is compiled with -fno-access-control so that it can read the module's private state (prev mix
 * volumes, filter processors, playback list nodes) without touching the reference sources.
 */
#include "../include/gas.h"

#include "audio_spatializer.h"
#include "audio_spatializer_3d.h"
#include "audio_spatializer_effect.h"
#include "audio_stream_player_spatial.h"
#include "register_types.h"
#include "scene/3d/audio_listener_3d.h"
#include "scene/3d/camera_3d.h"
#include "scene/3d/physics/area_3d.h"
#include "scene/3d/velocity_tracker_3d.h"
#include "scene/main/viewport.h"
#include "servers/audio/audio_server.h"
#ifdef GAS_REF_WITH_SHIM
#include "audio_spatializer_3d_gpu.h" /* integration/godot_module: the GPU-backed spatializer classes */
#include "gas_backend.h"
#endif

#include <algorithm>
#include <cmath>
#include <memory>
#include <vector>

namespace {

/* The AudioStreamPlayback of one voice.  Two feeding modes:
 *  - block mode (the C ABI's gas_mix_block contract: the caller supplies POST-lookahead frames): the harness
 *    writes frames [0,64) of the row into the list node's lookahead before the step and mix() delivers frames
 *    [64,F), so the module's playback_buffer[0..F) equals the row;
 *  - stream mode (gas_mix_block_stream: the caller supplies what AudioStreamPlayback::mix returned): mix()
 *    delivers `avail` new frames and the module's own lookahead / end-fade code does the rest. */
class HarnessPlayback : public AudioStreamPlayback {
	GDCLASS(HarnessPlayback, AudioStreamPlayback);

public:
	int voice = -1;
	const gas_frame *row = nullptr; /* this block's source row (NULL = nothing to deliver) */
	int avail = 0;                  /* frames mix() may deliver this block */
	int skip = 0;                   /* frames of the row consumed by the lookahead poke (block mode) */
	bool playing = false;
	float last_rate_scale = 0.0f;
	virtual void start(double p_from_pos = 0.0) override { playing = true; }
	virtual void stop() override { playing = false; }
	virtual bool is_playing() const override { return playing; }
	virtual int mix(AudioFrame *p_buffer, float p_rate_scale, int p_frames) override {
		last_rate_scale = p_rate_scale;
		int n = 0;
		if (row) {
			n = avail - skip;
			if (n > p_frames) {
				n = p_frames;
			}
			if (n < 0) {
				n = 0;
			}
			for (int i = 0; i < n; i++) {
				p_buffer[i] = AudioFrame(row[skip + i].l, row[skip + i].r);
			}
		}
		if (skip > 0) { /* block mode: the module must see a full block */
			for (int i = n; i < p_frames; i++) {
				p_buffer[i] = AudioFrame(0, 0);
			}
			return p_frames;
		}
		return n;
	}
};

struct RefInstance {
	AudioStreamPlayerSpatial *player = nullptr;
	Ref<AudioSpatializerInstance3D> companion; /* EFFECT kind: plays the script's _calculate_spatialization */
	int slot = -1;
	int kind = GAS_SPATIALIZER_3D;
	int effect_gain_binding = -1;
};

struct RefSpatializer {
	gas_spatializer pod;
	Ref<AudioSpatializer3D> res3d;       /* the 3D resource (for EFFECT: the formulas the script mirrors) */
	Ref<AudioSpatializerEffect> res_fx;  /* EFFECT kind */
	bool valid = false;
	bool gpu = false; // res3d is an AudioSpatializer3DGPU (integration/godot_module), see ref_use_gpu_shim
};

struct RefWorld {
	gas_config cfg;
	AudioServer server;
	Ref<World3D> world;
	std::vector<std::unique_ptr<Camera3D>> cams;
	std::vector<std::unique_ptr<Viewport>> vps;
	Viewport idle_vp; /* the players' viewport when there is no listener */
	Area3D area;
	std::vector<RefSpatializer> spat;
	std::vector<RefInstance> inst;
	std::vector<Ref<HarnessPlayback>> voice_pb;
	std::vector<int> voice_inst;
	std::vector<char> voice_fresh;
	std::vector<uint64_t> voice_seq; /* order of the ref_voice_init calls = start order */
	uint64_t next_seq = 1;
	/* per-call physics answers */
	const gas_area *cur_area = nullptr;
	int closest_calls = 0;
	bool stream_mode = false;
	bool use_gpu_shim = false;
};

String bus_name(int idx) {
	if (idx == 0) {
		return String("Master");
	}
	return String(std::string("Bus") + std::to_string(idx));
}

/* binds the singletons the module reaches through get_singleton() to this world for the duration of a call */
struct Scope {
	explicit Scope(RefWorld *w) {
		AudioServer::singleton_ptr() = &w->server;
		godot_lite::global_3d_panning_strength() = w->cfg.global_panning_strength;
		PhysicsDirectSpaceState3D &ss = PhysicsServer3D::get_singleton()->gl_space_state;
		ss.gl_intersect_point = [w](const PhysicsDirectSpaceState3D::PointParameters &, PhysicsDirectSpaceState3D::ShapeResult *r, int max) {
			if (!w->cur_area || max < 1) {
				return 0;
			}
			r[0].collider = &w->area;
			return 1;
		};
		ss.gl_closest_point = [w](RID, const Vector3 &) {
			const float *p = w->cur_area->closest_point[w->closest_calls < GAS_MAX_LISTENERS ? w->closest_calls : GAS_MAX_LISTENERS - 1];
			w->closest_calls++;
			return Vector3(p[0], p[1], p[2]);
		};
	}
};

void apply_3d_properties(AudioSpatializer3D *r, const gas_spatializer &s) {
	r->set_mix_channel_mode(s.mix_channel_mode != 0);
	r->set_attenuation_model((AudioSpatializer3D::AttenuationModel)s.attenuation_model);
	r->set_unit_size(s.unit_size);
	r->set_max_distance(s.max_distance);
	r->set_panning_strength(s.panning_strength);
	r->set_area_mask(s.area_mask);
	r->set_emission_angle_enabled(s.emission_angle_enabled != 0);
	r->set_emission_angle(s.emission_angle);
	r->set_emission_angle_filter_attenuation_db(s.emission_angle_filter_attenuation_db);
	r->set_attenuation_filter_cutoff_hz(s.attenuation_filter_cutoff_hz);
	r->set_attenuation_filter_db(s.attenuation_filter_db);
	r->set_doppler_tracking((AudioSpatializer3D::DopplerTracking)s.doppler_tracking);
	r->set_doppler_speed_of_sound(s.doppler_speed_of_sound);
}

void apply_effect(AudioEffectFilter *f, const gas_effect &e) {
	f->mode = (AudioFilterSW::Mode)e.mode;
	f->set_cutoff(e.cutoff_hz);
	f->set_resonance(e.resonance);
	f->set_gain(e.gain);
	int st = e.stages < 1 ? 1 : (e.stages > 4 ? 4 : e.stages);
	f->set_db((AudioEffectFilter::FilterDB)(st - 1));
}

AudioSpatializerInstance *instance_of(RefInstance &q) {
	return q.player ? q.player->spatializer.ptr() : nullptr;
}

void destroy_instance(RefInstance &q) {
	if (q.player) {
		q.companion.unref();
		q.player->spatializer.unref(); /* while the player it unhooks from is still whole */
		delete q.player;
		q.player = nullptr;
	}
}

void fill_params(RefWorld *w, const Ref<SpatializerParameters> &p, gas_params *out) {
	memset(out, 0, sizeof(*out));
	out->pitch_scale = 1.0f;
	out->attenuation_filter_cutoff_hz = 5000.0f;
	if (p.is_null()) {
		return;
	}
	Vector<Vector2> mv = p->get_mix_volumes();
	for (int c = 0; c < 4 && c < mv.size(); c++) {
		out->mix_volumes[c][0] = mv[c].x;
		out->mix_volumes[c][1] = mv[c].y;
	}
	out->pitch_scale = p->get_pitch_scale();
	out->update_parameters = p->should_update_parameters() ? 1 : 0;
	if (SpatializerParameters3D *p3 = Object::cast_to<SpatializerParameters3D>(*p)) {
		out->linear_attenuation = p3->get_linear_attenuation();
		out->attenuation_filter_cutoff_hz = p3->get_attenuation_filter_cutoff_hz();
	}
	Dictionary bv = p->get_bus_volumes();
	int n = 0;
	for (StringName key : bv.get_key_list()) {
		if (n >= GAS_MAX_BUSES_PER_PLAYBACK) {
			break;
		}
		out->bus[n] = w->server.gl_bus_index(key);
		Vector<Vector2> v = bv.get_valid(key);
		for (int c = 0; c < 4 && c < v.size(); c++) {
			out->bus_volumes[n][c][0] = v[c].x;
			out->bus_volumes[n][c][1] = v[c].y;
		}
		n++;
	}
	out->n_bus = n;
}

Ref<SpatializerParameters3D> params_from_pod(const gas_params *p) {
	Ref<SpatializerParameters3D> r;
	r.instantiate();
	Vector<Vector2> mv;
	mv.resize(4);
	for (int c = 0; c < 4; c++) {
		mv.write[c] = Vector2(p->mix_volumes[c][0], p->mix_volumes[c][1]);
	}
	r->set_mix_volumes(mv);
	r->set_pitch_scale(p->pitch_scale);
	r->set_update_parameters(p->update_parameters != 0);
	r->set_linear_attenuation(p->linear_attenuation);
	r->set_attenuation_filter_cutoff_hz(p->attenuation_filter_cutoff_hz);
	for (int k = 0; k < p->n_bus && k < GAS_MAX_BUSES_PER_PLAYBACK; k++) {
		Vector<Vector2> v;
		v.resize(4);
		for (int c = 0; c < 4; c++) {
			v.write[c] = Vector2(p->bus_volumes[k][c][0], p->bus_volumes[k][c][1]);
		}
		r->add_bus_volume(StringName(bus_name(p->bus[k])), v);
	}
	return r;
}

void set_listeners(RefWorld *w, int n, const gas_listener *ls) {
	while ((int)w->cams.size() < n) {
		w->cams.emplace_back(new Camera3D());
		w->vps.emplace_back(new Viewport());
	}
	w->world->gl_cameras.clear();
	for (int i = 0; i < n; i++) {
		Camera3D *c = w->cams[(size_t)i].get();
		Viewport *v = w->vps[(size_t)i].get();
		Transform3D t;
		for (int r = 0; r < 3; r++) {
			t.basis.rows[r] = Vector3(ls[i].basis[r * 3 + 0], ls[i].basis[r * 3 + 1], ls[i].basis[r * 3 + 2]);
		}
		t.origin = Vector3(ls[i].origin[0], ls[i].origin[1], ls[i].origin[2]);
		c->gl_global_transform = t;
		c->gl_velocity = Vector3(ls[i].velocity[0], ls[i].velocity[1], ls[i].velocity[2]);
		c->gl_viewport = v;
		c->gl_world_3d = w->world;
		v->gl_camera_3d = c;
		v->gl_is_audio_listener_3d = true;
		w->world->gl_cameras.insert(c);
	}
}

void pose_player(RefWorld *w, RefInstance &q, const gas_emitter *e, int n_listeners) {
	AudioStreamPlayerSpatial *pl = q.player;
	Transform3D t;
	/* only column 2 of the basis is read (emission direction is -Z, audio_spatializer_3d.cpp:380) */
	t.basis.set_column(2, Vector3(e->basis_z[0], e->basis_z[1], e->basis_z[2]));
	t.origin = Vector3(e->origin[0], e->origin[1], e->origin[2]);
	pl->gl_global_transform = t;
	pl->gl_world_3d = w->world;
	pl->gl_viewport = n_listeners > 0 ? w->vps[0].get() : &w->idle_vp;
	pl->set_volume_db(e->volume_db);
	pl->set_max_db(e->max_db);
	pl->set_pitch_scale(e->pitch_scale);
	/* an out-of-range bus index becomes a name AudioServer does not know: get_bus() falls back to Master
	 * (audio_stream_player_spatial.cpp:405-413) */
	pl->set_bus(StringName((e->bus >= 0 && e->bus < w->cfg.num_buses) ? bus_name(e->bus) : String("NoSuchBus")));
	Vector3 vel(e->velocity[0], e->velocity[1], e->velocity[2]);
	if (AudioSpatializerInstance3D *i3 = Object::cast_to<AudioSpatializerInstance3D>(instance_of(q))) {
		i3->velocity_tracker->gl_velocity = vel;
	}
#ifdef GAS_REF_WITH_SHIM
	if (AudioSpatializerInstance3DGPU *ig = Object::cast_to<AudioSpatializerInstance3DGPU>(instance_of(q))) {
		ig->velocity_tracker->gl_velocity = vel;
	}
#endif
	if (q.companion.is_valid()) {
		q.companion->velocity_tracker->gl_velocity = vel;
	}
}

void set_area(RefWorld *w, const gas_area *a) {
	w->cur_area = a;
	w->closest_calls = 0;
	if (a) {
		w->area.gl_audio_bus_override = a->override_bus != 0;
		w->area.gl_audio_bus_name = StringName(bus_name(a->bus));
		w->area.gl_use_reverb_bus = a->use_reverb != 0;
		w->area.gl_reverb_bus_name = StringName(bus_name(a->reverb_bus));
		w->area.gl_reverb_amount = a->reverb_amount;
		w->area.gl_reverb_uniformity = a->reverb_uniformity;
	}
}

} // namespace

extern "C" {

typedef struct RefWorld ref_world;

GAS_API ref_world *ref_create(const gas_config *cfg) {
	if (!cfg || cfg->max_instances <= 0 || cfg->max_voices <= 0 || cfg->max_spatializers <= 0 || cfg->num_buses < 1 ||
			cfg->num_buses > GAS_MAX_BUSES || cfg->speaker_mode < 0 || cfg->speaker_mode > 3) {
		return nullptr;
	}
	static bool registered = false;
	if (!registered) { /* the module's own registration entry point (register_types.cpp:40-60) */
		initialize_audio_spatializer_module(MODULE_INITIALIZATION_LEVEL_SCENE);
		registered = true;
	}
	RefWorld *w = new RefWorld();
	w->cfg = *cfg;
	w->server.gl_speaker_mode = (AudioServer::SpeakerMode)cfg->speaker_mode;
	w->server.gl_mix_rate = cfg->mix_rate;
	w->server.gl_bus_names.clear();
	for (int b = 0; b < cfg->num_buses; b++) {
		w->server.gl_bus_names.push_back(bus_name(b));
	}
	w->world.instantiate();
	w->spat.resize((size_t)cfg->max_spatializers);
	w->inst.resize((size_t)cfg->max_instances);
	w->voice_pb.resize((size_t)cfg->max_voices);
	w->voice_inst.assign((size_t)cfg->max_voices, -1);
	w->voice_fresh.assign((size_t)cfg->max_voices, 0);
	w->voice_seq.assign((size_t)cfg->max_voices, 0);
	/* playbacks are mixed in ascending (instance, channel) order: the order only fixes the float summation
	 * order across instances, which upstream leaves to the start order */
	w->server.gl_order_key = [w](AudioStreamPlayback *p) -> int64_t {
		AudioStreamPlaybackSpatial *sp = Object::cast_to<AudioStreamPlaybackSpatial>(p);
		if (!sp) {
			return INT64_MAX; // not a spatializer proxy (the GPU shim's feeders): after every proxy, as upstream's newest-first list has them
		}
		for (size_t i = 0; i < w->inst.size(); i++) {
			if (w->inst[i].player && w->inst[i].player->spatializer.ptr() == sp->spatializer) {
				return (int64_t)i * 8 + sp->channel;
			}
		}
		return 0;
	};
	return w;
}

GAS_API void ref_destroy(ref_world *w) {
	if (!w) {
		return;
	}
	Scope sc(w);
	for (RefInstance &q : w->inst) {
		destroy_instance(q);
	}
#ifdef GAS_REF_WITH_SHIM
	if (w->use_gpu_shim) {
		GasBackend::shutdown();
	}
#endif
	w->server.gl_playbacks.clear();
	delete w;
}

GAS_API int ref_error_count(void) { return godot_lite::error_log().count; }
GAS_API const char *ref_last_error(void) { return godot_lite::error_log().last.c_str(); }
GAS_API int ref_registered_classes(void) { return (int)godot_lite::registered_classes().size(); }

GAS_API int ref_set_speaker_mode(ref_world *w, int mode) {
	if (mode < 0 || mode > 3) {
		return GAS_ERR_INVALID;
	}
	w->cfg.speaker_mode = mode;
	w->server.gl_speaker_mode = (AudioServer::SpeakerMode)mode;
	return GAS_OK;
}
GAS_API int ref_set_mix_rate(ref_world *w, float hz) {
	w->cfg.mix_rate = hz;
	w->server.gl_mix_rate = hz;
	return GAS_OK;
}
GAS_API int ref_set_global_panning_strength(ref_world *w, float s) {
	w->cfg.global_panning_strength = s; /* read at instance construction (audio_spatializer_3d.cpp:633) */
	return GAS_OK;
}
/* upstream AudioServer mixes each playback through its own 64-frame lookahead; off by default because the
 * batched path is defined at the bus-accumulate input. */
/* 1 = AudioSpatializer3D resources are created as AudioSpatializer3DGPU (integration/godot_module): the same scene, the same
 * reference plumbing around it, the arithmetic on the device.  Only in the library built with the shim (libgas_ref_gpu.so);
 * one world at a time (the shim keeps one device context per process, like one AudioServer per engine). */
GAS_API int ref_use_gpu_shim(ref_world *w, int on) {
#ifdef GAS_REF_WITH_SHIM
	w->use_gpu_shim = on != 0;
	return GAS_OK;
#else
	(void)w;
	return on ? GAS_ERR_STATE : GAS_OK;
#endif
}
GAS_API int ref_has_gpu_shim(void) {
#ifdef GAS_REF_WITH_SHIM
	return 1;
#else
	return 0;
#endif
}

GAS_API int ref_set_server_lookahead(ref_world *w, int on) {
	w->server.gl_playback_lookahead = on != 0;
	return GAS_OK;
}

/* Property setters run the reference's validation (audio_spatializer_3d.cpp:670-760): a rejected value
 * leaves the old one in place and logs an error; the return value says whether any setter complained. */
GAS_API int ref_spatializer_set(ref_world *w, int slot, const gas_spatializer *s) {
	if (slot < 0 || slot >= w->cfg.max_spatializers || !s) {
		return GAS_ERR_INVALID;
	}
	Scope sc(w);
	RefSpatializer &r = w->spat[(size_t)slot];
	int errors0 = godot_lite::error_log().count;
	if (!r.valid || r.pod.kind != s->kind || r.gpu != w->use_gpu_shim) {
#ifdef GAS_REF_WITH_SHIM
		if (w->use_gpu_shim && s->kind == GAS_SPATIALIZER_3D) {
			Ref<AudioSpatializer3DGPU> g;
			g.instantiate();
			r.res3d = g;
		} else
#endif
		{
			r.res3d.instantiate();
		}
		r.gpu = w->use_gpu_shim;
		r.res_fx.unref();
		if (s->kind == GAS_SPATIALIZER_EFFECT) {
			r.res_fx.instantiate();
		}
		r.valid = true;
	}
	apply_3d_properties(r.res3d.ptr(), *s);
	if (s->kind == GAS_SPATIALIZER_EFFECT) {
		TypedArray<AudioEffect> fx;
		for (int k = 0; k < s->chain.n_effects && k < GAS_MAX_EFFECTS; k++) {
			Ref<AudioEffectFilter> f;
			f.instantiate();
			apply_effect(f.ptr(), s->chain.effects[k]);
			fx.push_back(f);
		}
		r.res_fx->set_audio_effects(fx);
	}
	r.pod = *s;
	return godot_lite::error_log().count == errors0 ? GAS_OK : GAS_ERR_INVALID;
}

GAS_API int ref_instance_init(ref_world *w, int n, const int32_t *instances, const int32_t *spatializers) {
	Scope sc(w);
	for (int i = 0; i < n; i++) {
		if (instances[i] < 0 || instances[i] >= w->cfg.max_instances || spatializers[i] < 0 || spatializers[i] >= w->cfg.max_spatializers ||
				!w->spat[(size_t)spatializers[i]].valid) {
			return GAS_ERR_INVALID;
		}
	}
	for (int i = 0; i < n; i++) {
		RefInstance &q = w->inst[(size_t)instances[i]];
		RefSpatializer &r = w->spat[(size_t)spatializers[i]];
		if (q.player) { /* drop the old proxies with the old instance */
			AudioSpatializerInstance *old = instance_of(q);
			if (old) {
				for (int c = 0; c < old->spatial_playbacks.size(); c++) {
					Ref<AudioStreamPlayback> pb = old->spatial_playbacks[c];
					w->server.gl_playbacks.remove_if([&pb](const std::unique_ptr<AudioServer::PlaybackNode> &nd) { return nd->stream_playback == pb; });
				}
			}
			destroy_instance(q);
		}
		q.slot = spatializers[i];
		q.kind = r.pod.kind;
		q.effect_gain_binding = r.pod.effect_gain_binding;
		q.player = new AudioStreamPlayerSpatial();
		q.player->gl_world_3d = w->world;
		q.player->gl_viewport = &w->idle_vp;
		q.player->gl_inside_tree = true;
		q.player->set_max_polyphony(1 << 20);
		if (r.pod.kind == GAS_SPATIALIZER_EFFECT) {
			q.player->set_spatializer(r.res_fx);
		} else {
			q.player->set_spatializer(r.res3d);
		}
		/* AudioStreamPlayerSpatial::_notification(ENTER_TREE) instantiates the spatializer and binds the player
		 * (audio_stream_player_spatial.cpp:48-56) */
		q.player->notification(Node::NOTIFICATION_ENTER_TREE);
		AudioSpatializerInstance *ins = instance_of(q);
		if (!ins) {
			return GAS_ERR_STATE;
		}
		if (r.pod.kind == GAS_SPATIALIZER_EFFECT) {
			/* AudioSpatializerInstanceEffect leaves _calculate_spatialization and _process_effects to a script
			 * (gd_spatializer_instance.gd:86-127).  The script of the example mirrors the C++ formulas, so the
			 * hook delegates to an AudioSpatializerInstance3D bound to the same player, and binds one effect's
			 * gain to the computed high-shelf gain like the example's _process_effects does. */
			q.companion = r.res3d->instantiate();
			q.companion->set_audio_player(q.player);
			AudioSpatializerInstance3D *comp = q.companion.ptr();
			ins->_gdvirtual__calculate_spatialization_hook = [comp](Ref<SpatializerParameters> &r_ret) {
				r_ret = comp->calculate_spatialization();
				return true;
			};
			AudioSpatializerInstanceEffect *fxi = Object::cast_to<AudioSpatializerInstanceEffect>(ins);
			int binding = r.pod.effect_gain_binding;
			fxi->_gdvirtual__process_effects_hook = [fxi, binding](Ref<SpatializerParameters> p, Ref<SpatializerPlaybackData>) {
				SpatializerParameters3D *p3 = Object::cast_to<SpatializerParameters3D>(*p);
				if (p3 && binding >= 0 && binding < (int)fxi->audio_effects.size()) {
					Ref<AudioEffectFilter> f = fxi->audio_effects[binding];
					if (f.is_valid()) {
						f->set_gain(p3->get_linear_attenuation());
					}
				}
				return true;
			};
		}
		/* block mode leaves voice lifetimes to the caller, like the C ABI's gas_mix_block: a NaN threshold never
		 * compares <= (audio_spatializer.cpp:465).  Stream mode restores the module's default (-80 dB). */
		ins->set_playback_disable_threshold_db(NAN);
	}
	return GAS_OK;
}

/* The reference registers the proxies when the first voice of an inactive instance starts
 * (audio_spatializer.cpp:75-95) and stops them when its last voice is gone (:484-491): both happen inside
 * ref_mix_block*.  These two exist so that the oracle's call sequence runs unchanged. */
GAS_API int ref_instance_start(ref_world *, int, const int32_t *) { return GAS_OK; }
GAS_API int ref_instance_stop(ref_world *, int, const int32_t *) { return GAS_OK; }

GAS_API int ref_voice_init(ref_world *w, int n, const int32_t *voices) {
	Scope sc(w);
	for (int i = 0; i < n; i++) {
		if (voices[i] < 0 || voices[i] >= w->cfg.max_voices) {
			return GAS_ERR_INVALID;
		}
	}
	for (int i = 0; i < n; i++) {
		int v = voices[i];
		if (w->voice_pb[(size_t)v].is_valid() && w->voice_inst[(size_t)v] >= 0) {
			AudioSpatializerInstance *ins = instance_of(w->inst[(size_t)w->voice_inst[(size_t)v]]);
			if (ins) {
				ins->stop_playback_stream(w->voice_pb[(size_t)v]);
			}
		}
		w->voice_pb[(size_t)v].unref();
		w->voice_inst[(size_t)v] = -1;
		w->voice_fresh[(size_t)v] = 1;
		w->voice_seq[(size_t)v] = w->next_seq++;
	}
	return GAS_OK;
}

GAS_API int ref_gain_compute(ref_world *w, int n, const gas_emitter *emitters, int n_listeners, const gas_listener *listeners,
		int n_areas, const gas_area *areas, gas_params *out_params) {
	if (n < 0 || n_listeners < 0 || n_listeners > GAS_MAX_LISTENERS) {
		return GAS_ERR_INVALID;
	}
	Scope sc(w);
	for (int i = 0; i < n; i++) {
		const gas_emitter *e = &emitters[i];
		if (e->instance < 0 || e->instance >= w->cfg.max_instances || e->area >= n_areas || !w->inst[(size_t)e->instance].player ||
				w->inst[(size_t)e->instance].slot != e->spatializer) {
			return GAS_ERR_INVALID;
		}
	}
	set_listeners(w, n_listeners, listeners);
	for (int i = 0; i < n; i++) {
		const gas_emitter *e = &emitters[i];
		RefInstance &q = w->inst[(size_t)e->instance];
		pose_player(w, q, e, n_listeners);
		set_area(w, e->area >= 0 ? &areas[e->area] : nullptr);
		/* what the physics tick does (audio_stream_player_spatial.cpp:74-79) */
		instance_of(q)->update_spatializer_parameters();
		if (out_params) {
			fill_params(w, instance_of(q)->get_spatializer_parameters(), &out_params[i]);
		}
	}
	w->cur_area = nullptr;
	return GAS_OK;
}

/* Parameters computed elsewhere: set_spatializer_parameters + the bus-map push of
 * update_spatializer_parameters (audio_spatializer.cpp:263-271), with get_bus_map being the module's. */
GAS_API int ref_params_set(ref_world *w, int n, const int32_t *instances, const gas_params *params) {
	Scope sc(w);
	for (int i = 0; i < n; i++) {
		if (instances[i] < 0 || instances[i] >= w->cfg.max_instances || !w->inst[(size_t)instances[i]].player) {
			return GAS_ERR_INVALID;
		}
	}
	for (int i = 0; i < n; i++) {
		AudioSpatializerInstance *ins = instance_of(w->inst[(size_t)instances[i]]);
		Ref<SpatializerParameters> p = params_from_pod(&params[i]);
		ins->set_spatializer_parameters(p);
		if (p->should_update_parameters()) {
			for (int c = 0; c < ins->spatial_playbacks.size(); c++) {
				Ref<AudioStreamPlayback> playback = ins->spatial_playbacks[c];
				AudioServer::get_singleton()->set_playback_bus_volumes_linear(playback, ins->get_bus_map(p, c));
			}
		}
	}
	return GAS_OK;
}

GAS_API int ref_params_get(ref_world *w, int n, const int32_t *instances, gas_params *out) {
	Scope sc(w);
	for (int i = 0; i < n; i++) {
		if (instances[i] < 0 || instances[i] >= w->cfg.max_instances || !w->inst[(size_t)instances[i]].player) {
			return GAS_ERR_INVALID;
		}
		fill_params(w, instance_of(w->inst[(size_t)instances[i]])->get_spatializer_parameters(), &out[i]);
	}
	return GAS_OK;
}

GAS_API int ref_effect_params_set(ref_world *w, int n, const int32_t *instances, const gas_effect_chain *chains) {
	Scope sc(w);
	for (int i = 0; i < n; i++) {
		if (instances[i] < 0 || instances[i] >= w->cfg.max_instances) {
			return GAS_ERR_INVALID;
		}
		AudioSpatializerInstanceEffect *fxi = Object::cast_to<AudioSpatializerInstanceEffect>(instance_of(w->inst[(size_t)instances[i]]));
		if (!fxi) {
			continue;
		}
		for (int k = 0; k < chains[i].n_effects && k < (int)fxi->audio_effects.size(); k++) {
			Ref<AudioEffectFilter> f = fxi->audio_effects[k];
			if (f.is_valid()) {
				apply_effect(f.ptr(), chains[i].effects[k]);
			}
		}
	}
	return GAS_OK;
}

static int mix_common(ref_world *w, int n_voices, const gas_voice *voices, const gas_frame *src, int src_rows, int frames,
		const int32_t *mixed_frames, gas_frame *bus_out, int32_t *active_out) {
	if (n_voices < 0 || frames <= 0) {
		return GAS_ERR_INVALID;
	}
	const bool stream = mixed_frames != nullptr;
	for (int i = 0; i < n_voices; i++) {
		const gas_voice &v = voices[i];
		if (v.voice < 0 || v.voice >= w->cfg.max_voices || v.instance < 0 || v.instance >= w->cfg.max_instances || v.src_row >= src_rows ||
				!w->inst[(size_t)v.instance].player) {
			return GAS_ERR_INVALID;
		}
	}
	Scope sc(w);
	/* start fresh voices in the order they were initialised (the module's SafeList then iterates newest first) */
	std::vector<int> fresh;
	for (int i = 0; i < n_voices; i++) {
		if (w->voice_fresh[(size_t)voices[i].voice]) {
			fresh.push_back(i);
		}
	}
	std::stable_sort(fresh.begin(), fresh.end(), [&](int a, int b) { return w->voice_seq[(size_t)voices[a].voice] < w->voice_seq[(size_t)voices[b].voice]; });
	for (int i : fresh) {
		const gas_voice &v = voices[i];
		AudioSpatializerInstance *ins = instance_of(w->inst[(size_t)v.instance]);
		ins->set_playback_disable_threshold_db(stream ? -80.0f : NAN);
		Ref<HarnessPlayback> pb;
		pb.instantiate();
		pb->voice = v.voice;
		w->voice_pb[(size_t)v.voice] = pb;
		w->voice_inst[(size_t)v.voice] = v.instance;
		w->voice_fresh[(size_t)v.voice] = 0;
		ins->start_playback_stream(pb, 0.0f); /* audio_spatializer.cpp:44-96 */
	}
	for (int i = 0; i < n_voices; i++) {
		const gas_voice &v = voices[i];
		Ref<HarnessPlayback> pb = w->voice_pb[(size_t)v.voice];
		if (pb.is_null()) {
			return GAS_ERR_STATE;
		}
		AudioSpatializerInstance *ins = instance_of(w->inst[(size_t)v.instance]);
		AudioSpatializerInstance::SpatialPlaybackListNode *node = ins->_find_playback_list_node(pb);
		pb->row = v.src_row >= 0 ? src + (size_t)v.src_row * (size_t)frames : nullptr;
		if (stream) {
			pb->skip = 0;
			pb->avail = v.src_row >= 0 ? mixed_frames[i] : 0;
		} else {
			pb->avail = frames;
			pb->skip = AudioSpatializerInstance::LOOKAHEAD_BUFFER_SIZE < frames ? AudioSpatializerInstance::LOOKAHEAD_BUFFER_SIZE : frames;
			if (node) {
				if (v.src_row < 0) {
					node->has_frames.clear(); /* the zero-filled playback buffer of audio_spatializer.cpp:405-408 */
				} else {
					node->has_frames.set();
					for (int k = 0; k < pb->skip; k++) {
						node->lookahead[k] = AudioFrame(pb->row[k].l, pb->row[k].r);
					}
				}
			}
		}
	}
	w->server.gl_mix_step(frames); /* AudioServer -> AudioStreamPlaybackSpatial::mix -> get_mixed_frames -> _mix_from_playback_list */
	const int channels = w->cfg.speaker_mode + 1;
	if (bus_out) {
		for (int b = 0; b < w->cfg.num_buses; b++) {
			for (int c = 0; c < channels; c++) {
				const std::vector<AudioFrame> &buf = w->server.gl_bus_buffers[(size_t)b][(size_t)c];
				gas_frame *o = bus_out + ((size_t)b * channels + c) * frames;
				for (int i = 0; i < frames; i++) {
					o[i].l = buf[(size_t)i].left;
					o[i].r = buf[(size_t)i].right;
				}
			}
		}
	}
	if (active_out) {
		for (int i = 0; i < n_voices; i++) {
			const gas_voice &v = voices[i];
			AudioSpatializerInstance *ins = instance_of(w->inst[(size_t)v.instance]);
			AudioSpatializerInstance::SpatialPlaybackListNode *node = ins->_find_playback_list_node(w->voice_pb[(size_t)v.voice]);
			active_out[i] = node ? ((node->active.is_set() ? 1 : 0) | (node->has_frames.is_set() ? 2 : 0)) : 0;
		}
	}
	return GAS_OK;
}

/* block mode: rows are post-lookahead playback buffers (the gas_mix_block contract) */
GAS_API int ref_mix_block(ref_world *w, int n_voices, const gas_voice *voices, const gas_frame *src, int src_rows, int frames,
		gas_frame *bus_out) {
	return mix_common(w, n_voices, voices, src, src_rows, frames, nullptr, bus_out, nullptr);
}

/* stream mode: row r holds the mixed_frames[i] frames AudioStreamPlayback::mix returned for voice i this block;
 * the module's own lookahead splice, end-of-stream fade and tail deactivation run (audio_spatializer.cpp:369-408,
 * :464-469).  active_out[i]: bit 0 = node still active after the block, bit 1 = has_frames; 0 = node deleted. */
GAS_API int ref_mix_block_stream(ref_world *w, int n_voices, const gas_voice *voices, const gas_frame *src, int src_rows, int frames,
		const int32_t *mixed_frames, gas_frame *bus_out, int32_t *active_out) {
	if (!mixed_frames) {
		return GAS_ERR_INVALID;
	}
	return mix_common(w, n_voices, voices, src, src_rows, frames, mixed_frames, bus_out, active_out);
}

GAS_API int ref_voice_state_export(ref_world *w, int n, const int32_t *voices, gas_voice_state *out) {
	Scope sc(w);
	for (int i = 0; i < n; i++) {
		memset(&out[i], 0, sizeof(gas_voice_state));
		int v = voices[i];
		if (v < 0 || v >= w->cfg.max_voices) {
			return GAS_ERR_INVALID;
		}
		if (w->voice_pb[(size_t)v].is_null() || w->voice_inst[(size_t)v] < 0) {
			continue;
		}
		AudioSpatializerInstance *ins = instance_of(w->inst[(size_t)w->voice_inst[(size_t)v]]);
		AudioSpatializerInstance::SpatialPlaybackListNode *node = ins ? ins->_find_playback_list_node(w->voice_pb[(size_t)v]) : nullptr;
		if (!node) {
			continue;
		}
		if (SpatializerPlaybackData3D *d = Object::cast_to<SpatializerPlaybackData3D>(*node->playback_data)) {
			for (int c = 0; c < 4; c++) {
				Vector2 pv = d->get_prev_mix_volume(c);
				out[i].prev_mix_volumes[c][0] = pv.x;
				out[i].prev_mix_volumes[c][1] = pv.y;
			}
			for (int k = 0; k < 8; k++) {
				const AudioFilterSW::Processor &p = d->filter_processors[k];
				gas_processor_state &o = out[i].filter_processors[k];
				o.b0 = p.gl_coeffs().b0;
				o.b1 = p.gl_coeffs().b1;
				o.b2 = p.gl_coeffs().b2;
				o.a1 = p.gl_coeffs().a1;
				o.a2 = p.gl_coeffs().a2;
				float h[4];
				p.gl_history(h);
				o.ha1 = h[0];
				o.ha2 = h[1];
				o.hb1 = h[2];
				o.hb2 = h[3];
			}
		} else if (SpatializerPlaybackDataEffect *fx = Object::cast_to<SpatializerPlaybackDataEffect>(*node->playback_data)) {
			Vector<Ref<AudioEffectInstance>> eis = fx->get_effect_instances();
			for (int j = 0; j < eis.size() && j < GAS_MAX_EFFECTS; j++) {
				AudioEffectFilterInstance *fi = Object::cast_to<AudioEffectFilterInstance>(*eis[j]);
				if (!fi) {
					continue;
				}
				for (int side = 0; side < 2; side++) {
					for (int s = 0; s < GAS_MAX_FILTER_STAGES; s++) {
						fi->gl_processor(side, s).gl_history(out[i].effect_history[j][side][s]);
					}
				}
			}
		}
	}
	return GAS_OK;
}

/* ---- scalar pieces, for known-answer comparisons with the oracle's orc_* twins ---------------------- */
struct ScalarRig {
	RefWorld *w;
	int32_t zero = 0;
	explicit ScalarRig(int speaker_mode, float global_panning, float mix_rate, const gas_spatializer *s) {
		gas_config cfg;
		memset(&cfg, 0, sizeof(cfg));
		cfg.max_instances = 1;
		cfg.max_voices = 1;
		cfg.max_frames = 4096;
		cfg.max_spatializers = 1;
		cfg.num_buses = GAS_MAX_BUSES;
		cfg.speaker_mode = speaker_mode;
		cfg.mix_rate = mix_rate;
		cfg.global_panning_strength = global_panning;
		w = ref_create(&cfg);
		ref_spatializer_set(w, 0, s);
		ref_instance_init(w, 1, &zero, &zero);
	}
	~ScalarRig() { ref_destroy(w); }
	AudioSpatializerInstance3D *i3() { return Object::cast_to<AudioSpatializerInstance3D>(instance_of(w->inst[0])); }
};

/* audio_spatializer_3d.cpp:123-151 */
GAS_API float ref_get_attenuation_db(const gas_spatializer *s, float volume_db, float max_db, float distance) {
	ScalarRig rig(0, 0.5f, 48000.0f, s);
	Scope sc(rig.w);
	rig.w->inst[0].player->set_volume_db(volume_db);
	rig.w->inst[0].player->set_max_db(max_db);
	return rig.i3()->get_attenuation_db(distance);
}

/* audio_spatializer_3d.cpp:112-121 (-> :103-110 stereo, :57-98 + :903-938 surround) */
GAS_API void ref_calc_output_vol(int speaker_mode, float global_panning, float panning_strength, const float dir[3], float out[4][2]) {
	gas_spatializer s;
	memset(&s, 0, sizeof(s));
	s.unit_size = 10.0f;
	s.panning_strength = panning_strength;
	s.emission_angle = 45.0f;
	s.doppler_speed_of_sound = 343.0f;
	ScalarRig rig(speaker_mode, global_panning, 48000.0f, &s);
	Scope sc(rig.w);
	Vector<Vector2> o;
	o.resize(4);
	for (int c = 0; c < 4; c++) {
		o.write[c] = Vector2(out[c][0], out[c][1]);
	}
	rig.i3()->calc_output_vol(Vector3(dir[0], dir[1], dir[2]), o);
	for (int c = 0; c < 4; c++) {
		out[c][0] = o[c].x;
		out[c][1] = o[c].y;
	}
}

/* SpeakerPlacementConfiguration, audio_spatializer_3d.cpp:903-938 with the module's default directions (:47-55,
 * reached through calc_output_vol_surround's update_speaker_configuration call) */
GAS_API void ref_spcap_calculate(int speaker_count, const float dir[3], float tightness, float volumes[7], float eff[7]) {
	int mode = speaker_count == 3 ? 1 : (speaker_count == 5 ? 2 : (speaker_count == 7 ? 3 : 0));
	gas_spatializer s;
	memset(&s, 0, sizeof(s));
	s.unit_size = 10.0f;
	s.panning_strength = 1.0f;
	s.emission_angle = 45.0f;
	s.doppler_speed_of_sound = 343.0f;
	ScalarRig rig(mode, 0.5f, 48000.0f, &s);
	Scope sc(rig.w);
	Vector<Vector2> o;
	o.resize(4);
	rig.i3()->calc_output_vol_surround(Vector3(dir[0], dir[1], dir[2]), tightness, o); /* configures spcap for speaker_count */
	SpeakerPlacementConfiguration *sp = rig.i3()->base->spcap;
	for (int i = 0; i < 7; i++) {
		volumes[i] = 0.0f;
		eff[i] = 0.0f;
	}
	sp->calculate(Vector3(dir[0], dir[1], dir[2]), tightness, (unsigned int)speaker_count, volumes);
	for (unsigned int i = 0; i < sp->get_speaker_count(); i++) {
		eff[i] = sp->speakers[i].effective_number_of_speakers;
	}
}

/* audio_spatializer.cpp:274-324 */
GAS_API int ref_get_bus_map(const gas_params *p, int mix_channels, int channel, int out_bus[6], float out_vol[6][4][2]) {
	gas_spatializer s;
	memset(&s, 0, sizeof(s));
	s.unit_size = 10.0f;
	s.panning_strength = 1.0f;
	s.emission_angle = 45.0f;
	s.doppler_speed_of_sound = 343.0f;
	s.mix_channel_mode = mix_channels;
	ScalarRig rig(3, 0.5f, 48000.0f, &s);
	Scope sc(rig.w);
	HashMap<StringName, Vector<AudioFrame>> m = rig.i3()->get_bus_map(params_from_pod(p), channel);
	int n = 0;
	for (const KeyValue<StringName, Vector<AudioFrame>> &kv : m) {
		out_bus[n] = rig.w->server.gl_bus_index(kv.key);
		for (int c = 0; c < 4; c++) {
			out_vol[n][c][0] = kv.value[c].left;
			out_vol[n][c][1] = kv.value[c].right;
		}
		n++;
	}
	return n;
}

static Ref<SpatializerPlaybackData3D> playback_data_from(const gas_voice_state *st) {
	Ref<SpatializerPlaybackData3D> d;
	d.instantiate();
	for (int c = 0; c < 4; c++) {
		d->set_prev_mix_volume(c, Vector2(st->prev_mix_volumes[c][0], st->prev_mix_volumes[c][1]));
	}
	for (int k = 0; k < 8; k++) {
		AudioFilterSW::Processor &p = d->filter_processors[k];
		const gas_processor_state &s = st->filter_processors[k];
		p.coeffs.b0 = s.b0;
		p.coeffs.b1 = s.b1;
		p.coeffs.b2 = s.b2;
		p.coeffs.a1 = s.a1;
		p.coeffs.a2 = s.a2;
		p.ha1 = s.ha1;
		p.ha2 = s.ha2;
		p.hb1 = s.hb1;
		p.hb2 = s.hb2;
	}
	return d;
}
static void playback_data_to(const Ref<SpatializerPlaybackData3D> &d, gas_voice_state *st) {
	for (int c = 0; c < 4; c++) {
		Vector2 pv = d->get_prev_mix_volume(c);
		st->prev_mix_volumes[c][0] = pv.x;
		st->prev_mix_volumes[c][1] = pv.y;
	}
	for (int k = 0; k < 8; k++) {
		const AudioFilterSW::Processor &p = d->filter_processors[k];
		gas_processor_state &s = st->filter_processors[k];
		s.b0 = p.coeffs.b0;
		s.b1 = p.coeffs.b1;
		s.b2 = p.coeffs.b2;
		s.a1 = p.coeffs.a1;
		s.a2 = p.coeffs.a2;
		s.ha1 = p.ha1;
		s.ha2 = p.ha2;
		s.hb1 = p.hb1;
		s.hb2 = p.hb2;
	}
}

/* audio_spatializer_3d.cpp:491-552 / :554-609 on one voice state */
GAS_API void ref_process_frames_3d(const gas_params *p, gas_voice_state *st, float mix_rate, gas_frame *out, const gas_frame *src, int frames) {
	gas_spatializer s;
	memset(&s, 0, sizeof(s));
	s.unit_size = 10.0f;
	s.panning_strength = 1.0f;
	s.emission_angle = 45.0f;
	s.doppler_speed_of_sound = 343.0f;
	ScalarRig rig(3, 0.5f, mix_rate, &s);
	Scope sc(rig.w);
	Ref<SpatializerPlaybackData3D> d = playback_data_from(st);
	std::vector<AudioFrame> in((size_t)frames), o((size_t)frames);
	for (int i = 0; i < frames; i++) {
		in[(size_t)i] = AudioFrame(src[i].l, src[i].r);
	}
	rig.i3()->process_frames(params_from_pod(p), d, o.data(), in.data(), frames);
	for (int i = 0; i < frames; i++) {
		out[i].l = o[(size_t)i].left;
		out[i].r = o[(size_t)i].right;
	}
	playback_data_to(d, st);
}
GAS_API void ref_mix_channel_3d(const gas_params *p, gas_voice_state *st, float mix_rate, int channel, gas_frame *out, const gas_frame *src, int frames) {
	gas_spatializer s;
	memset(&s, 0, sizeof(s));
	s.unit_size = 10.0f;
	s.panning_strength = 1.0f;
	s.emission_angle = 45.0f;
	s.doppler_speed_of_sound = 343.0f;
	s.mix_channel_mode = 1;
	ScalarRig rig(3, 0.5f, mix_rate, &s);
	Scope sc(rig.w);
	Ref<SpatializerPlaybackData3D> d = playback_data_from(st);
	std::vector<AudioFrame> in((size_t)frames), o((size_t)frames);
	for (int i = 0; i < frames; i++) {
		in[(size_t)i] = AudioFrame(src[i].l, src[i].r);
	}
	rig.i3()->mix_channel(params_from_pod(p), d, channel, o.data(), in.data(), frames);
	for (int i = 0; i < frames; i++) {
		out[i].l = o[(size_t)i].left;
		out[i].r = o[(size_t)i].right;
	}
	playback_data_to(d, st);
}

/* upstream AudioFilterSW::prepare_coefficients as restated in godot_lite (NOT reference code): lets a test
 * check that the two independent restatements (oracle, godot-lite) agree. */
GAS_API void ref_filter_prepare_coefficients(int mode, float cutoff, float resonance, float gain, int stages, float sampling_rate, float out[5]) {
	AudioFilterSW f;
	f.set_mode((AudioFilterSW::Mode)mode);
	f.set_cutoff(cutoff);
	f.set_resonance(resonance);
	f.set_gain(gain);
	f.set_stages(stages);
	f.set_sampling_rate(sampling_rate);
	AudioFilterSW::Coeffs c;
	f.prepare_coefficients(&c);
	out[0] = c.b0;
	out[1] = c.b1;
	out[2] = c.b2;
	out[3] = c.a1;
	out[4] = c.a2;
}

} // extern "C"
