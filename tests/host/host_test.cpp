// host_test — exercises the C++ host mirror (godot-audio-spatializer_b200/host) the way a Godot-side caller would.
//   host_test validate                 : setter / parameter validation (reference ERR_FAIL_* behaviour); no GPU needed
//   host_test scene <in.bin> <out.bin> : plays a scene read from in.bin through BatchMixer on cuda:0, writes the bus buffers
// in.bin : int32 {V, F, blocks, speaker_mode, num_buses, mix_channel_mode, custom_last, has_area}, gas_area,
//          then per block: gas_emitter[V], float[V][F][2]
// out.bin: per block float[num_buses][channels][F][2]
#include "../../godot-audio-spatializer_b200/host/audio_spatializer_gpu.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

using namespace gas;

#define CHECK(c)                                                  \
	do {                                                          \
		if (!(c)) {                                               \
			fprintf(stderr, "CHECK failed: %s (line %d)\n", #c, __LINE__); \
			return 1;                                             \
		}                                                         \
	} while (0)

// a custom spatializer in the style of the reference's plugin API: its instance supplies its own parameters
class FixedInstance : public AudioSpatializerInstance3D {
public:
	Ref<SpatializerParameters> calculate_spatialization() override {
		auto p = std::make_shared<SpatializerParameters3D>();
		std::vector<Vector2> v(4, Vector2{ 0.25f, 0.5f });
		p->set_mix_volumes(v);
		p->add_bus_volume(0, v);
		p->set_pitch_scale(1.0f);
		p->set_linear_attenuation(0.0f);
		p->set_update_parameters(true);
		return p;
	}
};
class FixedSpatializer : public AudioSpatializer3D {
public:
	Ref<AudioSpatializerInstance> instantiate() override {
		auto i = std::make_shared<FixedInstance>();
		i->base = shared_from_this();
		i->mix_channel_mode = get_mix_channel_mode();
		return i;
	}
};

static int validate() {
	auto s = std::make_shared<AudioSpatializer3D>();
	// defaults, audio_spatializer_3d.h:171-188
	CHECK(s->get_attenuation_model() == AudioSpatializer3D::ATTENUATION_INVERSE_DISTANCE);
	CHECK(s->get_unit_size() == 10.0f && s->get_max_distance() == 0.0f && s->get_panning_strength() == 1.0f);
	CHECK(s->get_area_mask() == 1u && !s->is_emission_angle_enabled() && s->get_emission_angle() == 45.0f);
	CHECK(s->get_emission_angle_filter_attenuation_db() == -12.0f && s->get_attenuation_filter_cutoff_hz() == 5000.0f);
	CHECK(s->get_attenuation_filter_db() == -24.0f && s->get_doppler_speed_of_sound() == 343.0f && !s->get_mix_channel_mode());
	// validation: invalid values are refused and leave the property untouched (audio_spatializer_3d.cpp:671,696,729,738,759)
	CHECK(!s->set_max_distance(-1.0f) && s->get_max_distance() == 0.0f);
	CHECK(s->set_max_distance(50.0f) && s->get_max_distance() == 50.0f);
	CHECK(!s->set_emission_angle(91.0f) && !s->set_emission_angle(-0.5f) && s->get_emission_angle() == 45.0f);
	CHECK(s->set_emission_angle(90.0f));
	CHECK(!s->set_attenuation_model(4) && !s->set_attenuation_model(-1) && s->set_attenuation_model(2));
	CHECK(!s->set_panning_strength(-0.1f) && s->get_panning_strength() == 1.0f && s->set_panning_strength(0.0f));
	CHECK(!s->set_doppler_speed_of_sound(0.0f) && s->set_doppler_speed_of_sound(1.0f));
	// SpatializerParameters: volumes must have exactly 4 entries (spatializer_parameters.cpp:36,45)
	SpatializerParameters3D p;
	CHECK(!p.set_mix_volumes(std::vector<Vector2>(3)) && p.get_mix_volumes().size() == 4);
	CHECK(!p.add_bus_volume(1, std::vector<Vector2>(5)) && p.get_bus_volumes().empty());
	CHECK(p.add_bus_volume(3, std::vector<Vector2>(4, Vector2{ 1.f, 2.f })) && p.add_bus_volume(1, std::vector<Vector2>(4)));
	CHECK(p.add_bus_volume(3, std::vector<Vector2>(4, Vector2{ 5.f, 6.f }))); // same key: overwritten in place
	CHECK(p.get_bus_volumes().size() == 2 && p.get_bus_volumes()[0].first == 3 && p.get_bus_volumes()[0].second[0].x == 5.f);
	gas_params pod;
	p.set_linear_attenuation(0.5f);
	p.to_pod(pod);
	CHECK(pod.n_bus == 2 && pod.bus[0] == 3 && pod.bus[1] == 1 && pod.bus_volumes[0][2][1] == 6.f && pod.linear_attenuation == 0.5f);
	SpatializerParameters3D q;
	q.from_pod(pod);
	CHECK(q.get_bus_volumes().size() == 2 && q.get_linear_attenuation() == 0.5f);
	// instance: NaN volume rejected (audio_stream_player_spatial.cpp:193)
	AudioSpatializerInstance3D inst;
	CHECK(!inst.set_volume_db(nanf("")) && inst.set_volume_db(-6.0f));
	CHECK(inst.should_process_frames() && !inst.should_mix_channels()); // resource default mix_channel_mode = false wins (Q19)
	auto e = std::make_shared<AudioSpatializerEffect>();
	gas_effect fx{ GAS_FILTER_HIGHSHELF, 4000.f, 1.f, 0.3f, 1 };
	for (int i = 0; i < GAS_MAX_EFFECTS; i++) {
		CHECK(e->add_effect(fx));
	}
	CHECK(!e->add_effect(fx) && e->get_effect_count() == GAS_MAX_EFFECTS && !e->set_effect(7, fx));
	printf("validate ok\n");
	return 0;
}

static int scene(const char *in_path, const char *out_path) {
	FILE *fi = fopen(in_path, "rb");
	CHECK(fi);
	int32_t h[8];
	CHECK(fread(h, sizeof(int32_t), 8, fi) == 8);
	const int V = h[0], F = h[1], blocks = h[2], speaker_mode = h[3], num_buses = h[4], mode_b = h[5], custom_last = h[6], has_area = h[7];
	gas_area area;
	CHECK(fread(&area, sizeof(area), 1, fi) == 1);
	BatchMixerConfig cfg;
	cfg.max_instances = V;
	cfg.max_voices = V;
	cfg.max_frames = F;
	cfg.num_buses = num_buses;
	cfg.speaker_mode = speaker_mode;
	cfg.mix_rate = 48000.f;
	BatchMixer mixer(cfg);
	if (!mixer.ok()) {
		fprintf(stderr, "no device: %s\n", last_error().c_str());
		return 2;
	}
	auto spat = std::make_shared<AudioSpatializer3D>();
	spat->set_mix_channel_mode(mode_b != 0);
	auto fixed = std::make_shared<FixedSpatializer>();
	fixed->set_mix_channel_mode(mode_b != 0);
	std::vector<Ref<AudioSpatializerInstance>> inst(V);
	for (int i = 0; i < V; i++) {
		inst[i] = mixer.instantiate((custom_last && i == V - 1) ? std::static_pointer_cast<AudioSpatializer>(fixed) : std::static_pointer_cast<AudioSpatializer>(spat));
		CHECK(inst[i]);
	}
	gas_listener l;
	memset(&l, 0, sizeof(l));
	l.basis[0] = l.basis[4] = l.basis[8] = 1.f;
	mixer.set_listeners({ l });
	const int C = speaker_mode + 1;
	std::vector<gas_emitter> em(V);
	std::vector<AudioFrame> src((size_t)V * F), bus((size_t)num_buses * C * F);
	FILE *fo = fopen(out_path, "wb");
	CHECK(fo);
	for (int b = 0; b < blocks; b++) {
		CHECK(fread(em.data(), sizeof(gas_emitter), V, fi) == (size_t)V);
		CHECK(fread(src.data(), sizeof(AudioFrame), (size_t)V * F, fi) == (size_t)V * F);
		for (int i = 0; i < V; i++) {
			Transform3D t;
			t.origin = Vector3{ em[i].origin[0], em[i].origin[1], em[i].origin[2] };
			t.basis.rows[0][2] = em[i].basis_z[0], t.basis.rows[1][2] = em[i].basis_z[1], t.basis.rows[2][2] = em[i].basis_z[2];
			inst[i]->set_global_transform(t);
			inst[i]->set_volume_db(em[i].volume_db);
			inst[i]->set_max_db(em[i].max_db);
			inst[i]->set_pitch_scale(em[i].pitch_scale);
			inst[i]->set_bus(em[i].bus);
			inst[i]->set_area((has_area && em[i].area >= 0) ? &area : nullptr);
			if (b == 0) {
				CHECK(inst[i]->start_playback_stream());
			}
		}
		CHECK(mixer.update_spatializer_parameters()); // physics tick
		std::vector<const AudioFrame *> ptrs(V);
		for (int i = 0; i < V; i++) {
			ptrs[i] = &src[(size_t)i * F];
		}
		CHECK(mixer.mix(F, ptrs, bus.data())); // audio mix step
		CHECK(fwrite(bus.data(), sizeof(AudioFrame), bus.size(), fo) == bus.size());
	}
	// the frame count may not change under a running mixer (audio_spatializer.cpp:336-338)
	std::vector<const AudioFrame *> ptrs(V, src.data());
	CHECK(!mixer.mix(F + 1, ptrs, bus.data()));
	auto pd = std::dynamic_pointer_cast<SpatializerPlaybackData3D>(inst[0]->playbacks[0]);
	CHECK(pd);
	if (mode_b) {
		auto prm = inst[0]->get_spatializer_parameters();
		CHECK(prm && pd->get_prev_mix_volume(0).x == prm->get_mix_volumes()[0].x); // :608 prev <- volumes[channel]
	}
	fclose(fo);
	fclose(fi);
	printf("scene ok\n");
	return 0;
}

int main(int argc, char **argv) {
	if (argc >= 2 && !strcmp(argv[1], "validate")) {
		return validate();
	}
	if (argc >= 4 && !strcmp(argv[1], "scene")) {
		return scene(argv[2], argv[3]);
	}
	fprintf(stderr, "usage: host_test validate | scene in.bin out.bin\n");
	return 64;
}
