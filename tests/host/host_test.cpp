// host_test — exercises the C++ host mirror (godot-audio-spatializer_b200/host) the way a Godot-side caller would.
//   host_test validate                 : setter / parameter validation (reference ERR_FAIL_* behaviour); no GPU needed
//   host_test scene <in.bin> <out.bin> : plays a scene read from in.bin through BatchMixer on cuda:0, writes the bus buffers
//   host_test stream <in.bin> <out.bin>: the same through BatchMixer::mix_streams (voice lifecycle on the device): streams of
//                                        given lengths, per block the bus buffers and the number of playbacks still alive
//   host_test percall                  : the per-call virtuals (process_frames / mix_channel) and the refusal of overriders
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

// in.bin : int32 {V, F, blocks, speaker_mode, num_buses, mix_channel_mode}, int32 length[V], then per block gas_emitter[V], float[V][F][2]
// out.bin: per block float[num_buses][channels][F][2], int32 alive_after
static int stream(const char *in_path, const char *out_path) {
	FILE *fi = fopen(in_path, "rb");
	CHECK(fi);
	int32_t h[6];
	CHECK(fread(h, sizeof(int32_t), 6, fi) == 6);
	const int V = h[0], F = h[1], blocks = h[2], speaker_mode = h[3], num_buses = h[4], mode_b = h[5];
	std::vector<int32_t> length(V);
	CHECK(fread(length.data(), sizeof(int32_t), V, fi) == (size_t)V);
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
	std::vector<Ref<AudioSpatializerInstance>> inst(V);
	std::vector<Ref<SpatializerPlaybackData>> pb(V);
	std::vector<int64_t> pos(V, 0);
	for (int i = 0; i < V; i++) {
		inst[i] = mixer.instantiate(spat);
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
			inst[i]->set_global_transform(t);
			inst[i]->set_volume_db(em[i].volume_db);
			inst[i]->set_max_db(em[i].max_db);
			if (b == 0) {
				pb[i] = inst[i]->start_playback_stream();
				CHECK(pb[i]);
			}
		}
		CHECK(mixer.update_spatializer_parameters());
		// the live playbacks in the mixer's order, with what their streams still deliver
		auto order = mixer.playback_order();
		std::vector<const AudioFrame *> ptrs;
		std::vector<int> counts;
		for (auto &d : order) {
			int i = 0;
			while (i < V && pb[i] != d) {
				i++;
			}
			CHECK(i < V);
			const int64_t left = length[i] - pos[i];
			const int n = (int)(left < 0 ? 0 : (left > F ? F : left));
			ptrs.push_back(n > 0 ? &src[(size_t)i * F] : nullptr);
			counts.push_back(n);
			pos[i] += n;
		}
		std::vector<Ref<SpatializerPlaybackData>> finished;
		CHECK(mixer.mix_streams(F, ptrs, counts, bus.data(), &finished));
		for (auto &d : finished) { // a finished playback is no longer in any instance's list
			for (int i = 0; i < V; i++) {
				if (pb[i] == d) {
					CHECK(!inst[i]->is_playback_active());
				}
			}
		}
		const int32_t alive = (int32_t)mixer.playback_order().size();
		CHECK(fwrite(bus.data(), sizeof(AudioFrame), bus.size(), fo) == bus.size());
		CHECK(fwrite(&alive, sizeof(alive), 1, fo) == 1);
	}
	fclose(fo);
	fclose(fi);
	printf("stream ok\n");
	return 0;
}

// an instance whose subclass brings its own per-voice DSP: BatchMixer must refuse to batch it
class CustomDspInstance : public FixedInstance {
public:
	bool uses_builtin_dsp() const override { return false; }
	void process_frames(const Ref<SpatializerParameters> &p, const Ref<SpatializerPlaybackData> &d, AudioFrame *out, const AudioFrame *src, int n) override {
		AudioSpatializerInstance::process_frames(p, d, out, src, n); // built-in part first
		for (int i = 0; i < n; i++) {
			out[i].l *= 0.5f;
		}
	}
};
class CustomDspSpatializer : public AudioSpatializer3D {
public:
	Ref<AudioSpatializerInstance> instantiate() override {
		auto i = std::make_shared<CustomDspInstance>();
		i->base = shared_from_this();
		i->mix_channel_mode = get_mix_channel_mode();
		return i;
	}
};

static int percall() {
	BatchMixerConfig cfg;
	cfg.max_instances = 4;
	cfg.max_voices = 4;
	cfg.max_frames = 64;
	cfg.speaker_mode = GAS_SPEAKER_SURROUND_71;
	BatchMixer mixer(cfg);
	if (!mixer.ok()) {
		fprintf(stderr, "no device: %s\n", last_error().c_str());
		return 2;
	}
	const int F = 50; // any frame count, like the reference's virtuals
	std::vector<AudioFrame> src(F), out(F);
	for (int i = 0; i < F; i++) {
		src[i] = AudioFrame{ 0.01f * (i + 1), -0.02f * (i + 1) };
	}
	// Mode A, filter gain 0 (< 0.001): process_frames is a copy (audio_spatializer_3d.cpp:531-534) and stores the prev volume (:537-551)
	auto fa = std::make_shared<FixedSpatializer>();
	auto ia = mixer.instantiate(fa);
	auto pa = ia->start_playback_stream();
	CHECK(ia && pa && mixer.update_spatializer_parameters());
	ia->process_frames(nullptr, pa, out.data(), src.data(), F);
	for (int i = 0; i < F; i++) {
		CHECK(out[i].l == src[i].l && out[i].r == src[i].r);
	}
	auto pda = std::dynamic_pointer_cast<SpatializerPlaybackData3D>(pa);
	CHECK(pda && pda->get_prev_mix_volume(0).x == 0.25f && pda->get_prev_mix_volume(0).y == 0.5f);
	// Mode B: mix_channel ramps from the previous volume (0 on the first call) to volumes[channel] with t = i / F (:589-604)
	auto fb = std::make_shared<FixedSpatializer>();
	fb->set_mix_channel_mode(true);
	auto ib = mixer.instantiate(fb);
	auto pbk = ib->start_playback_stream();
	CHECK(ib && pbk && mixer.update_spatializer_parameters());
	ib->mix_channel(nullptr, pbk, 2, out.data(), src.data(), F);
	for (int i = 0; i < F; i++) {
		const float t = (float)i / F;
		const float vl = 0.25f * t + (1 - t) * 0.0f, vr = 0.5f * t + (1 - t) * 0.0f;
		CHECK(out[i].l == vl * src[i].l && out[i].r == vr * src[i].r);
	}
	ib->mix_channel(nullptr, pbk, 2, out.data(), src.data(), F); // second call: prev == new, the ramp is flat
	CHECK(fabsf(out[7].l - 0.25f * src[7].l) < 1e-7f && fabsf(out[7].r - 0.5f * src[7].r) < 1e-7f);
	// wrong playback data => no-op with an error (ERR_FAIL_COND_MSG, :493-494)
	out[0].l = 123.f;
	ib->mix_channel(nullptr, pa, 0, out.data(), src.data(), F);
	CHECK(out[0].l == 123.f && last_error().find("Unexpected SpatializerPlaybackData") != std::string::npos);
	ib->mix_channel(nullptr, pbk, 4, out.data(), src.data(), F); // ERR_FAIL_INDEX of get_filter_processor (:888)
	CHECK(out[0].l == 123.f);
	// the batched step still runs with built-in instances only ...
	std::vector<AudioFrame> bus((size_t)cfg.num_buses * 4 * 64), blk(64);
	std::vector<const AudioFrame *> ptrs(2, blk.data());
	CHECK(mixer.mix(64, ptrs, bus.data()));
	// ... and refuses an instance that brings its own per-voice DSP
	auto fc = std::make_shared<CustomDspSpatializer>();
	auto ic = mixer.instantiate(fc);
	auto pc = ic->start_playback_stream();
	CHECK(ic && pc && mixer.update_spatializer_parameters());
	ptrs.push_back(blk.data());
	CHECK(!mixer.mix(64, ptrs, bus.data()) && last_error().find("uses_builtin_dsp") != std::string::npos);
	ic->process_frames(nullptr, pc, out.data(), src.data(), F); // its override still reaches the built-in part per call
	CHECK(out[3].l == 0.5f * src[3].l && out[3].r == src[3].r);
	ic->stop_playback_stream(pc);
	ptrs.pop_back();
	CHECK(mixer.mix(64, ptrs, bus.data()));
	printf("percall ok\n");
	return 0;
}

int main(int argc, char **argv) {
	if (argc >= 2 && !strcmp(argv[1], "validate")) {
		return validate();
	}
	if (argc >= 4 && !strcmp(argv[1], "scene")) {
		return scene(argv[2], argv[3]);
	}
	if (argc >= 4 && !strcmp(argv[1], "stream")) {
		return stream(argv[2], argv[3]);
	}
	if (argc >= 2 && !strcmp(argv[1], "percall")) {
		return percall();
	}
	fprintf(stderr, "usage: host_test validate | scene in.bin out.bin | stream in.bin out.bin | percall\n");
	return 64;
}
