/*
 * godot_lite_audio.h — stand-ins for upstream servers/audio/{audio_server,audio_stream,audio_effect,
 * audio_filter_sw,effects/audio_effect_filter}.h.  TEST INFRASTRUCTURE ONLY (see godot_lite_core.h).
 *
 * Upstream pin: godotengine/godot 4.x (the reference's example project declares feature "4.6",
 * examples/godot-gd-spatializer/project.godot:19; no commit hash is pinned anywhere in the reference).
 * AudioFrame, AudioFilterSW(::Processor), AudioEffectFilter(Instance) and the AudioServer mix step are
 * restated from upstream as recalled — they are third-party code that is not
 * under /root/reference.  Everything the module itself computes comes from the module's own sources.
 */
#pragma once

#include "godot_lite_core.h"

#include <memory>

/* ---- AudioFrame (core/math/audio_frame.h) ------------------------------------------------------------ */
struct AudioFrame {
	float left = 0, right = 0;
	_ALWAYS_INLINE_ AudioFrame() {}
	_ALWAYS_INLINE_ AudioFrame(float p_left, float p_right) :
			left(p_left), right(p_right) {}
	_ALWAYS_INLINE_ AudioFrame(const Vector2 &p_v2) :
			left(p_v2.x), right(p_v2.y) {}
	_ALWAYS_INLINE_ operator Vector2() const { return Vector2(left, right); }
	_ALWAYS_INLINE_ AudioFrame operator+(const AudioFrame &p_frame) const { return AudioFrame(left + p_frame.left, right + p_frame.right); }
	_ALWAYS_INLINE_ AudioFrame operator-(const AudioFrame &p_frame) const { return AudioFrame(left - p_frame.left, right - p_frame.right); }
	_ALWAYS_INLINE_ AudioFrame operator*(const AudioFrame &p_frame) const { return AudioFrame(left * p_frame.left, right * p_frame.right); }
	_ALWAYS_INLINE_ AudioFrame operator*(float p_sample) const { return AudioFrame(left * p_sample, right * p_sample); }
	_ALWAYS_INLINE_ void operator+=(const AudioFrame &p_frame) {
		left += p_frame.left;
		right += p_frame.right;
	}
	_ALWAYS_INLINE_ void operator*=(const AudioFrame &p_frame) {
		left *= p_frame.left;
		right *= p_frame.right;
	}
	_ALWAYS_INLINE_ void operator*=(float p_sample) {
		left *= p_sample;
		right *= p_sample;
	}
	_ALWAYS_INLINE_ AudioFrame lerp(const AudioFrame &p_b, float p_t) const {
		AudioFrame res = *this;
		res.left += (p_t * (p_b.left - left));
		res.right += (p_t * (p_b.right - right));
		return res;
	}
};
_ALWAYS_INLINE_ AudioFrame operator*(float p_scalar, const AudioFrame &p_frame) {
	return AudioFrame(p_frame.left * p_scalar, p_frame.right * p_scalar);
}
_ALWAYS_INLINE_ AudioFrame operator*(int32_t p_scalar, const AudioFrame &p_frame) {
	return AudioFrame(p_frame.left * p_scalar, p_frame.right * p_scalar);
}

/* ---- AudioFilterSW (servers/audio/audio_filter_sw.{h,cpp}) ---------------------------------------------- */
class AudioFilterSW {
public:
	struct Coeffs {
		float a1 = 0, a2 = 0;
		float b0 = 0, b1 = 0, b2 = 0;
	};
	enum Mode { BANDPASS, HIGHPASS, LOWPASS, NOTCH, PEAK, BANDLIMIT, LOWSHELF, HIGHSHELF };

	class Processor { /* direct form I, per-sample interpolated coefficients */
		AudioFilterSW *filter = nullptr;
		Coeffs coeffs;
		float ha1 = 0, ha2 = 0, hb1 = 0, hb2 = 0;
		Coeffs incr_coeffs;

	public:
		void set_filter(AudioFilterSW *p_filter, bool p_clear_history = true) {
			if (p_clear_history) {
				ha1 = ha2 = hb1 = hb2 = 0;
			}
			filter = p_filter;
		}
		void update_coeffs(int p_interp_buffer_len = 0) {
			if (!filter) {
				return;
			}
			if (p_interp_buffer_len) { // interpolate
				Coeffs old_coeffs = coeffs;
				filter->prepare_coefficients(&coeffs);
				incr_coeffs.a1 = (coeffs.a1 - old_coeffs.a1) / p_interp_buffer_len;
				incr_coeffs.a2 = (coeffs.a2 - old_coeffs.a2) / p_interp_buffer_len;
				incr_coeffs.b0 = (coeffs.b0 - old_coeffs.b0) / p_interp_buffer_len;
				incr_coeffs.b1 = (coeffs.b1 - old_coeffs.b1) / p_interp_buffer_len;
				incr_coeffs.b2 = (coeffs.b2 - old_coeffs.b2) / p_interp_buffer_len;
				coeffs = old_coeffs;
			} else {
				filter->prepare_coefficients(&coeffs);
			}
		}
		_ALWAYS_INLINE_ void process_one(float &p_sample) {
			float pre = p_sample;
			p_sample = (p_sample * coeffs.b0 + hb1 * coeffs.b1 + hb2 * coeffs.b2 + ha1 * coeffs.a1 + ha2 * coeffs.a2);
			ha2 = ha1;
			hb2 = hb1;
			hb1 = pre;
			ha1 = p_sample;
		}
		_ALWAYS_INLINE_ void process_one_interp(float &p_sample) {
			float pre = p_sample;
			p_sample = (p_sample * coeffs.b0 + hb1 * coeffs.b1 + hb2 * coeffs.b2 + ha1 * coeffs.a1 + ha2 * coeffs.a2);
			ha2 = ha1;
			hb2 = hb1;
			hb1 = pre;
			ha1 = p_sample;
			coeffs.b0 += incr_coeffs.b0;
			coeffs.b1 += incr_coeffs.b1;
			coeffs.b2 += incr_coeffs.b2;
			coeffs.a1 += incr_coeffs.a1;
			coeffs.a2 += incr_coeffs.a2;
		}
		/* godot-lite only: state read-out for the test harness */
		const Coeffs &gl_coeffs() const { return coeffs; }
		void gl_history(float out[4]) const {
			out[0] = ha1;
			out[1] = ha2;
			out[2] = hb1;
			out[3] = hb2;
		}
	};

private:
	float cutoff = 5000;
	float resonance = 0.5;
	float gain = 1.0;
	float sampling_rate = 44100;
	int stages = 1;
	Mode mode = LOWPASS;

public:
	void set_mode(Mode p_mode) { mode = p_mode; }
	void set_cutoff(float p_cutoff) { cutoff = p_cutoff; }
	void set_resonance(float p_resonance) { resonance = p_resonance; }
	void set_gain(float p_gain) { gain = p_gain; }
	void set_sampling_rate(float p_srate) { sampling_rate = p_srate; }
	void set_stages(int p_stages) { stages = p_stages; }

	void prepare_coefficients(Coeffs *p_coeffs) {
		int sr_limit = (sampling_rate / 2) + 512;

		double final_cutoff = (cutoff > sr_limit) ? sr_limit : cutoff;
		if (final_cutoff < 1) {
			final_cutoff = 1; // don't allow less than this
		}

		double omega = Math::TAU * final_cutoff / sampling_rate;

		double sin_v = Math::sin(omega);
		double cos_v = Math::cos(omega);

		double Q = resonance;
		if (Q <= 0.0) {
			Q = 0.0001;
		}

		if (mode == BANDPASS) {
			Q *= 2.0;
		} else if (mode == PEAK) {
			Q *= 3.0;
		}

		double tmpgain = gain;

		if (tmpgain < 0.001) {
			tmpgain = 0.001;
		}

		if (stages > 1) {
			Q = (Q > 1.0 ? Math::pow(Q, 1.0 / stages) : Q);
			tmpgain = Math::pow(tmpgain, 1.0 / (stages + 1));
		}
		double alpha = sin_v / (2 * Q);

		double a0 = 1.0 + alpha;

		switch (mode) {
			case LOWPASS: {
				p_coeffs->b0 = (1.0 - cos_v) / 2.0;
				p_coeffs->b1 = 1.0 - cos_v;
				p_coeffs->b2 = (1.0 - cos_v) / 2.0;
				p_coeffs->a1 = -2.0 * cos_v;
				p_coeffs->a2 = 1.0 - alpha;
			} break;
			case HIGHPASS: {
				p_coeffs->b0 = (1.0 + cos_v) / 2.0;
				p_coeffs->b1 = -(1.0 + cos_v);
				p_coeffs->b2 = (1.0 + cos_v) / 2.0;
				p_coeffs->a1 = -2.0 * cos_v;
				p_coeffs->a2 = 1.0 - alpha;
			} break;
			case BANDPASS: {
				p_coeffs->b0 = alpha * sqrt(Q + 1);
				p_coeffs->b1 = 0.0;
				p_coeffs->b2 = -alpha * sqrt(Q + 1);
				p_coeffs->a1 = -2.0 * cos_v;
				p_coeffs->a2 = 1.0 - alpha;
			} break;
			case NOTCH: {
				p_coeffs->b0 = 1.0;
				p_coeffs->b1 = -2.0 * cos_v;
				p_coeffs->b2 = 1.0;
				p_coeffs->a1 = -2.0 * cos_v;
				p_coeffs->a2 = 1.0 - alpha;
			} break;
			case PEAK: {
				p_coeffs->b0 = (1.0 + alpha * tmpgain);
				p_coeffs->b1 = (-2.0 * cos_v);
				p_coeffs->b2 = (1.0 - alpha * tmpgain);
				p_coeffs->a1 = -2 * cos_v;
				p_coeffs->a2 = (1 - alpha / tmpgain);
			} break;
			case BANDLIMIT: {
				// this one is extra tricky
				double hicutoff = resonance;
				double centercutoff = (cutoff + resonance) / 2.0;
				double bandwidth = (Math::log(centercutoff) - Math::log(hicutoff)) / Math::log((double)2);
				omega = Math::TAU * centercutoff / sampling_rate;
				alpha = Math::sin(omega) * sinh(Math::log((double)2) / 2 * bandwidth * omega / Math::sin(omega));
				a0 = 1 + alpha;

				p_coeffs->b0 = alpha;
				p_coeffs->b1 = 0;
				p_coeffs->b2 = -alpha;
				p_coeffs->a1 = -2 * Math::cos(omega);
				p_coeffs->a2 = 1 - alpha;
			} break;
			case LOWSHELF: {
				double tmpq = Math::sqrt(Q);
				if (tmpq <= 0) {
					tmpq = 0.001;
				}
				double beta = Math::sqrt(tmpgain) / tmpq;

				a0 = (tmpgain + 1.0) + (tmpgain - 1.0) * cos_v + beta * sin_v;
				p_coeffs->b0 = tmpgain * ((tmpgain + 1.0) - (tmpgain - 1.0) * cos_v + beta * sin_v);
				p_coeffs->b1 = 2.0 * tmpgain * ((tmpgain - 1.0) - (tmpgain + 1.0) * cos_v);
				p_coeffs->b2 = tmpgain * ((tmpgain + 1.0) - (tmpgain - 1.0) * cos_v - beta * sin_v);
				p_coeffs->a1 = -2.0 * ((tmpgain - 1.0) + (tmpgain + 1.0) * cos_v);
				p_coeffs->a2 = ((tmpgain + 1.0) + (tmpgain - 1.0) * cos_v - beta * sin_v);
			} break;
			case HIGHSHELF: {
				double tmpq = Math::sqrt(Q);
				if (tmpq <= 0) {
					tmpq = 0.001;
				}
				double beta = Math::sqrt(tmpgain) / tmpq;

				a0 = (tmpgain + 1.0) - (tmpgain - 1.0) * cos_v + beta * sin_v;
				p_coeffs->b0 = tmpgain * ((tmpgain + 1.0) + (tmpgain - 1.0) * cos_v + beta * sin_v);
				p_coeffs->b1 = -2.0 * tmpgain * ((tmpgain - 1.0) + (tmpgain + 1.0) * cos_v);
				p_coeffs->b2 = tmpgain * ((tmpgain + 1.0) + (tmpgain - 1.0) * cos_v - beta * sin_v);
				p_coeffs->a1 = 2.0 * ((tmpgain - 1.0) - (tmpgain + 1.0) * cos_v);
				p_coeffs->a2 = (tmpgain + 1.0) - (tmpgain - 1.0) * cos_v - beta * sin_v;
			} break;
		}

		p_coeffs->b0 /= a0;
		p_coeffs->b1 /= a0;
		p_coeffs->b2 /= a0;
		p_coeffs->a1 /= 0.0 - a0;
		p_coeffs->a2 /= 0.0 - a0;
	}
};

/* ---- AudioStream / AudioStreamPlayback (servers/audio/audio_stream.h) ------------------------------------- */
class AudioStreamPlayback : public RefCounted {
	GDCLASS(AudioStreamPlayback, RefCounted);

public:
	virtual void start(double p_from_pos = 0.0) {}
	virtual void stop() {}
	virtual bool is_playing() const { return false; }
	virtual int get_loop_count() const { return 0; }
	virtual double get_playback_position() const { return 0; }
	virtual void seek(double p_time) {}
	virtual int mix(AudioFrame *p_buffer, float p_rate_scale, int p_frames) { return 0; }
	virtual void tag_used_streams() {}
	virtual void set_parameter(const StringName &p_name, const Variant &p_value) {}
	bool get_is_sample() const { return false; }
};

class AudioStream : public Resource {
	GDCLASS(AudioStream, Resource);

public:
	struct Parameter {
		PropertyInfo property;
		Variant default_value;
	};
	virtual Ref<AudioStreamPlayback> instantiate_playback() { return Ref<AudioStreamPlayback>(); }
	virtual bool is_monophonic() const { return false; }
	virtual void get_parameter_list(List<Parameter> *r_parameters) {}
};

/* ---- AudioEffect / AudioEffectFilter (servers/audio/audio_effect.h, effects/audio_effect_filter.{h,cpp}) --- */
class AudioEffectInstance : public RefCounted {
	GDCLASS(AudioEffectInstance, RefCounted);

public:
	virtual void process(const AudioFrame *p_src_frames, AudioFrame *p_dst_frames, int p_frame_count) {}
};
class AudioEffect : public Resource {
	GDCLASS(AudioEffect, Resource);

public:
	virtual Ref<AudioEffectInstance> instantiate() { return Ref<AudioEffectInstance>(); }
};

class AudioServer;
class AudioEffectFilter;
class AudioEffectFilterInstance : public AudioEffectInstance {
	GDCLASS(AudioEffectFilterInstance, AudioEffectInstance);
	friend class AudioEffectFilter;
	Ref<AudioEffectFilter> base;
	AudioFilterSW filter;
	AudioFilterSW::Processor filter_process[2][4];

	template <int S>
	void _process_filter(const AudioFrame *p_src_frames, AudioFrame *p_dst_frames, int p_frame_count) {
		for (int i = 0; i < p_frame_count; i++) {
			float f = p_src_frames[i].left;
			filter_process[0][0].process_one(f);
			if constexpr (S > 1) {
				filter_process[0][1].process_one(f);
			}
			if constexpr (S > 2) {
				filter_process[0][2].process_one(f);
			}
			if constexpr (S > 3) {
				filter_process[0][3].process_one(f);
			}
			p_dst_frames[i].left = f;
		}
		for (int i = 0; i < p_frame_count; i++) {
			float f = p_src_frames[i].right;
			filter_process[1][0].process_one(f);
			if constexpr (S > 1) {
				filter_process[1][1].process_one(f);
			}
			if constexpr (S > 2) {
				filter_process[1][2].process_one(f);
			}
			if constexpr (S > 3) {
				filter_process[1][3].process_one(f);
			}
			p_dst_frames[i].right = f;
		}
	}

public:
	AudioEffectFilterInstance() {
		for (int i = 0; i < 2; i++) {
			for (int j = 0; j < 4; j++) {
				filter_process[i][j].set_filter(&filter);
			}
		}
	}
	virtual void process(const AudioFrame *p_src_frames, AudioFrame *p_dst_frames, int p_frame_count) override;
	const AudioFilterSW::Processor &gl_processor(int side, int stage) const { return filter_process[side][stage]; }
};

class AudioEffectFilter : public AudioEffect {
	GDCLASS(AudioEffectFilter, AudioEffect);

public:
	enum FilterDB { FILTER_6DB, FILTER_12DB, FILTER_18DB, FILTER_24DB };
	friend class AudioEffectFilterInstance;
	AudioFilterSW::Mode mode;
	float cutoff = 2000;
	float resonance = 0.5;
	float gain = 1.0;
	FilterDB db = FILTER_6DB;

	void set_cutoff(float p_freq) { cutoff = p_freq; }
	float get_cutoff() const { return cutoff; }
	void set_resonance(float p_amount) { resonance = p_amount; }
	float get_resonance() const { return resonance; }
	void set_gain(float p_amount) { gain = p_amount; }
	float get_gain() const { return gain; }
	void set_db(FilterDB p_db) { db = p_db; }
	FilterDB get_db() const { return db; }

	virtual Ref<AudioEffectInstance> instantiate() override {
		Ref<AudioEffectFilterInstance> ins;
		ins.instantiate();
		ins->base = Ref<AudioEffectFilter>(this);
		return ins;
	}
	virtual Ref<Resource> duplicate(bool p_subresources = false) const override {
		Ref<AudioEffectFilter> r;
		r.instantiate();
		r->mode = mode;
		r->cutoff = cutoff;
		r->resonance = resonance;
		r->gain = gain;
		r->db = db;
		return r;
	}
	AudioEffectFilter(AudioFilterSW::Mode p_mode = AudioFilterSW::LOWPASS) :
			mode(p_mode) {}
};

/* ---- AudioServer (servers/audio/audio_server.{h,cpp}): globals + the playback mix step ------------------------ */
typedef void (*AudioCallback)(void *p_userdata);

class AudioServer : public Object {
	GDCLASS(AudioServer, Object);

public:
	enum SpeakerMode {
		SPEAKER_MODE_STEREO,
		SPEAKER_SURROUND_31,
		SPEAKER_SURROUND_51,
		SPEAKER_SURROUND_71,
	};
	enum {
		MAX_CHANNELS_PER_BUS = 4,
		MAX_BUSES_PER_PLAYBACK = 6,
		LOOKAHEAD_BUFFER_SIZE = 64,
	};
	typedef ::AudioCallback AudioCallback;

	/* upstream AudioStreamPlaybackBusDetails / AudioStreamPlaybackListNode */
	struct BusDetails {
		bool bus_active[MAX_BUSES_PER_PLAYBACK] = {};
		StringName bus[MAX_BUSES_PER_PLAYBACK];
		AudioFrame volume[MAX_BUSES_PER_PLAYBACK][MAX_CHANNELS_PER_BUS];
	};
	struct PlaybackNode {
		Ref<AudioStreamPlayback> stream_playback;
		BusDetails bus_details;
		BusDetails prev_bus_details;
		AudioFrame lookahead[LOOKAHEAD_BUFFER_SIZE];
		bool paused = false;
		bool fading_out = false; /* FADE_OUT_TO_DELETION */
		int64_t order_key = 0;
	};

	/* godot-lite configuration, set by the harness */
	SpeakerMode gl_speaker_mode = SPEAKER_MODE_STEREO;
	float gl_mix_rate = 44100;
	std::vector<String> gl_bus_names{ String("Master") };
	/* upstream mixes every playback through its own 64-frame lookahead; the batched mixer (and the oracle)
	 * are defined at the bus-accumulate input, so this is off unless a test turns it on */
	bool gl_playback_lookahead = false;
	std::function<int64_t(AudioStreamPlayback *)> gl_order_key; /* mix order of the playbacks (sum order only) */
	std::list<std::unique_ptr<PlaybackNode>> gl_playbacks;
	std::vector<std::pair<AudioCallback, void *>> gl_listener_changed;
	std::vector<std::vector<std::vector<AudioFrame>>> gl_bus_buffers; /* [bus][channel][frame] */

	static AudioServer *&singleton_ptr() {
		static AudioServer *s = nullptr;
		return s;
	}
	static AudioServer *get_singleton() { return singleton_ptr(); }

	SpeakerMode get_speaker_mode() const { return gl_speaker_mode; }
	int get_channel_count() const { return (int)gl_speaker_mode + 1; }
	float get_mix_rate() const { return gl_mix_rate; }
	int get_bus_count() const { return (int)gl_bus_names.size(); }
	String get_bus_name(int p_bus) const { return gl_bus_names[(size_t)p_bus]; }

	void add_listener_changed_callback(AudioCallback p_callback, void *p_userdata) { gl_listener_changed.emplace_back(p_callback, p_userdata); }
	void remove_listener_changed_callback(AudioCallback p_callback, void *p_userdata) {
		for (size_t i = 0; i < gl_listener_changed.size(); i++) {
			if (gl_listener_changed[i].first == p_callback && gl_listener_changed[i].second == p_userdata) {
				gl_listener_changed.erase(gl_listener_changed.begin() + i);
				return;
			}
		}
	}

	PlaybackNode *gl_find(const Ref<AudioStreamPlayback> &p_playback) {
		for (auto &n : gl_playbacks) {
			if (n->stream_playback == p_playback) {
				return n.get();
			}
		}
		return nullptr;
	}
	static void gl_fill_details(BusDetails &d, const HashMap<StringName, Vector<AudioFrame>> &p_bus_volumes) {
		d = BusDetails();
		int idx = 0;
		for (const KeyValue<StringName, Vector<AudioFrame>> &pair : p_bus_volumes) {
			if (pair.value.size() < MAX_CHANNELS_PER_BUS) {
				continue;
			}
			if (idx >= MAX_BUSES_PER_PLAYBACK) {
				break;
			}
			d.bus_active[idx] = true;
			d.bus[idx] = pair.key;
			for (int c = 0; c < MAX_CHANNELS_PER_BUS; c++) {
				d.volume[idx][c] = pair.value[c];
			}
			idx++;
		}
	}
	void start_playback_stream(Ref<AudioStreamPlayback> p_playback, const HashMap<StringName, Vector<AudioFrame>> &p_bus_volumes,
			float p_start_time = 0, float p_pitch_scale = 1, float p_highshelf_gain = 0, float p_attenuation_cutoff_hz = 0) {
		ERR_FAIL_COND(p_playback.is_null());
		std::unique_ptr<PlaybackNode> n(new PlaybackNode());
		n->stream_playback = p_playback;
		n->stream_playback->start(p_start_time);
		gl_fill_details(n->bus_details, p_bus_volumes);
		n->order_key = gl_order_key ? gl_order_key(p_playback.ptr()) : (int64_t)gl_playbacks.size();
		gl_playbacks.push_back(std::move(n));
	}
	void stop_playback_stream(Ref<AudioStreamPlayback> p_playback) {
		if (PlaybackNode *n = gl_find(p_playback)) {
			n->fading_out = true;
		}
	}
	void set_playback_bus_volumes_linear(Ref<AudioStreamPlayback> p_playback, const HashMap<StringName, Vector<AudioFrame>> &p_bus_volumes) {
		ERR_FAIL_COND(p_bus_volumes.size() > MAX_BUSES_PER_PLAYBACK);
		if (PlaybackNode *n = gl_find(p_playback)) {
			gl_fill_details(n->bus_details, p_bus_volumes);
		}
	}
	void set_playback_paused(Ref<AudioStreamPlayback> p_playback, bool p_paused) {
		if (PlaybackNode *n = gl_find(p_playback)) {
			n->paused = p_paused;
		}
	}
	bool is_playback_paused(Ref<AudioStreamPlayback> p_playback) {
		PlaybackNode *n = gl_find(p_playback);
		return n ? n->paused : false;
	}
	int get_bus_index(const StringName &p_bus_name) const { /* upstream: -1 when there is no such bus */
		for (size_t i = 0; i < gl_bus_names.size(); i++) {
			if (gl_bus_names[i] == String(p_bus_name)) {
				return (int)i;
			}
		}
		return -1;
	}
	int gl_bus_index(const StringName &p_name) const { /* thread_find_bus_index: unknown => Master */
		for (size_t i = 0; i < gl_bus_names.size(); i++) {
			if (gl_bus_names[i] == String(p_name)) {
				return (int)i;
			}
		}
		return 0;
	}

	/* upstream AudioServer::_mix_step_for_channel with p_highshelf_gain == 0 (the module never sets it) */
	static void _mix_step_for_channel(AudioFrame *p_out_buf, AudioFrame *p_source_buf, AudioFrame p_vol_start, AudioFrame p_vol_final, unsigned int buffer_size) {
		for (unsigned int frame_idx = 0; frame_idx < buffer_size; frame_idx++) {
			float lerp_param = (float)frame_idx / buffer_size;
			p_out_buf[frame_idx] += (p_vol_final * lerp_param + (1 - lerp_param) * p_vol_start) * p_source_buf[frame_idx];
		}
	}

	/* upstream AudioServer::_mix_step, playback part: every playback is mixed into its active buses with
	 * volumes ramped from the previous step's (looked up by bus name; absent => 0), buses that
	 * disappeared fade out to 0, then prev <- current. */
	void gl_mix_step(int buffer_size) {
		const int channel_count = get_channel_count();
		gl_bus_buffers.assign(gl_bus_names.size(), std::vector<std::vector<AudioFrame>>((size_t)channel_count, std::vector<AudioFrame>((size_t)buffer_size)));
		std::vector<PlaybackNode *> order;
		for (auto &n : gl_playbacks) {
			order.push_back(n.get());
		}
		std::stable_sort(order.begin(), order.end(), [](PlaybackNode *a, PlaybackNode *b) { return a->order_key < b->order_key; });
		std::vector<AudioFrame> mix_buffer((size_t)buffer_size + LOOKAHEAD_BUFFER_SIZE);
		std::vector<PlaybackNode *> to_delete;
		for (PlaybackNode *playback : order) {
			if (playback->paused) {
				continue;
			}
			const bool fading_out = playback->fading_out;
			AudioFrame *buf = mix_buffer.data();
			std::fill(mix_buffer.begin(), mix_buffer.end(), AudioFrame(0, 0));
			AudioFrame *mix_at = buf;
			if (gl_playback_lookahead) {
				for (int i = 0; i < LOOKAHEAD_BUFFER_SIZE; i++) {
					buf[i] = playback->lookahead[i];
				}
				mix_at = &buf[LOOKAHEAD_BUFFER_SIZE];
			}
			unsigned int mixed_frames = (unsigned int)playback->stream_playback->mix(mix_at, 1.0f, buffer_size);
			bool awaiting_deletion = false;
			if (mixed_frames != (unsigned int)buffer_size) {
				float fadeout_base = 0.94;
				float fadeout_coefficient = 1;
				for (unsigned int idx = mixed_frames; idx < (unsigned int)buffer_size; idx++) {
					fadeout_coefficient *= fadeout_base;
					buf[idx] *= fadeout_coefficient;
				}
				awaiting_deletion = true;
			} else if (gl_playback_lookahead) {
				for (int i = 0; i < LOOKAHEAD_BUFFER_SIZE; i++) {
					playback->lookahead[i] = buf[buffer_size + i];
				}
			}
			BusDetails bus_details = playback->bus_details;
			for (int idx = 0; idx < MAX_BUSES_PER_PLAYBACK; idx++) {
				if (!bus_details.bus_active[idx]) {
					continue;
				}
				int bus_idx = gl_bus_index(bus_details.bus[idx]);
				int prev_bus_idx = -1;
				for (int search_idx = 0; search_idx < MAX_BUSES_PER_PLAYBACK; search_idx++) {
					if (!playback->prev_bus_details.bus_active[search_idx]) {
						continue;
					}
					if (playback->prev_bus_details.bus[search_idx] == bus_details.bus[idx]) {
						prev_bus_idx = search_idx;
					}
				}
				for (int channel_idx = 0; channel_idx < channel_count; channel_idx++) {
					AudioFrame *channel_buf = gl_bus_buffers[(size_t)bus_idx][(size_t)channel_idx].data();
					if (fading_out) {
						bus_details.volume[idx][channel_idx] = AudioFrame(0, 0);
					}
					AudioFrame channel_vol = bus_details.volume[idx][channel_idx];
					AudioFrame prev_channel_vol = AudioFrame(0, 0);
					if (prev_bus_idx != -1) {
						prev_channel_vol = playback->prev_bus_details.volume[prev_bus_idx][channel_idx];
					}
					_mix_step_for_channel(channel_buf, buf, prev_channel_vol, channel_vol, (unsigned int)buffer_size);
				}
			}
			for (int idx = 0; idx < MAX_BUSES_PER_PLAYBACK; idx++) {
				if (!playback->prev_bus_details.bus_active[idx]) {
					continue;
				}
				int bus_idx = gl_bus_index(playback->prev_bus_details.bus[idx]);
				int current_bus_idx = -1;
				for (int search_idx = 0; search_idx < MAX_BUSES_PER_PLAYBACK; search_idx++) {
					if (bus_details.bus_active[search_idx] && bus_details.bus[search_idx] == playback->prev_bus_details.bus[idx]) {
						current_bus_idx = search_idx;
					}
				}
				if (current_bus_idx != -1) {
					continue; // handled above
				}
				for (int channel_idx = 0; channel_idx < channel_count; channel_idx++) {
					AudioFrame *channel_buf = gl_bus_buffers[(size_t)bus_idx][(size_t)channel_idx].data();
					AudioFrame prev_channel_vol = playback->prev_bus_details.volume[idx][channel_idx];
					_mix_step_for_channel(channel_buf, buf, prev_channel_vol, AudioFrame(0, 0), (unsigned int)buffer_size); // fade out to silence
				}
			}
			playback->prev_bus_details = bus_details;
			if (awaiting_deletion || fading_out) {
				to_delete.push_back(playback);
			}
		}
		for (PlaybackNode *d : to_delete) {
			gl_playbacks.remove_if([d](const std::unique_ptr<PlaybackNode> &n) { return n.get() == d; });
		}
	}
};

inline void AudioEffectFilterInstance::process(const AudioFrame *p_src_frames, AudioFrame *p_dst_frames, int p_frame_count) {
	filter.set_cutoff(base->cutoff);
	filter.set_gain(base->gain);
	filter.set_resonance(base->resonance);
	filter.set_mode(base->mode);
	int stages = int(base->db) + 1;
	filter.set_stages(stages);
	filter.set_sampling_rate(AudioServer::get_singleton()->get_mix_rate());

	for (int i = 0; i < 2; i++) {
		for (int c = 0; c < 4; c++) {
			filter_process[i][c].update_coeffs();
		}
	}

	if (stages == 1) {
		_process_filter<1>(p_src_frames, p_dst_frames, p_frame_count);
	} else if (stages == 2) {
		_process_filter<2>(p_src_frames, p_dst_frames, p_frame_count);
	} else if (stages == 3) {
		_process_filter<3>(p_src_frames, p_dst_frames, p_frame_count);
	} else if (stages == 4) {
		_process_filter<4>(p_src_frames, p_dst_frames, p_frame_count);
	}
}
