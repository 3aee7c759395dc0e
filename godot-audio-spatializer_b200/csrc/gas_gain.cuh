// gas_gain.cuh — calculate_spatialization for one emitter on a few lanes (see gas_gain.cu for the stand-alone kernel; the
// control warps of the step kernel, gas_mix_stream.cu, run the same code beside the streaming of the previous block).
// Include only from translation units compiled with -fmad=false.
#pragma once

#include "gas_internal.h"

#define CMP_EPSILON 0.00001

namespace gasgain {


struct V3 {
	float x, y, z;
};
static __device__ __forceinline__ float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static __device__ __forceinline__ float len3(V3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
static __device__ __forceinline__ V3 sub3(V3 a, V3 b) { return V3{ a.x - b.x, a.y - b.y, a.z - b.z }; }
static __device__ __forceinline__ V3 mul3(V3 a, float s) { return V3{ a.x * s, a.y * s, a.z * s }; }
// out of line (like the double transcendentals below): the kernel runs once per thread straight through ~70 KB of code,
// and its instruction-fetch stalls shrink with every call site that shares one copy
static __device__ __noinline__ V3 norm3(V3 a) { // upstream Vector3::normalized
	float l2 = a.x * a.x + a.y * a.y + a.z * a.z;
	if (l2 == 0.0f) {
		return V3{ 0.f, 0.f, 0.f };
	}
	float l = sqrtf(l2);
	return V3{ a.x / l, a.y / l, a.z / l };
}

struct Xf { // upstream Transform3D: Basis rows + origin
	float m[3][3];
	V3 o;
};
static __device__ __forceinline__ V3 col(const Xf &t, int c) { return V3{ t.m[0][c], t.m[1][c], t.m[2][c] }; }
static __device__ __forceinline__ void set_col(Xf &t, int c, V3 v) {
	t.m[0][c] = v.x;
	t.m[1][c] = v.y;
	t.m[2][c] = v.z;
}
static __device__ void orthonormalize(Xf &t) { // upstream Basis::orthonormalize (Gram-Schmidt)
	V3 x = col(t, 0), y = col(t, 1), z = col(t, 2);
	x = norm3(x);
	y = sub3(y, mul3(x, dot3(x, y)));
	y = norm3(y);
	z = sub3(sub3(z, mul3(x, dot3(x, z))), mul3(y, dot3(y, z)));
	z = norm3(z);
	set_col(t, 0, x);
	set_col(t, 1, y);
	set_col(t, 2, z);
}
static __device__ __forceinline__ V3 bxform(const Xf &t, V3 v) {
	return V3{ t.m[0][0] * v.x + t.m[0][1] * v.y + t.m[0][2] * v.z,
		t.m[1][0] * v.x + t.m[1][1] * v.y + t.m[1][2] * v.z,
		t.m[2][0] * v.x + t.m[2][1] * v.y + t.m[2][2] * v.z };
}
static __device__ __forceinline__ V3 bxform_inv(const Xf &t, V3 v) {
	return V3{ t.m[0][0] * v.x + t.m[1][0] * v.y + t.m[2][0] * v.z,
		t.m[0][1] * v.x + t.m[1][1] * v.y + t.m[2][1] * v.z,
		t.m[0][2] * v.x + t.m[1][2] * v.y + t.m[2][2] * v.z };
}
static __device__ void affine_invert(Xf &t) { // upstream Transform3D::affine_invert
#define CF(r1, c1, r2, c2) (t.m[r1][c1] * t.m[r2][c2] - t.m[r1][c2] * t.m[r2][c1])
	float co0 = CF(1, 1, 2, 2), co1 = CF(1, 2, 2, 0), co2 = CF(1, 0, 2, 1);
	float det = t.m[0][0] * co0 + t.m[0][1] * co1 + t.m[0][2] * co2;
	float s = 1.0f / det;
	float n[3][3];
	n[0][0] = co0 * s;
	n[0][1] = CF(0, 2, 2, 1) * s;
	n[0][2] = CF(0, 1, 1, 2) * s;
	n[1][0] = co1 * s;
	n[1][1] = CF(0, 0, 2, 2) * s;
	n[1][2] = CF(0, 2, 1, 0) * s;
	n[2][0] = co2 * s;
	n[2][1] = CF(0, 1, 2, 0) * s;
	n[2][2] = CF(0, 0, 1, 1) * s;
#undef CF
	for (int i = 0; i < 3; i++) {
		for (int j = 0; j < 3; j++) {
			t.m[i][j] = n[i][j];
		}
	}
	t.o = bxform(t, V3{ -t.o.x, -t.o.y, -t.o.z });
}
static __device__ __forceinline__ V3 xform(const Xf &t, V3 v) {
	V3 r = bxform(t, v);
	return V3{ r.x + t.o.x, r.y + t.o.y, r.z + t.o.z };
}
static __device__ __forceinline__ Xf load_xf12(const float *m) { // 9 basis floats (rows) + origin
	Xf t;
	t.m[0][0] = m[0]; t.m[0][1] = m[1]; t.m[0][2] = m[2];
	t.m[1][0] = m[3]; t.m[1][1] = m[4]; t.m[1][2] = m[5];
	t.m[2][0] = m[6]; t.m[2][1] = m[7]; t.m[2][2] = m[8];
	t.o = V3{ m[9], m[10], m[11] };
	return t;
}
static __device__ __forceinline__ void store_xf12(float *m, const Xf &t) {
	m[0] = t.m[0][0]; m[1] = t.m[0][1]; m[2] = t.m[0][2];
	m[3] = t.m[1][0]; m[4] = t.m[1][1]; m[5] = t.m[1][2];
	m[6] = t.m[2][0]; m[7] = t.m[2][1]; m[8] = t.m[2][2];
	m[9] = t.o.x; m[10] = t.o.y; m[11] = t.o.z;
}
static __device__ Xf load_xf(const gas_listener &l) {
	Xf t;
	for (int i = 0; i < 3; i++) {
		for (int j = 0; j < 3; j++) {
			t.m[i][j] = l.basis[i * 3 + j];
		}
	}
	t.o = V3{ l.origin[0], l.origin[1], l.origin[2] };
	return t;
}

// upstream Math::db_to_linear(float) / linear_to_db(double)
static __device__ __noinline__ float db_to_linear_f(float db) {
	float a = db * (float)0.11512925464970228420089957273422;
	return (float)exp((double)a);
}
static __device__ __noinline__ double log_d(double x) { return log(x); }
static __device__ __noinline__ double pow_d(double a, double b) { return pow(a, b); }
static __device__ __noinline__ double acos_d(double x) { return acos(x); }
static __device__ __noinline__ double log2_d(double x) { return log2(x); }
static __device__ __forceinline__ double linear_to_db_d(double lin) { return log_d(lin) * 8.6858896380650365530225783783321; }

// reference audio_spatializer_3d.cpp:123-151
static __device__ float attenuation_db(const gas_spatializer &s, float volume_db, float max_db, float dist) {
	float att = 0.f;
	switch (s.attenuation_model) {
		case GAS_ATTENUATION_INVERSE_DISTANCE:
			att = (float)linear_to_db_d(1.0 / ((double)(dist / s.unit_size) + CMP_EPSILON));
			break;
		case GAS_ATTENUATION_INVERSE_SQUARE_DISTANCE: {
			float d = dist / s.unit_size;
			d *= d;
			att = (float)linear_to_db_d(1.0 / ((double)d + CMP_EPSILON));
		} break;
		case GAS_ATTENUATION_LOGARITHMIC:
			att = (float)(-20.0 * log_d((double)(dist / s.unit_size) + CMP_EPSILON));
			break;
		default:
			break;
	}
	att += volume_db;
	if (att > max_db) {
		att = max_db;
	}
	return att;
}

// reference audio_spatializer_3d.cpp:57-98 + :903-938.  Lane `l` of the emitter's NL-lane group (mask gm,
// first lane gbase) owns speakers l, l + NL, ...; sums run in speaker order exactly like the reference loop.
template <int NL>
static __device__ __forceinline__ void output_vol_surround(unsigned gm, int gbase, int l, const GlobalCfg &g, V3 src, float tightness, float out[4][2]) {
	constexpr int S = 8 / NL; // speakers per lane
	const int speaker_mode = g.speaker_mode;
	const int count = speaker_mode == GAS_SPEAKER_SURROUND_31 ? 3 : (speaker_mode == GAS_SPEAKER_SURROUND_51 ? 5 : (speaker_mode == GAS_SPEAKER_SURROUND_71 ? 7 : 2));
	float sq[S];
#pragma unroll
	for (int k = 0; k < S; k++) {
		const int spk = l + k * NL;
		sq[k] = 0.f;
		if (spk < count) {
			const V3 dl{ g.spk_dir[spk][0], g.spk_dir[spk][1], g.spk_dir[spk][2] };
			const float eff = g.spk_eff[spk]; // :911-915, precomputed per speaker mode
			// :929-933.  pow(x, 1) and pow(x, 2) are exact in one rounding (x, x * x), which is what a correctly rounded
			// pow returns: the default 3d_panning_strength / panning_strength (tightness 1) never pays for the general pow
			const double base1 = 1.0 + (double)dot3(dl, src);
			const double pw = tightness == 1.0f ? base1 : (tightness == 2.0f ? base1 * base1 : pow_d(base1, (double)tightness));
			const float gain = (float)(0.5 * pw / (double)eff);
			sq[k] = gain * gain;
		}
	}
	float sum = 0.f;
#pragma unroll
	for (int i = 0; i < 7; i++) {
		const float v = NL == 1 ? sq[i % S] : __shfl_sync(gm, sq[i / NL], gbase + (i % NL));
		if (i < count) {
			sum += v;
		}
	}
	float mine[S];
#pragma unroll
	for (int k = 0; k < S; k++) {
		mine[k] = sqrtf(sq[k] / sum); // :935-937
	}
	float vol[7];
#pragma unroll
	for (int i = 0; i < 7; i++) {
		const float v = NL == 1 ? mine[i % S] : __shfl_sync(gm, mine[i / NL], gbase + (i % NL));
		vol[i] = i < count ? v : 0.f;
	}
	switch (speaker_mode) {
		case GAS_SPEAKER_SURROUND_71:
			out[3][0] = vol[5];
			out[3][1] = vol[6];
		case GAS_SPEAKER_SURROUND_51:
			out[2][0] = vol[3];
			out[2][1] = vol[4];
		case GAS_SPEAKER_SURROUND_31:
			out[1][0] = vol[2];
			out[1][1] = 1.0f; // LFE — always full power (Q9)
		default:
			out[0][0] = vol[0];
			out[0][1] = vol[1];
	}
}

// reference audio_spatializer_3d.cpp:103-110
static __device__ __noinline__ void output_vol_stereo(V3 dir, float pan_strength, float out[4][2]) {
	double flatrad = sqrt((double)(dir.x * dir.x + dir.z * dir.z));
	double g = (1.0 - (double)pan_strength) * (1.0 - (double)pan_strength);
	g = g < 0.0 ? 0.0 : (g > 1.0 ? 1.0 : g);
	double f = (1.0 - g) / (1.0 + g);
	double cosx = (double)dir.x / (flatrad == 0.0 ? 1.0 : flatrad);
	cosx = cosx < -1.0 ? -1.0 : (cosx > 1.0 ? 1.0 : cosx);
	double fcosx = cosx * f;
	out[0][0] = (float)sqrt((-fcosx + 1.0) / 2.0);
	out[0][1] = (float)sqrt((fcosx + 1.0) / 2.0);
}

// reference audio_spatializer_3d.cpp:112-121
template <int NL>
static __device__ __forceinline__ void output_vol(unsigned gm, int gbase, int l, const GlobalCfg &g, const gas_spatializer &s, V3 dir, float out[4][2]) {
	if (g.speaker_mode == GAS_SPEAKER_MODE_STEREO) {
		output_vol_stereo(dir, g.global_panning * s.panning_strength, out);
	} else {
		float tightness = g.global_panning * 2.0f;
		tightness *= s.panning_strength;
		output_vol_surround<NL>(gm, gbase, l, g, dir, tightness, out);
	}
}

static __device__ __forceinline__ float lerpf(float a, float b, float w) { return a + (b - a) * w; }

// reference audio_spatializer_3d.cpp:154-197
// reference audio_spatializer_3d.cpp:154-191: the uniformity > 0 branch (rare: out of line)
struct Vol8 {
	float v[4][2];
};
template <int NL>
static __device__ __noinline__ Vol8 reverb_vol_uniform(unsigned gm, int gbase, int l, const GlobalCfg &g, int attenuation_model, float unit_size,
		float panning_strength, float volume_db, float max_db, float uniformity, float area_send, V3 listener_area_pos, const Vol8 dir8) {
	gas_spatializer s; // the three fields attenuation_db / output_vol read
	s.attenuation_model = attenuation_model;
	s.unit_size = unit_size;
	s.panning_strength = panning_strength;
	const int chan = g.channels;
	float rev[4][2];
	float direct[4][2];
	for (int i = 0; i < 4; i++) {
		rev[i][0] = rev[i][1] = 0.f;
		direct[i][0] = dir8.v[i][0];
		direct[i][1] = dir8.v[i][1];
	}
	float distance = len3(listener_area_pos);
	float attenuation = db_to_linear_f(attenuation_db(s, volume_db, max_db, distance));
	const float center_val[4] = { 0.5f, 0.25f, 0.16666f, 0.125f };
	float cv = center_val[chan - 1];
	if (attenuation < 1.0f) {
		V3 rp = listener_area_pos;
		rp.y = 0.f;
		rp = norm3(rp);
		output_vol<NL>(gm, gbase, l, g, s, rp, rev);
		for (int i = 0; i < chan; i++) {
			rev[i][0] = lerpf(rev[i][0], cv, attenuation);
			rev[i][1] = lerpf(rev[i][1], cv, attenuation);
		}
	} else {
		for (int i = 0; i < chan; i++) {
			rev[i][0] = rev[i][1] = cv;
		}
	}
	for (int i = 0; i < chan; i++) {
		rev[i][0] = lerpf(direct[i][0], rev[i][0] * attenuation, uniformity);
		rev[i][1] = lerpf(direct[i][1], rev[i][1] * attenuation, uniformity);
		rev[i][0] *= area_send;
		rev[i][1] *= area_send;
	}
	Vol8 out;
	for (int i = 0; i < 4; i++) {
		out.v[i][0] = rev[i][0];
		out.v[i][1] = rev[i][1];
	}
	return out;
}

// reference audio_spatializer_3d.cpp:154-197
template <int NL>
static __device__ __forceinline__ void reverb_vol(unsigned gm, int gbase, int l, const GlobalCfg &g, const gas_spatializer &s, const gas_emitter &e,
		const gas_area &a, V3 listener_area_pos, const float direct[4][2], float rev[4][2]) {
	for (int i = 0; i < 4; i++) {
		rev[i][0] = rev[i][1] = 0.f;
	}
	if (a.reverb_uniformity > 0.0f) {
		Vol8 d8;
		for (int i = 0; i < 4; i++) {
			d8.v[i][0] = direct[i][0];
			d8.v[i][1] = direct[i][1];
		}
		const Vol8 r8 = reverb_vol_uniform<NL>(gm, gbase, l, g, s.attenuation_model, s.unit_size, s.panning_strength, e.volume_db, e.max_db,
				a.reverb_uniformity, a.reverb_amount, listener_area_pos, d8);
		for (int i = 0; i < 4; i++) {
			rev[i][0] = r8.v[i][0];
			rev[i][1] = r8.v[i][1];
		}
	} else {
		const float area_send = a.reverb_amount;
		for (int i = 0; i < 4; i++) {
			rev[i][0] = direct[i][0] * area_send;
			rev[i][1] = direct[i][1] * area_send;
		}
	}
}

static __device__ __forceinline__ int resolve_bus(const GlobalCfg &g, int bus) { // audio_stream_player_spatial.cpp:405-413
	return (bus >= 0 && bus < g.num_buses) ? bus : 0;
}

// AudioSpatializerInstance::get_bus_map for all proxy channels at once (audio_spatializer.cpp:274-324):
// Mode B proxies normalise by the mix volume and mask to their own pair; Mode A sends mix_volumes to
// every bus (Q15).
static __device__ void push_bus_map(const gas_params &p, bool mix_channels, BusDetails &d) {
	int n = p.n_bus < GAS_MAX_BUSES_PER_PLAYBACK ? p.n_bus : GAS_MAX_BUSES_PER_PLAYBACK;
	d.n = n;
	for (int k = 0; k < GAS_MAX_BUSES_PER_PLAYBACK; k++) {
		d.bus[k] = k < n ? p.bus[k] : 0;
		for (int c = 0; c < 4; c++) {
			float l = 0.f, r = 0.f;
			if (k < n) {
				if (mix_channels) {
					if (p.mix_volumes[c][0] > 0.0f) {
						l = p.bus_volumes[k][c][0] / p.mix_volumes[c][0];
					}
					if (p.mix_volumes[c][1] > 0.0f) {
						r = p.bus_volumes[k][c][1] / p.mix_volumes[c][1];
					}
				} else {
					l = p.mix_volumes[c][0];
					r = p.mix_volumes[c][1];
				}
			}
			d.vol[k][c][0] = l;
			d.vol[k][c][1] = r;
		}
	}
}

static __device__ __forceinline__ bool inst_mix_channels(const DevTables &t, int q) {
	return (t.inst_mode[q] & 0xff) == MODE_B;
}

// set_spatializer_parameters + bus-map push (audio_spatializer.cpp:258-272)
static __device__ void commit_params(const DevTables &t, int q, const gas_params &p_in) {
	// bus slots beyond n_bus carry nothing (the reference's arrays have exactly n_bus entries): stored as zeros, which is what
	// gain_emitter relies on when it leaves slots it does not use untouched
	gas_params p = p_in;
	for (int b = p.n_bus < 0 ? 0 : p.n_bus; b < GAS_MAX_BUSES_PER_PLAYBACK; b++) {
		p.bus[b] = 0;
		for (int c = 0; c < GAS_MAX_CHANNELS_PER_BUS; c++) {
			p.bus_volumes[b][c][0] = p.bus_volumes[b][c][1] = 0.f;
		}
	}
	t.inst_params[q] = p;
	if (p.update_parameters && t.inst_active[q]) {
		push_bus_map(p, inst_mix_channels(t, q), t.inst_cur[q]);
	}
}

// this lane's element of a [pair][side] table without dynamic indexing (keeps the table in registers)
static __device__ __forceinline__ float pick(const float v[4][2], int c, int x) {
	float r = 0.f;
#pragma unroll
	for (int cc = 0; cc < 4; cc++) {
#pragma unroll
		for (int xx = 0; xx < 2; xx++) {
			r = (cc == c && xx == x) ? v[cc][xx] : r;
		}
	}
	return r;
}

// Per-listener part of calculate_spatialization (audio_spatializer_3d.cpp:335-342, :350-352, :408-409): the transforms
// every emitter would derive from the listener again.
static __device__ void listener_precompute(const gas_listener &L, ListenerPre &p) {
	const Xf lt = load_xf(L);
	Xf inv = lt;
	orthonormalize(inv);
	Xf on = inv;
	affine_invert(inv);
	Xf inv2 = lt;
	affine_invert(inv2);
	store_xf12(p.inv, inv);
	store_xf12(p.inv2, inv2);
	store_xf12(p.on, on);
}

// :378-385 (rare: out of line)
static __device__ __noinline__ float emission_angle_db(float emission_angle, float emission_angle_filter_attenuation_db, V3 basis_z, V3 global_pos,
		V3 listener_origin, float db_att) {
	V3 listenertopos = sub3(global_pos, listener_origin);
	float c = dot3(norm3(listenertopos), norm3(basis_z));
	float ac = c < -1.0f ? (float)3.14159265358979323846 : (c > 1.0f ? 0.0f : (float)acos_d((double)c));
	float angle = ac * (float)(180.0 / 3.14159265358979323846);
	if (angle > emission_angle) {
		db_att -= -emission_angle_filter_attenuation_db;
	}
	return db_att;
}

// :405-427 (rare: out of line)
struct Pitch2 {
	float scale, weight;
};
static __device__ __noinline__ Pitch2 doppler_listener(float pitch_scale, float speed_of_sound, V3 listener_velocity, const ListenerPre *pre, V3 linear_velocity,
		V3 local_pos, float weight, Pitch2 acc) {
	const Xf on = load_xf12(pre->on);
	V3 local_velocity = bxform_inv(on, sub3(linear_velocity, listener_velocity));
	if (!(local_velocity.x == 0.f && local_velocity.y == 0.f && local_velocity.z == 0.f)) {
		float approaching = dot3(norm3(local_pos), norm3(local_velocity));
		float velocity = len3(local_velocity);
		float dps = pitch_scale * speed_of_sound / (speed_of_sound + velocity * approaching);
		dps = (double)dps < 0.125 ? 0.125f : ((double)dps > 8.0 ? 8.0f : dps);
		acc.scale += weight * (float)log2_d((double)dps);
		acc.weight += weight;
	}
	return acc;
}

// NL lanes per emitter (8, 4 or 2).  The scalar chain is evaluated redundantly by the NL lanes; the SPCAP speaker gains
// and the stores are dealt across them (lane l owns elements l, l + NL, ... of the [pair][side] tables).  Fewer lanes
// = fewer, longer threads: slower on an empty GPU (the chain is latency-bound), but a smaller footprint beside the
// mix kernels, which is what a step pays for (see launch_gain for the measured shapes).
// Returns the instance the emitter belongs to (-1: record skipped).
template <int NL>
static __device__ __forceinline__ int gain_emitter(const DevTables &t, const GlobalCfg &g, int i, int l, int gbase, unsigned gm,
		const gas_emitter *__restrict__ emitters, int n_listeners, const gas_listener *__restrict__ listeners, const ListenerPre *__restrict__ pre,
		const gas_area *__restrict__ areas, int n_areas, gas_params *__restrict__ out, int32_t *__restrict__ inst_seq = nullptr, int seq_val = 0,
		unsigned long long *dbg = nullptr) {
	constexpr int S = 8 / NL;
#define GAS_GAIN_STAMP(k_)                                            \
	if (dbg) {                                                        \
		dbg[k_] = gas_globaltimer();                                  \
	}
	gas_emitter e = emitters[i];
	// Device-resident emitter records are not seen by the host: a record that points outside the tables is skipped (its
	// instance keeps its parameters), an area index outside the resident areas counts as "no area" — like the prologue
	// does with voice records.  Uniform over the emitter's lanes, so the shuffles below stay converged.
	if (e.instance < 0 || e.instance >= g.max_instances || e.spatializer < 0 || e.spatializer >= g.max_spatializers) {
		return -1;
	}
	if (e.area >= n_areas) {
		e.area = -1;
	}
	// everything behind the emitter record is requested at once: the fields of its spatializer, its area and
	// the state of its instance
	const gas_spatializer *sp = &t.spat[e.spatializer];
	gas_spatializer s;
	s.attenuation_model = sp->attenuation_model;
	s.unit_size = sp->unit_size;
	s.max_distance = sp->max_distance;
	s.panning_strength = sp->panning_strength;
	s.emission_angle_enabled = sp->emission_angle_enabled;
	s.emission_angle = sp->emission_angle;
	s.emission_angle_filter_attenuation_db = sp->emission_angle_filter_attenuation_db;
	s.attenuation_filter_cutoff_hz = sp->attenuation_filter_cutoff_hz;
	s.attenuation_filter_db = sp->attenuation_filter_db;
	s.doppler_tracking = sp->doppler_tracking;
	s.doppler_speed_of_sound = sp->doppler_speed_of_sound;
	const int q = e.instance;
	// bus slots the instance's records hold at the moment: slots this computation leaves unused are cleared only if they were in use
	// (parameters handed in through gas_params_set may carry up to six; calculate_spatialization produces at most two)
	const int old_n_bus = t.inst_params[q].n_bus;
	const int old_n_cur = t.inst_cur[q].n;
	const bool was_further = t.inst_was_further[q] != 0;
	const int q_active = t.inst_active[q];
	const bool mix_channels = inst_mix_channels(t, q);
	const bool has_area = e.area >= 0;
	const gas_area *ap = has_area ? &areas[e.area] : nullptr;
	struct {
		float pitch_scale, linear_attenuation, attenuation_filter_cutoff_hz;
		int update_parameters;
	} prm;
	prm.pitch_scale = 1.0f;
	prm.linear_attenuation = 0.0f;
	prm.attenuation_filter_cutoff_hz = 5000.0f;
	prm.update_parameters = 0;

	GAS_GAIN_STAMP(23) // emitter record here
	const V3 global_pos{ e.origin[0], e.origin[1], e.origin[2] };
	V3 linear_velocity{ 0.f, 0.f, 0.f };
	const bool doppler = s.doppler_tracking != GAS_DOPPLER_TRACKING_DISABLED;
	if (doppler) { // :297-299
		linear_velocity = V3{ e.velocity[0], e.velocity[1], e.velocity[2] };
	}
	float log_pitch_scale = 0.f, log_pitch_weight = 0.f;
	float output_volume[4][2], reverb_volume[4][2], tmp_volume[4][2], tmp_reverb[4][2];
	for (int c = 0; c < 4; c++) {
		output_volume[c][0] = output_volume[c][1] = 0.f;
		reverb_volume[c][0] = reverb_volume[c][1] = 0.f;
	}
	bool in_range_any = false;
	const bool area_reverb_uniform = has_area && ap->use_reverb && ap->reverb_uniformity > 0.0f;

	for (int li = 0; li < n_listeners; li++) { // :323
		const gas_listener L = listeners[li];
		// the listener's transforms (orthonormalised inverse, plain affine inverse, orthonormalised basis) are the same for
		// every emitter: computed once per listener upload by k_listener_pre with the very code the reference runs per emitter
		const Xf inv = load_xf12(pre[li].inv);
		const V3 local_pos = xform(inv, global_pos); // :342
		const float dist = len3(local_pos);          // :344
		V3 listener_area_pos{ 0.f, 0.f, 0.f };
		if (area_reverb_uniform) { // :350-353 (plain affine inverse, not orthonormalised)
			const Xf inv2 = load_xf12(pre[li].inv2);
			listener_area_pos = xform(inv2, V3{ ap->closest_point[li][0], ap->closest_point[li][1], ap->closest_point[li][2] });
		}
		GAS_GAIN_STAMP(24) // tables + listener here
		float multiplier = db_to_linear_f(attenuation_db(s, e.volume_db, e.max_db, dist)); // :359
		if (s.max_distance > 0.f) { // :361-373
			float total_max = s.max_distance;
			if (area_reverb_uniform) {
				float lap = len3(listener_area_pos);
				total_max = total_max > lap ? total_max : lap;
			}
			if (dist > total_max || total_max > s.max_distance) {
				continue;
			}
			double m = 1.0 - (double)(dist / s.max_distance);
			m = 0.0 > m ? 0.0 : m;
			multiplier = (float)((double)multiplier * m);
		}
		in_range_any = true;

		double mm = 1.0 < (double)multiplier ? 1.0 : (double)multiplier;
		float db_att = (float)((1.0 - mm) * (double)s.attenuation_filter_db); // :376
		if (s.emission_angle_enabled) { // :378-385
			db_att = emission_angle_db(s.emission_angle, s.emission_angle_filter_attenuation_db, V3{ e.basis_z[0], e.basis_z[1], e.basis_z[2] }, global_pos,
					V3{ L.origin[0], L.origin[1], L.origin[2] }, db_att);
		}
		prm.linear_attenuation = db_to_linear_f(db_att); // :387, last listener wins (Q6)
		prm.attenuation_filter_cutoff_hz = s.attenuation_filter_cutoff_hz;

		GAS_GAIN_STAMP(25) // attenuation + filter gain done
		for (int c = 0; c < 4; c++) {
			tmp_volume[c][0] = tmp_volume[c][1] = 0.f;
		}
		output_vol<NL>(gm, gbase, l, g, s, local_pos, tmp_volume); // :391 — un-normalised direction (Q1)
		for (int c = 0; c < 4; c++) {            // :393-396
			tmp_volume[c][0] = multiplier * tmp_volume[c][0];
			tmp_volume[c][1] = multiplier * tmp_volume[c][1];
			output_volume[c][0] = output_volume[c][0] > tmp_volume[c][0] ? output_volume[c][0] : tmp_volume[c][0];
			output_volume[c][1] = output_volume[c][1] > tmp_volume[c][1] ? output_volume[c][1] : tmp_volume[c][1];
		}
		if (has_area && ap->use_reverb) { // :399-402
			reverb_vol<NL>(gm, gbase, l, g, s, e, *ap, listener_area_pos, tmp_volume, tmp_reverb);
			for (int c = 0; c < 4; c++) {
				reverb_volume[c][0] = reverb_volume[c][0] > tmp_reverb[c][0] ? reverb_volume[c][0] : tmp_reverb[c][0];
				reverb_volume[c][1] = reverb_volume[c][1] > tmp_reverb[c][1] ? reverb_volume[c][1] : tmp_reverb[c][1];
			}
		}
		if (doppler) { // :405-427
			float weight = 0.f;
			for (int c = 0; c < 4; c++) {
				weight = weight > tmp_volume[c][0] ? weight : tmp_volume[c][0];
				weight = weight > tmp_volume[c][1] ? weight : tmp_volume[c][1];
			}
			const Pitch2 pa = doppler_listener(e.pitch_scale, s.doppler_speed_of_sound, V3{ L.velocity[0], L.velocity[1], L.velocity[2] }, &pre[li],
					linear_velocity, local_pos, weight, Pitch2{ log_pitch_scale, log_pitch_weight });
			log_pitch_scale = pa.scale;
			log_pitch_weight = pa.weight;
		}
	}
	GAS_GAIN_STAMP(26) // listeners done
	if (log_pitch_weight > 0.f) { // :430-434
		prm.pitch_scale = (float)pow_d(2.0, (double)(log_pitch_scale / log_pitch_weight));
	} else {
		prm.pitch_scale = e.pitch_scale;
	}
	// :437-461 — bus entries in Dictionary insertion order; a second add to the same bus overwrites its volumes
	int n_bus = 0, bus0 = 0, bus1 = 0;
	bool slot0_is_reverb = false;
	if (in_range_any) {
		if (has_area) {
			bus0 = resolve_bus(g, ap->override_bus ? ap->bus : e.bus);
			n_bus = 1;
			if (ap->use_reverb) {
				const int rb = resolve_bus(g, ap->reverb_bus);
				if (rb == bus0) {
					slot0_is_reverb = true;
				} else {
					bus1 = rb;
					n_bus = 2;
				}
			}
		} else {
			bus0 = resolve_bus(g, e.bus);
			n_bus = 1;
		}
	}
	float mv[S], bv0[S], bv1[S]; // this lane's elements (pair, side) = ((l + k L) >> 1, (l + k L) & 1)
#pragma unroll
	for (int k = 0; k < S; k++) {
		const int el = l + k * NL;
		mv[k] = pick(output_volume, el >> 1, el & 1); // :463
		const float rv = pick(reverb_volume, el >> 1, el & 1);
		bv0[k] = n_bus > 0 ? (slot0_is_reverb ? rv : mv[k]) : 0.f;
		bv1[k] = n_bus > 1 ? rv : 0.f;
	}
	const bool skip = !in_range_any && was_further; // :466-467
	__syncwarp(gm); // every lane has read was_further before lane 0 rewrites it
	if (!skip) {
		prm.update_parameters = 1;
	}
	// set_spatializer_parameters + bus-map push (audio_spatializer.cpp:258-272), the (pair, side) elements dealt over the lanes
#pragma unroll
	for (int pass = 0; pass < 2; pass++) {
		gas_params *P = pass == 0 ? &t.inst_params[q] : (out ? &out[i] : nullptr);
		if (!P) {
			continue;
		}
#pragma unroll
		for (int k = 0; k < S; k++) {
			const int el = l + k * NL;
			P->mix_volumes[el >> 1][el & 1] = mv[k];
#pragma unroll
			for (int b = 0; b < GAS_MAX_BUSES_PER_PLAYBACK; b++) {
				if (b < 2 || pass == 1 || b < old_n_bus) {
					P->bus_volumes[b][el >> 1][el & 1] = b == 0 ? bv0[k] : (b == 1 ? bv1[k] : 0.f);
				}
			}
		}
		if (l == 0) {
			P->pitch_scale = prm.pitch_scale;
			P->linear_attenuation = prm.linear_attenuation;
			P->attenuation_filter_cutoff_hz = prm.attenuation_filter_cutoff_hz;
			P->update_parameters = prm.update_parameters;
			P->n_bus = n_bus;
#pragma unroll
			for (int b = 0; b < GAS_MAX_BUSES_PER_PLAYBACK; b++) {
				if (b < 2 || pass == 1 || b < old_n_bus) {
					P->bus[b] = b == 0 ? bus0 : (b == 1 ? bus1 : 0);
				}
			}
		}
	}
	if (l == 0) {
		t.inst_was_further[q] = in_range_any ? 0 : 1;
	}
	if (prm.update_parameters && q_active) { // get_bus_map of all proxy channels (audio_spatializer.cpp:274-324)
		BusDetails *d = &t.inst_cur[q];
#pragma unroll
		for (int k = 0; k < S; k++) {
			const int el = l + k * NL;
#pragma unroll
			for (int b = 0; b < GAS_MAX_BUSES_PER_PLAYBACK; b++) {
				float w = 0.f;
				if (b < n_bus) {
					const float bv = b == 0 ? bv0[k] : bv1[k];
					w = mix_channels ? (mv[k] > 0.0f ? bv / mv[k] : 0.f) : mv[k];
				}
				if (b < 2 || b < old_n_cur) {
					d->vol[b][el >> 1][el & 1] = w;
				}
			}
		}
		if (l == 0) {
			d->n = n_bus;
#pragma unroll
			for (int b = 0; b < GAS_MAX_BUSES_PER_PLAYBACK; b++) {
				if (b < 2 || b < old_n_cur) {
					d->bus[b] = b == 0 ? (n_bus > 0 ? bus0 : 0) : (b == 1 && n_bus > 1 ? bus1 : 0);
				}
			}
		}
	}
	GAS_GAIN_STAMP(27) // stores issued
	if (inst_seq) {
		// in-kernel gains (step kernel): tell the planner of this block that the instance's parameters are in place
		__syncwarp(gm);
		if (l == 0) {
			gas_st_release_gpu_s32(inst_seq + q, seq_val);
		}
	}
	return q;
}


} // namespace gasgain
