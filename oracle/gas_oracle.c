/*
 * gas_oracle.c — CPU oracle (scalar C restatement of the reference hot path).
 * TEST INFRASTRUCTURE ONLY; pinned against the reference's own code by oracle/_ref — see gas_oracle.h.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp (no -ffast-math): Godot's release optimisation level
 * without FMA contraction, so every float operation below rounds where the reference's does on a
 * baseline x86-64 build.
 */
#include "gas_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define CMP_EPSILON 0.00001 /* upstream core/math/math_defs.h (double literal) */

/* ---------------------------------------------------------------------------------------------
 * upstream Math::* (core/math/math_funcs.h), SURVEY Appendix A
 * ------------------------------------------------------------------------------------------- */
float orc_db_to_linear_f(float db) {
	return expf(db * (float)0.11512925464970228420089957273422);
}
double orc_db_to_linear_d(double db) {
	return exp(db * 0.11512925464970228420089957273422);
}
float orc_linear_to_db_f(float lin) {
	return logf(lin) * (float)8.6858896380650365530225783783321;
}
double orc_linear_to_db_d(double lin) {
	return log(lin) * 8.6858896380650365530225783783321;
}

/* upstream Vector3 (real_t = float) */
typedef struct v3 {
	float x, y, z;
} v3;
static inline float v3_dot(v3 a, v3 b) {
	return a.x * b.x + a.y * b.y + a.z * b.z;
}
static inline float v3_length(v3 a) {
	return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z);
}
static inline v3 v3_normalized(v3 a) {
	float lengthsq = a.x * a.x + a.y * a.y + a.z * a.z;
	if (lengthsq == 0) {
		v3 z = { 0, 0, 0 };
		return z;
	}
	float length = sqrtf(lengthsq);
	v3 r = { a.x / length, a.y / length, a.z / length };
	return r;
}
static inline v3 v3_sub(v3 a, v3 b) {
	v3 r = { a.x - b.x, a.y - b.y, a.z - b.z };
	return r;
}
static inline v3 v3_scale(v3 a, float s) {
	v3 r = { a.x * s, a.y * s, a.z * s };
	return r;
}

/* upstream Basis (rows[3]) / Transform3D */
typedef struct xf3 {
	float m[3][3];
	v3 o;
} xf3;
static v3 xf_col(const xf3 *t, int c) {
	v3 r = { t->m[0][c], t->m[1][c], t->m[2][c] };
	return r;
}
static void xf_set_col(xf3 *t, int c, v3 v) {
	t->m[0][c] = v.x;
	t->m[1][c] = v.y;
	t->m[2][c] = v.z;
}
/* Basis::orthonormalize (Gram-Schmidt) */
static void xf_orthonormalize(xf3 *t) {
	v3 x = xf_col(t, 0), y = xf_col(t, 1), z = xf_col(t, 2);
	x = v3_normalized(x);
	y = v3_sub(y, v3_scale(x, v3_dot(x, y)));
	y = v3_normalized(y);
	z = v3_sub(v3_sub(z, v3_scale(x, v3_dot(x, z))), v3_scale(y, v3_dot(y, z)));
	z = v3_normalized(z);
	xf_set_col(t, 0, x);
	xf_set_col(t, 1, y);
	xf_set_col(t, 2, z);
}
static inline v3 basis_xform(const xf3 *t, v3 v) {
	v3 r = { t->m[0][0] * v.x + t->m[0][1] * v.y + t->m[0][2] * v.z,
		t->m[1][0] * v.x + t->m[1][1] * v.y + t->m[1][2] * v.z,
		t->m[2][0] * v.x + t->m[2][1] * v.y + t->m[2][2] * v.z };
	return r;
}
/* Basis::xform_inv: transposed multiply */
static inline v3 basis_xform_inv(const xf3 *t, v3 v) {
	v3 r = { t->m[0][0] * v.x + t->m[1][0] * v.y + t->m[2][0] * v.z,
		t->m[0][1] * v.x + t->m[1][1] * v.y + t->m[2][1] * v.z,
		t->m[0][2] * v.x + t->m[1][2] * v.y + t->m[2][2] * v.z };
	return r;
}
/* Transform3D::affine_invert: Basis::invert (cofactors) then origin = basis.xform(-origin) */
static void xf_affine_invert(xf3 *t) {
#define COFAC(r1, c1, r2, c2) (t->m[r1][c1] * t->m[r2][c2] - t->m[r1][c2] * t->m[r2][c1])
	float co[3] = { COFAC(1, 1, 2, 2), COFAC(1, 2, 2, 0), COFAC(1, 0, 2, 1) };
	float det = t->m[0][0] * co[0] + t->m[0][1] * co[1] + t->m[0][2] * co[2];
	float s = 1.0f / det;
	float n[3][3] = {
		{ co[0] * s, COFAC(0, 2, 2, 1) * s, COFAC(0, 1, 1, 2) * s },
		{ co[1] * s, COFAC(0, 0, 2, 2) * s, COFAC(0, 2, 1, 0) * s },
		{ co[2] * s, COFAC(0, 1, 2, 0) * s, COFAC(0, 0, 1, 1) * s },
	};
#undef COFAC
	memcpy(t->m, n, sizeof(n));
	v3 neg = { -t->o.x, -t->o.y, -t->o.z };
	t->o = basis_xform(t, neg);
}
static inline v3 xf_xform(const xf3 *t, v3 v) {
	v3 r = basis_xform(t, v);
	r.x += t->o.x;
	r.y += t->o.y;
	r.z += t->o.z;
	return r;
}
static xf3 xf_from_listener(const gas_listener *l) {
	xf3 t;
	memcpy(t.m, l->basis, sizeof(t.m));
	t.o.x = l->origin[0];
	t.o.y = l->origin[1];
	t.o.z = l->origin[2];
	return t;
}

/* ---------------------------------------------------------------------------------------------
 * SPCAP — audio_spatializer_3d.cpp:47-55 (speaker directions), :903-916, :926-938
 * ------------------------------------------------------------------------------------------- */
static void spcap_dirs(v3 d[7]) {
	static const float raw[7][3] = {
		{ -1, 0, -1 }, { 1, 0, -1 }, { 0, 0, -1 }, { -1, 0, 1 }, { 1, 0, 1 }, { -1, 0, 0 }, { 1, 0, 0 }
	};
	for (int i = 0; i < 7; i++) {
		v3 v = { raw[i][0], raw[i][1], raw[i][2] };
		d[i] = v3_normalized(v);
	}
}

void orc_spcap_effective_speakers(int speaker_count, float eff[7]) {
	v3 d[7];
	spcap_dirs(d);
	for (int i = 0; i < 7; i++) {
		eff[i] = 0.0f;
	}
	for (int i = 0; i < speaker_count; i++) { /* :911-915 — float += double, narrowed every iteration */
		for (int j = 0; j < speaker_count; j++) {
			eff[i] = (float)((double)eff[i] + 0.5 * (1.0 + (double)v3_dot(d[i], d[j])));
		}
	}
}

void orc_spcap_calculate(int speaker_count, const float dir[3], float tightness, float volumes[7]) {
	v3 d[7];
	float eff[7];
	float sq[7];
	spcap_dirs(d);
	orc_spcap_effective_speakers(speaker_count, eff);
	v3 src = { dir[0], dir[1], dir[2] };
	float sum_squared_gains = 0.0f;
	for (int i = 0; i < speaker_count; i++) { /* :929-933 */
		float initial_gain = (float)(0.5 * pow(1.0 + (double)v3_dot(d[i], src), (double)tightness) / (double)eff[i]);
		sq[i] = initial_gain * initial_gain;
		sum_squared_gains += sq[i];
	}
	for (int i = 0; i < speaker_count; i++) { /* :935-937 */
		volumes[i] = sqrtf(sq[i] / sum_squared_gains);
	}
}

/* audio_spatializer_3d.cpp:57-98 */
void orc_calc_output_vol_surround(int speaker_mode, const float dir[3], float tightness, float out[4][2]) {
	int speaker_count = 0;
	switch (speaker_mode) {
		case GAS_SPEAKER_MODE_STEREO:
			speaker_count = 2;
			break;
		case GAS_SPEAKER_SURROUND_31:
			speaker_count = 3;
			break;
		case GAS_SPEAKER_SURROUND_51:
			speaker_count = 5;
			break;
		case GAS_SPEAKER_SURROUND_71:
			speaker_count = 7;
			break;
	}
	float volumes[7] = { 0 };
	orc_spcap_calculate(speaker_count, dir, tightness, volumes);
	switch (speaker_mode) {
		case GAS_SPEAKER_SURROUND_71:
			out[3][0] = volumes[5];
			out[3][1] = volumes[6];
			/* fallthrough */
		case GAS_SPEAKER_SURROUND_51:
			out[2][0] = volumes[3];
			out[2][1] = volumes[4];
			/* fallthrough */
		case GAS_SPEAKER_SURROUND_31:
			out[1][0] = volumes[2];
			out[1][1] = 1.0f; /* LFE - always full power */
			/* fallthrough */
		case GAS_SPEAKER_MODE_STEREO:
			out[0][0] = volumes[0];
			out[0][1] = volumes[1];
			break;
	}
}

/* audio_spatializer_3d.cpp:103-110 */
void orc_calc_output_vol_stereo(const float dir[3], float pan_strength, float out[4][2]) {
	double flatrad = sqrt((double)(dir[0] * dir[0] + dir[2] * dir[2]));
	double g = (1.0 - pan_strength) * (1.0 - pan_strength);
	g = g < 0.0 ? 0.0 : (g > 1.0 ? 1.0 : g);
	double f = (1.0 - g) / (1.0 + g);
	double cosx = dir[0] / (flatrad == 0.0 ? 1.0 : flatrad);
	cosx = cosx < -1.0 ? -1.0 : (cosx > 1.0 ? 1.0 : cosx);
	double fcosx = cosx * f;
	out[0][0] = (float)sqrt((-fcosx + 1.0) / 2.0);
	out[0][1] = (float)sqrt((fcosx + 1.0) / 2.0);
}

/* audio_spatializer_3d.cpp:112-121 */
void orc_calc_output_vol(int speaker_mode, float global_panning, float panning_strength, const float dir[3], float out[4][2]) {
	if (speaker_mode == GAS_SPEAKER_MODE_STEREO) {
		orc_calc_output_vol_stereo(dir, global_panning * panning_strength, out);
	} else {
		float tightness = global_panning * 2.0f;
		tightness *= panning_strength;
		orc_calc_output_vol_surround(speaker_mode, dir, tightness, out);
	}
}

/* audio_spatializer_3d.cpp:123-151 */
float orc_get_attenuation_db(const gas_spatializer *s, float volume_db, float max_db, float p_distance) {
	float att = 0;
	switch (s->attenuation_model) {
		case GAS_ATTENUATION_INVERSE_DISTANCE: {
			att = (float)orc_linear_to_db_d(1.0 / ((p_distance / s->unit_size) + CMP_EPSILON));
		} break;
		case GAS_ATTENUATION_INVERSE_SQUARE_DISTANCE: {
			float d = (p_distance / s->unit_size);
			d *= d;
			att = (float)orc_linear_to_db_d(1.0 / (d + CMP_EPSILON));
		} break;
		case GAS_ATTENUATION_LOGARITHMIC: {
			att = (float)(-20 * log(p_distance / s->unit_size + CMP_EPSILON));
		} break;
		case GAS_ATTENUATION_DISABLED:
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

static inline float lerpf(float from, float to, float w) { /* upstream Math::lerp */
	return from + (to - from) * w;
}

/* audio_spatializer_3d.cpp:154-197 */
static void calc_reverb_vol(const gas_config *cfg, const gas_spatializer *s, const gas_emitter *e, const gas_area *area,
		v3 listener_area_pos, float direct[4][2], float reverb[4][2]) {
	memset(reverb, 0, sizeof(float) * 8);
	float uniformity = area->reverb_uniformity;
	float area_send = area->reverb_amount;
	int chan_count = cfg->speaker_mode + 1;
	if (uniformity > 0.0) {
		float distance = v3_length(listener_area_pos);
		float attenuation = orc_db_to_linear_f(orc_get_attenuation_db(s, e->volume_db, e->max_db, distance));
		const float center_val[4] = { 0.5f, 0.25f, 0.16666f, 0.125f };
		float cv = center_val[chan_count - 1];
		if (attenuation < 1.0) {
			v3 rev_pos = listener_area_pos;
			rev_pos.y = 0;
			rev_pos = v3_normalized(rev_pos);
			float rp[3] = { rev_pos.x, rev_pos.y, rev_pos.z };
			orc_calc_output_vol(cfg->speaker_mode, cfg->global_panning_strength, s->panning_strength, rp, reverb);
			for (int i = 0; i < chan_count; i++) {
				reverb[i][0] = lerpf(reverb[i][0], cv, attenuation);
				reverb[i][1] = lerpf(reverb[i][1], cv, attenuation);
			}
		} else {
			for (int i = 0; i < chan_count; i++) {
				reverb[i][0] = cv;
				reverb[i][1] = cv;
			}
		}
		for (int i = 0; i < chan_count; i++) {
			reverb[i][0] = lerpf(direct[i][0], reverb[i][0] * attenuation, uniformity);
			reverb[i][1] = lerpf(direct[i][1], reverb[i][1] * attenuation, uniformity);
			reverb[i][0] *= area_send;
			reverb[i][1] *= area_send;
		}
	} else {
		for (int i = 0; i < 4; i++) {
			reverb[i][0] = direct[i][0] * area_send;
			reverb[i][1] = direct[i][1] * area_send;
		}
	}
}

/* SpatializerParameters::add_bus_volume: Dictionary assignment keeps the first insertion position */
static void params_add_bus_volume(gas_params *p, int bus, float vol[4][2]) {
	for (int i = 0; i < p->n_bus; i++) {
		if (p->bus[i] == bus) {
			memcpy(p->bus_volumes[i], vol, sizeof(float) * 8);
			return;
		}
	}
	if (p->n_bus >= GAS_MAX_BUSES_PER_PLAYBACK) {
		return;
	}
	p->bus[p->n_bus] = bus;
	memcpy(p->bus_volumes[p->n_bus], vol, sizeof(float) * 8);
	p->n_bus++;
}

static int resolve_bus(const gas_config *cfg, int bus) { /* audio_stream_player_spatial.cpp:405-413 */
	return (bus >= 0 && bus < cfg->num_buses) ? bus : 0;
}

/* AudioSpatializerInstance3D::calculate_spatialization, audio_spatializer_3d.cpp:277-489, minus the
 * scene / physics look-ups whose results arrive in gas_emitter / gas_listener / gas_area. */
void orc_calculate_spatialization(const gas_config *cfg, const gas_spatializer *s, const gas_emitter *e,
		int n_listeners, const gas_listener *listeners, const gas_area *area, int *was_further, gas_params *out) {
	gas_params prm;
	memset(&prm, 0, sizeof(prm));
	prm.pitch_scale = 1.0f;                      /* spatializer_parameters.h:48 */
	prm.linear_attenuation = 0.0f;               /* audio_spatializer_3d.h:67 */
	prm.attenuation_filter_cutoff_hz = 5000.0f;  /* audio_spatializer_3d.h:68 */

	v3 global_pos = { e->origin[0], e->origin[1], e->origin[2] };
	v3 linear_velocity = { 0, 0, 0 };
	if (s->doppler_tracking != GAS_DOPPLER_TRACKING_DISABLED) { /* :297-299 */
		linear_velocity.x = e->velocity[0];
		linear_velocity.y = e->velocity[1];
		linear_velocity.z = e->velocity[2];
	}
	float log_pitch_scale = 0.0f;
	float log_pitch_weight = 0.0f;
	float output_volume[4][2] = { { 0 } };
	float reverb_volume[4][2] = { { 0 } };
	float tmp_volume[4][2];
	float tmp_reverb[4][2];
	int has_any_listener_in_range = 0;
	const int area_reverb_uniform = area && area->use_reverb && area->reverb_uniformity > 0;

	for (int li = 0; li < n_listeners; li++) { /* :323 */
		const gas_listener *L = &listeners[li];
		xf3 lt = xf_from_listener(L);
		xf3 inv = lt;
		xf_orthonormalize(&inv);
		xf_affine_invert(&inv);
		const v3 local_pos = xf_xform(&inv, global_pos); /* :342 */
		const float dist = v3_length(local_pos);        /* :344 */

		v3 listener_area_pos = { 0, 0, 0 };
		if (area_reverb_uniform) { /* :350-353 — NOT orthonormalized */
			v3 area_sound_pos = { area->closest_point[li][0], area->closest_point[li][1], area->closest_point[li][2] };
			xf3 inv2 = lt;
			xf_affine_invert(&inv2);
			listener_area_pos = xf_xform(&inv2, area_sound_pos);
		}

		float multiplier = orc_db_to_linear_f(orc_get_attenuation_db(s, e->volume_db, e->max_db, dist)); /* :359 */
		if (s->max_distance > 0) { /* :361-373 */
			float total_max = s->max_distance;
			if (area_reverb_uniform) {
				float lap = v3_length(listener_area_pos);
				total_max = total_max > lap ? total_max : lap;
			}
			if (dist > total_max || total_max > s->max_distance) {
				continue;
			}
			double m = 1.0 - (dist / s->max_distance);
			m = 0 > m ? 0 : m;
			multiplier = (float)((double)multiplier * m);
		}
		has_any_listener_in_range = 1;

		double mm = 1.0 < (double)multiplier ? 1.0 : (double)multiplier;
		float db_att = (float)((1.0 - mm) * (double)s->attenuation_filter_db); /* :376 */

		if (s->emission_angle_enabled) { /* :378-385 */
			v3 lo = { L->origin[0], L->origin[1], L->origin[2] };
			v3 listenertopos = v3_sub(global_pos, lo);
			v3 bz = { e->basis_z[0], e->basis_z[1], e->basis_z[2] };
			float c = v3_dot(v3_normalized(listenertopos), v3_normalized(bz));
			float ac = c < -1.0f ? (float)3.14159265358979323846 : (c > 1.0f ? 0.0f : acosf(c)); /* upstream Math::acos clamps */
			float angle = ac * (float)(180.0 / 3.14159265358979323846);
			if (angle > s->emission_angle) {
				db_att -= -s->emission_angle_filter_attenuation_db;
			}
		}
		prm.linear_attenuation = orc_db_to_linear_f(db_att);                  /* :387 — last listener wins */
		prm.attenuation_filter_cutoff_hz = s->attenuation_filter_cutoff_hz; /* :388 */

		memset(tmp_volume, 0, sizeof(tmp_volume)); /* :390 */
		float lp[3] = { local_pos.x, local_pos.y, local_pos.z };
		orc_calc_output_vol(cfg->speaker_mode, cfg->global_panning_strength, s->panning_strength, lp, tmp_volume); /* :391 — NOT normalised (Q1) */
		for (int k = 0; k < 4; k++) { /* :393-396 */
			tmp_volume[k][0] = multiplier * tmp_volume[k][0];
			tmp_volume[k][1] = multiplier * tmp_volume[k][1];
			output_volume[k][0] = output_volume[k][0] > tmp_volume[k][0] ? output_volume[k][0] : tmp_volume[k][0];
			output_volume[k][1] = output_volume[k][1] > tmp_volume[k][1] ? output_volume[k][1] : tmp_volume[k][1];
		}
		if (area && area->use_reverb) { /* :399-402 */
			calc_reverb_vol(cfg, s, e, area, listener_area_pos, tmp_volume, tmp_reverb);
			for (int k = 0; k < 4; k++) {
				reverb_volume[k][0] = reverb_volume[k][0] > tmp_reverb[k][0] ? reverb_volume[k][0] : tmp_reverb[k][0];
				reverb_volume[k][1] = reverb_volume[k][1] > tmp_reverb[k][1] ? reverb_volume[k][1] : tmp_reverb[k][1];
			}
		}
		if (s->doppler_tracking != GAS_DOPPLER_TRACKING_DISABLED) { /* :405-427 */
			v3 lv = { L->velocity[0], L->velocity[1], L->velocity[2] };
			xf3 on = lt;
			xf_orthonormalize(&on);
			v3 local_velocity = basis_xform_inv(&on, v3_sub(linear_velocity, lv));
			if (!(local_velocity.x == 0 && local_velocity.y == 0 && local_velocity.z == 0)) {
				float approaching = v3_dot(v3_normalized(local_pos), v3_normalized(local_velocity));
				float velocity = v3_length(local_velocity);
				float dps = e->pitch_scale * s->doppler_speed_of_sound / (s->doppler_speed_of_sound + velocity * approaching);
				dps = (double)dps < (1 / 8.0) ? (float)(1 / 8.0) : ((double)dps > 8.0 ? 8.0f : dps);
				float weight = 0.0f; /* _get_max_volume, :268-275 */
				for (int k = 0; k < 4; k++) {
					weight = weight > tmp_volume[k][0] ? weight : tmp_volume[k][0];
					weight = weight > tmp_volume[k][1] ? weight : tmp_volume[k][1];
				}
				log_pitch_scale += weight * log2f(dps);
				log_pitch_weight += weight;
			}
		}
	}

	if (log_pitch_weight > 0) { /* :430-434 */
		prm.pitch_scale = powf(2.0f, log_pitch_scale / log_pitch_weight);
	} else {
		prm.pitch_scale = e->pitch_scale;
	}

	if (has_any_listener_in_range) { /* :437-461 */
		if (area) {
			if (area->override_bus) {
				params_add_bus_volume(&prm, resolve_bus(cfg, area->bus), output_volume);
			} else {
				params_add_bus_volume(&prm, resolve_bus(cfg, e->bus), output_volume);
			}
			if (area->use_reverb) {
				params_add_bus_volume(&prm, resolve_bus(cfg, area->reverb_bus), reverb_volume);
			}
		} else {
			params_add_bus_volume(&prm, resolve_bus(cfg, e->bus), output_volume);
		}
	}
	memcpy(prm.mix_volumes, output_volume, sizeof(output_volume)); /* :463 */

	const int skip_setting_volumes = !has_any_listener_in_range && *was_further; /* :466 */
	*was_further = !has_any_listener_in_range;                                  /* :467 */
	if (!skip_setting_volumes) {
		prm.update_parameters = 1; /* :471 */
	}
	*out = prm;
}

/* AudioSpatializerInstance::get_bus_map, audio_spatializer.cpp:274-324 */
int orc_get_bus_map(const gas_params *p, int mix_channels, int channel, int out_bus[6], float out_vol[6][4][2]) {
	int idx = 0;
	if (channel < 0 || channel >= GAS_MAX_CHANNELS_PER_BUS) {
		return 0;
	}
	for (int k = 0; k < p->n_bus; k++) {
		if (idx >= GAS_MAX_BUSES_PER_PLAYBACK) {
			break;
		}
		out_bus[idx] = p->bus[k];
		for (int c = 0; c < GAS_MAX_CHANNELS_PER_BUS; c++) {
			if (mix_channels) {
				float left = 0.0f, right = 0.0f;
				if (c == channel) {
					if (p->mix_volumes[c][0] > 0.0) {
						left = p->bus_volumes[k][c][0] / p->mix_volumes[c][0];
					}
					if (p->mix_volumes[c][1] > 0.0) {
						right = p->bus_volumes[k][c][1] / p->mix_volumes[c][1];
					}
				}
				out_vol[idx][c][0] = left;
				out_vol[idx][c][1] = right;
			} else {
				out_vol[idx][c][0] = p->mix_volumes[c][0]; /* Q15: mix volumes to every bus */
				out_vol[idx][c][1] = p->mix_volumes[c][1];
			}
		}
		idx++;
	}
	return idx;
}

/* upstream AudioFilterSW::prepare_coefficients: all arithmetic in double, each
 * coefficient narrowed to float when stored and again after the division by a0. */
void orc_filter_prepare_coefficients(int mode, float cutoff, float resonance, float gain, int stages, float sampling_rate, float out[5]) {
	int sr_limit = (int)((sampling_rate / 2) + 512);
	double final_cutoff = (cutoff > sr_limit) ? sr_limit : cutoff;
	if (final_cutoff < 1) {
		final_cutoff = 1;
	}
	const double TAU = 6.2831853071795864769252867666;
	double omega = TAU * final_cutoff / sampling_rate;
	double sin_v = sin(omega);
	double cos_v = cos(omega);
	double Q = resonance;
	if (Q <= 0.0) {
		Q = 0.0001;
	}
	if (mode == GAS_FILTER_BANDPASS) {
		Q *= 2.0;
	} else if (mode == GAS_FILTER_PEAK) {
		Q *= 3.0;
	}
	double tmpgain = gain;
	if (tmpgain < 0.001) {
		tmpgain = 0.001;
	}
	if (stages > 1) {
		Q = (Q > 1.0 ? pow(Q, 1.0 / stages) : Q);
		tmpgain = pow(tmpgain, 1.0 / (stages + 1));
	}
	double alpha = sin_v / (2 * Q);
	double a0 = 1.0 + alpha;
	float b0 = 0, b1 = 0, b2 = 0, a1 = 0, a2 = 0;
	switch (mode) {
		case GAS_FILTER_LOWPASS: {
			b0 = (float)((1.0 - cos_v) / 2.0);
			b1 = (float)(1.0 - cos_v);
			b2 = (float)((1.0 - cos_v) / 2.0);
			a1 = (float)(-2.0 * cos_v);
			a2 = (float)(1.0 - alpha);
		} break;
		case GAS_FILTER_HIGHPASS: {
			b0 = (float)((1.0 + cos_v) / 2.0);
			b1 = (float)(-(1.0 + cos_v));
			b2 = (float)((1.0 + cos_v) / 2.0);
			a1 = (float)(-2.0 * cos_v);
			a2 = (float)(1.0 - alpha);
		} break;
		case GAS_FILTER_BANDPASS: {
			b0 = (float)(alpha * sqrt(Q + 1));
			b1 = 0.0f;
			b2 = (float)(-alpha * sqrt(Q + 1));
			a1 = (float)(-2.0 * cos_v);
			a2 = (float)(1.0 - alpha);
		} break;
		case GAS_FILTER_NOTCH: {
			b0 = 1.0f;
			b1 = (float)(-2.0 * cos_v);
			b2 = 1.0f;
			a1 = (float)(-2.0 * cos_v);
			a2 = (float)(1.0 - alpha);
		} break;
		case GAS_FILTER_PEAK: {
			b0 = (float)(1.0 + alpha * tmpgain);
			b1 = (float)(-2.0 * cos_v);
			b2 = (float)(1.0 - alpha * tmpgain);
			a1 = (float)(-2 * cos_v);
			a2 = (float)(1 - alpha / tmpgain);
		} break;
		case GAS_FILTER_BANDLIMIT: {
			double hicutoff = resonance;
			double centercutoff = (cutoff + resonance) / 2.0;
			double bandwidth = (log(centercutoff) - log(hicutoff)) / log((double)2);
			omega = TAU * centercutoff / sampling_rate;
			alpha = sin(omega) * sinh(log((double)2) / 2 * bandwidth * omega / sin(omega));
			a0 = 1 + alpha;
			b0 = (float)alpha;
			b1 = 0;
			b2 = (float)-alpha;
			a1 = (float)(-2 * cos(omega));
			a2 = (float)(1 - alpha);
		} break;
		case GAS_FILTER_LOWSHELF: {
			double tmpq = sqrt(Q);
			if (tmpq <= 0) {
				tmpq = 0.001;
			}
			double beta = sqrt(tmpgain) / tmpq;
			a0 = (tmpgain + 1.0) + (tmpgain - 1.0) * cos_v + beta * sin_v;
			b0 = (float)(tmpgain * ((tmpgain + 1.0) - (tmpgain - 1.0) * cos_v + beta * sin_v));
			b1 = (float)(2.0 * tmpgain * ((tmpgain - 1.0) - (tmpgain + 1.0) * cos_v));
			b2 = (float)(tmpgain * ((tmpgain + 1.0) - (tmpgain - 1.0) * cos_v - beta * sin_v));
			a1 = (float)(-2.0 * ((tmpgain - 1.0) + (tmpgain + 1.0) * cos_v));
			a2 = (float)((tmpgain + 1.0) + (tmpgain - 1.0) * cos_v - beta * sin_v);
		} break;
		case GAS_FILTER_HIGHSHELF:
		default: {
			double tmpq = sqrt(Q);
			if (tmpq <= 0) {
				tmpq = 0.001;
			}
			double beta = sqrt(tmpgain) / tmpq;
			a0 = (tmpgain + 1.0) - (tmpgain - 1.0) * cos_v + beta * sin_v;
			b0 = (float)(tmpgain * ((tmpgain + 1.0) + (tmpgain - 1.0) * cos_v + beta * sin_v));
			b1 = (float)(-2.0 * tmpgain * ((tmpgain - 1.0) + (tmpgain + 1.0) * cos_v));
			b2 = (float)(tmpgain * ((tmpgain + 1.0) + (tmpgain - 1.0) * cos_v - beta * sin_v));
			a1 = (float)(2.0 * ((tmpgain - 1.0) - (tmpgain + 1.0) * cos_v));
			a2 = (float)((tmpgain + 1.0) - (tmpgain - 1.0) * cos_v - beta * sin_v);
		} break;
	}
	out[0] = (float)((double)b0 / a0);
	out[1] = (float)((double)b1 / a0);
	out[2] = (float)((double)b2 / a0);
	out[3] = (float)((double)a1 / (0.0 - a0));
	out[4] = (float)((double)a2 / (0.0 - a0));
}

/* ---------------------------------------------------------------------------------------------
 * per-voice mix functions, float32 (reference arithmetic) and float64 (shadow)
 * ------------------------------------------------------------------------------------------- */
typedef struct proc64 {
	double b0, b1, b2, a1, a2, ha1, ha2, hb1, hb2;
} proc64;
typedef struct vstate64 {
	double prev_mix_volumes[GAS_MAX_CHANNELS_PER_BUS][2];
	proc64 filter_processors[2 * GAS_MAX_CHANNELS_PER_BUS];
	double effect_history[GAS_MAX_EFFECTS][2][GAS_MAX_FILTER_STAGES][4];
} vstate64;
typedef struct frame64 {
	double l, r;
} frame64;

#define REAL float
#define SUF(n) n##_f32
#define PROC gas_processor_state
#define VSTATE gas_voice_state
#define FRAME gas_frame
#include "gas_oracle_mix.inc"
#undef REAL
#undef SUF
#undef PROC
#undef VSTATE
#undef FRAME

#define REAL double
#define SUF(n) n##_f64
#define PROC proc64
#define VSTATE vstate64
#define FRAME frame64
#include "gas_oracle_mix.inc"
#undef REAL
#undef SUF
#undef PROC
#undef VSTATE
#undef FRAME

void orc_process_frames_3d(const gas_params *p, gas_voice_state *st, float mix_rate, gas_frame *out, const gas_frame *src, int frames) {
	process_frames_3d_f32(p, st, mix_rate, out, src, frames);
}
void orc_mix_channel_3d(const gas_params *p, gas_voice_state *st, float mix_rate, int channel, gas_frame *out, const gas_frame *src, int frames) {
	mix_channel_3d_f32(p, st, mix_rate, channel, out, src, frames);
}
void orc_process_frames_effect(const gas_effect_chain *chain, gas_voice_state *st, float mix_rate, gas_frame *out, const gas_frame *src, int frames) {
	process_frames_effect_f32(chain, st, mix_rate, out, src, frames);
}

/* ---------------------------------------------------------------------------------------------
 * world
 * ------------------------------------------------------------------------------------------- */
typedef struct bus_details { /* upstream AudioStreamPlaybackBusDetails, shared by an instance's proxies */
	int n;
	int bus[GAS_MAX_BUSES_PER_PLAYBACK];
	float vol[GAS_MAX_BUSES_PER_PLAYBACK][GAS_MAX_CHANNELS_PER_BUS][2];
} bus_details;

typedef struct orc_instance {
	int spatializer;
	/* latched by AudioSpatializer3D::instantiate (audio_spatializer_3d.cpp:645-652: ins->mix_channel_mode = mix_channel_mode) */
	int kind, mix_channel_mode, effect_gain_binding;
	gas_params params;
	int was_further;
	int active; /* playback_active: proxies registered with AudioServer */
	bus_details cur, prev;
	gas_effect_chain fx;
} orc_instance;

struct orc_world {
	gas_config cfg;
	gas_spatializer *spat;
	orc_instance *inst;
	gas_voice_state *vs;
	vstate64 *vs64;
	/* start order of the voices: SafeList::insert puts a new SpatialPlaybackListNode at the HEAD of playback_list
	 * (audio_spatializer.cpp:74), so _mix_from_playback_list meets the most recently started voice first */
	uint64_t *start_seq;
	uint64_t next_seq;
	/* SpatialPlaybackListNode lifecycle (audio_spatializer.h:55-66), used by orc_mix_block_stream only */
	gas_frame (*lookahead)[GAS_LOOKAHEAD_BUFFER_SIZE];
	unsigned char *v_active, *v_has_frames;
	float *inst_threshold_db; /* playback_disable_threshold_db, audio_spatializer.h:87 */
	double last_mix_s, last_gain_s;
};

static double now_s(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void spat_defaults(gas_spatializer *s) { /* audio_spatializer_3d.h:171-188 */
	memset(s, 0, sizeof(*s));
	s->kind = GAS_SPATIALIZER_3D;
	s->attenuation_model = GAS_ATTENUATION_INVERSE_DISTANCE;
	s->unit_size = 10.0f;
	s->max_distance = 0.0f;
	s->panning_strength = 1.0f;
	s->area_mask = 1;
	s->emission_angle_enabled = 0;
	s->emission_angle = 45.0f;
	s->emission_angle_filter_attenuation_db = -12.0f;
	s->attenuation_filter_cutoff_hz = 5000.0f;
	s->attenuation_filter_db = -24.0f;
	s->doppler_tracking = GAS_DOPPLER_TRACKING_DISABLED;
	s->doppler_speed_of_sound = 343.0f;
	s->mix_channel_mode = 0;
	s->effect_gain_binding = -1;
}

static void params_defaults(gas_params *p) {
	memset(p, 0, sizeof(*p));
	p->pitch_scale = 1.0f;
	p->attenuation_filter_cutoff_hz = 5000.0f;
}

orc_world *orc_create(const gas_config *cfg) {
	if (!cfg || cfg->max_instances <= 0 || cfg->max_voices <= 0 || cfg->max_spatializers <= 0 ||
			cfg->num_buses < 1 || cfg->num_buses > GAS_MAX_BUSES || cfg->speaker_mode < 0 || cfg->speaker_mode > 3) {
		return NULL;
	}
	orc_world *w = (orc_world *)calloc(1, sizeof(orc_world));
	w->cfg = *cfg;
	w->spat = (gas_spatializer *)calloc(cfg->max_spatializers, sizeof(gas_spatializer));
	w->inst = (orc_instance *)calloc(cfg->max_instances, sizeof(orc_instance));
	w->vs = (gas_voice_state *)calloc(cfg->max_voices, sizeof(gas_voice_state));
	w->vs64 = (vstate64 *)calloc(cfg->max_voices, sizeof(vstate64));
	w->start_seq = (uint64_t *)calloc(cfg->max_voices, sizeof(uint64_t));
	w->next_seq = 1;
	w->lookahead = calloc(cfg->max_voices, sizeof(*w->lookahead));
	w->v_active = (unsigned char *)calloc(cfg->max_voices, 1);
	w->v_has_frames = (unsigned char *)calloc(cfg->max_voices, 1);
	w->inst_threshold_db = (float *)malloc(sizeof(float) * cfg->max_instances);
	for (int i = 0; i < cfg->max_instances; i++) {
		w->inst_threshold_db[i] = -80.0f;
	}
	for (int i = 0; i < cfg->max_spatializers; i++) {
		spat_defaults(&w->spat[i]);
	}
	for (int i = 0; i < cfg->max_instances; i++) {
		params_defaults(&w->inst[i].params);
	}
	return w;
}

void orc_destroy(orc_world *w) {
	if (!w) {
		return;
	}
	free(w->spat);
	free(w->inst);
	free(w->vs);
	free(w->vs64);
	free(w->start_seq);
	free(w->lookahead);
	free(w->v_active);
	free(w->v_has_frames);
	free(w->inst_threshold_db);
	free(w);
}

int orc_set_speaker_mode(orc_world *w, int mode) {
	if (mode < 0 || mode > 3) {
		return GAS_ERR_INVALID;
	}
	w->cfg.speaker_mode = mode;
	return GAS_OK;
}
int orc_set_mix_rate(orc_world *w, float hz) {
	if (!(hz > 0)) {
		return GAS_ERR_INVALID;
	}
	w->cfg.mix_rate = hz;
	return GAS_OK;
}
int orc_set_global_panning_strength(orc_world *w, float s) {
	w->cfg.global_panning_strength = s;
	return GAS_OK;
}

static int spat_valid(const gas_spatializer *s) { /* audio_spatializer_3d.cpp:670-672,695-697,728-730,737-739,758-760 */
	if (s->kind != GAS_SPATIALIZER_3D && s->kind != GAS_SPATIALIZER_EFFECT) {
		return 0;
	}
	if (s->max_distance < 0.0) {
		return 0;
	}
	if (s->emission_angle < 0 || s->emission_angle > 90) {
		return 0;
	}
	if (s->attenuation_model < 0 || s->attenuation_model >= 4) {
		return 0;
	}
	if (s->panning_strength < 0) {
		return 0;
	}
	if (s->doppler_speed_of_sound <= 0) {
		return 0;
	}
	if (s->chain.n_effects < 0 || s->chain.n_effects > GAS_MAX_EFFECTS) {
		return 0;
	}
	return 1;
}

int orc_spatializer_set(orc_world *w, int slot, const gas_spatializer *s) {
	if (slot < 0 || slot >= w->cfg.max_spatializers || !s || !spat_valid(s)) {
		return GAS_ERR_INVALID;
	}
	w->spat[slot] = *s;
	return GAS_OK;
}

static int inst_mix_channels(const orc_world *w, const orc_instance *q) {
	(void)w;
	return q->kind == GAS_SPATIALIZER_3D && q->mix_channel_mode;
}

/* get_bus_map for every proxy channel folded into one table: entry [bus][c] is what proxy c (Mode B)
 * or the single proxy (Mode A) sends to pair c of that bus (audio_spatializer.cpp:266-270, :295-319). */
static void inst_push_bus_map(const orc_world *w, orc_instance *q) {
	int mc = inst_mix_channels(w, q);
	bus_details d;
	memset(&d, 0, sizeof(d));
	if (mc) {
		for (int c = 0; c < GAS_MAX_CHANNELS_PER_BUS; c++) {
			int bus[6];
			float vol[6][4][2];
			int n = orc_get_bus_map(&q->params, 1, c, bus, vol);
			d.n = n;
			for (int k = 0; k < n; k++) {
				d.bus[k] = bus[k];
				d.vol[k][c][0] = vol[k][c][0];
				d.vol[k][c][1] = vol[k][c][1];
			}
		}
	} else {
		int bus[6];
		float vol[6][4][2];
		int n = orc_get_bus_map(&q->params, 0, 0, bus, vol);
		d.n = n;
		for (int k = 0; k < n; k++) {
			d.bus[k] = bus[k];
			memcpy(d.vol[k], vol[k], sizeof(float) * 8);
		}
	}
	q->cur = d;
}

int orc_instance_init(orc_world *w, int n, const int32_t *instances, const int32_t *spatializers) {
	for (int i = 0; i < n; i++) {
		if (instances[i] < 0 || instances[i] >= w->cfg.max_instances || spatializers[i] < 0 || spatializers[i] >= w->cfg.max_spatializers) {
			return GAS_ERR_INVALID;
		}
	}
	for (int i = 0; i < n; i++) {
		orc_instance *q = &w->inst[instances[i]];
		memset(q, 0, sizeof(*q));
		q->spatializer = spatializers[i];
		q->kind = w->spat[q->spatializer].kind;
		q->mix_channel_mode = w->spat[q->spatializer].mix_channel_mode;
		q->effect_gain_binding = w->spat[q->spatializer].effect_gain_binding;
		params_defaults(&q->params);
		q->fx = w->spat[q->spatializer].chain;
		w->inst_threshold_db[instances[i]] = -80.0f;
	}
	return GAS_OK;
}

int orc_instance_start(orc_world *w, int n, const int32_t *instances) {
	for (int i = 0; i < n; i++) {
		if (instances[i] < 0 || instances[i] >= w->cfg.max_instances) {
			return GAS_ERR_INVALID;
		}
	}
	for (int i = 0; i < n; i++) {
		orc_instance *q = &w->inst[instances[i]];
		q->active = 1;
		memset(&q->prev, 0, sizeof(q->prev));
		inst_push_bus_map(w, q);
	}
	return GAS_OK;
}

int orc_instance_stop(orc_world *w, int n, const int32_t *instances) {
	for (int i = 0; i < n; i++) {
		if (instances[i] < 0 || instances[i] >= w->cfg.max_instances) {
			return GAS_ERR_INVALID;
		}
	}
	for (int i = 0; i < n; i++) {
		w->inst[instances[i]].active = 0;
	}
	return GAS_OK;
}

int orc_voice_init(orc_world *w, int n, const int32_t *voices) {
	for (int i = 0; i < n; i++) {
		if (voices[i] < 0 || voices[i] >= w->cfg.max_voices) {
			return GAS_ERR_INVALID;
		}
	}
	for (int i = 0; i < n; i++) {
		memset(&w->vs[voices[i]], 0, sizeof(gas_voice_state));
		memset(&w->vs64[voices[i]], 0, sizeof(vstate64));
		w->start_seq[voices[i]] = w->next_seq++;
		/* start_playback_stream, audio_spatializer.cpp:57-72: lookahead zeroed, active and has_frames set */
		memset(w->lookahead[voices[i]], 0, sizeof(w->lookahead[0]));
		w->v_active[voices[i]] = 1;
		w->v_has_frames[voices[i]] = 1;
	}
	return GAS_OK;
}

static void inst_set_params(orc_world *w, orc_instance *q, const gas_params *p) {
	q->params = *p; /* set_spatializer_parameters, audio_spatializer.cpp:263 */
	if (p->update_parameters && q->active) { /* :265-271 (spatial_playbacks is empty while inactive) */
		inst_push_bus_map(w, q);
	}
}

int orc_gain_compute(orc_world *w, int n, const gas_emitter *emitters, int n_listeners, const gas_listener *listeners,
		int n_areas, const gas_area *areas, gas_params *out_params) {
	if (n < 0 || n_listeners < 0 || n_listeners > GAS_MAX_LISTENERS) {
		return GAS_ERR_INVALID;
	}
	for (int i = 0; i < n; i++) {
		const gas_emitter *e = &emitters[i];
		if (e->instance < 0 || e->instance >= w->cfg.max_instances || e->spatializer < 0 || e->spatializer >= w->cfg.max_spatializers ||
				e->area >= n_areas) {
			return GAS_ERR_INVALID;
		}
	}
	double t0 = now_s();
	for (int i = 0; i < n; i++) {
		const gas_emitter *e = &emitters[i];
		orc_instance *q = &w->inst[e->instance];
		gas_params p;
		orc_calculate_spatialization(&w->cfg, &w->spat[e->spatializer], e, n_listeners, listeners,
				e->area >= 0 ? &areas[e->area] : NULL, &q->was_further, &p);
		inst_set_params(w, q, &p);
		if (out_params) {
			out_params[i] = p;
		}
	}
	w->last_gain_s = now_s() - t0;
	return GAS_OK;
}

int orc_params_set(orc_world *w, int n, const int32_t *instances, const gas_params *params) {
	for (int i = 0; i < n; i++) {
		if (instances[i] < 0 || instances[i] >= w->cfg.max_instances || params[i].n_bus < 0 || params[i].n_bus > GAS_MAX_BUSES_PER_PLAYBACK) {
			return GAS_ERR_INVALID;
		}
	}
	for (int i = 0; i < n; i++) {
		gas_params p = params[i];
		for (int k = 0; k < p.n_bus; k++) {
			p.bus[k] = resolve_bus(&w->cfg, p.bus[k]);
		}
		inst_set_params(w, &w->inst[instances[i]], &p);
	}
	return GAS_OK;
}

int orc_params_get(orc_world *w, int n, const int32_t *instances, gas_params *out) {
	for (int i = 0; i < n; i++) {
		if (instances[i] < 0 || instances[i] >= w->cfg.max_instances) {
			return GAS_ERR_INVALID;
		}
		out[i] = w->inst[instances[i]].params;
	}
	return GAS_OK;
}

int orc_effect_params_set(orc_world *w, int n, const int32_t *instances, const gas_effect_chain *chains) {
	for (int i = 0; i < n; i++) {
		if (instances[i] < 0 || instances[i] >= w->cfg.max_instances || chains[i].n_effects < 0 || chains[i].n_effects > GAS_MAX_EFFECTS) {
			return GAS_ERR_INVALID;
		}
	}
	for (int i = 0; i < n; i++) {
		w->inst[instances[i]].fx = chains[i];
	}
	return GAS_OK;
}

/* One instance: AudioSpatializerInstance::_mix_from_playback_list (audio_spatializer.cpp:326-471) over
 * the voices idx[0..nv) followed by the AudioServer step of its proxies into `bus` (and `bus64`). */
typedef struct scratch {
	gas_frame *zero, *process, *temp, *mix[GAS_MAX_CHANNELS_PER_BUS];
	frame64 *src64, *process64, *temp64, *mix64[GAS_MAX_CHANNELS_PER_BUS];
} scratch;

static void scratch_alloc(scratch *s, int frames, int want64) {
	memset(s, 0, sizeof(*s));
	s->zero = (gas_frame *)calloc(frames, sizeof(gas_frame));
	s->process = (gas_frame *)calloc(frames, sizeof(gas_frame));
	s->temp = (gas_frame *)calloc(frames, sizeof(gas_frame));
	for (int c = 0; c < GAS_MAX_CHANNELS_PER_BUS; c++) {
		s->mix[c] = (gas_frame *)calloc(frames, sizeof(gas_frame));
	}
	if (want64) {
		s->src64 = (frame64 *)calloc(frames, sizeof(frame64));
		s->process64 = (frame64 *)calloc(frames, sizeof(frame64));
		s->temp64 = (frame64 *)calloc(frames, sizeof(frame64));
		for (int c = 0; c < GAS_MAX_CHANNELS_PER_BUS; c++) {
			s->mix64[c] = (frame64 *)calloc(frames, sizeof(frame64));
		}
	}
}
static void scratch_free(scratch *s) {
	free(s->zero);
	free(s->process);
	free(s->temp);
	free(s->src64);
	free(s->process64);
	free(s->temp64);
	for (int c = 0; c < GAS_MAX_CHANNELS_PER_BUS; c++) {
		free(s->mix[c]);
		free(s->mix64[c]);
	}
}

static int details_find(const bus_details *d, int bus) {
	for (int k = 0; k < d->n; k++) {
		if (d->bus[k] == bus) {
			return k;
		}
	}
	return -1;
}

static void mix_instance(orc_world *w, int qi, const gas_voice *voices, const int *idx, int nv, const gas_frame *src, int frames,
		gas_frame *bus, gas_frame *peaks, double *bus64, scratch *sc) {
	orc_instance *q = &w->inst[qi];
	const int kind = q->kind; /* latched at instantiate() */
	const int channels = w->cfg.speaker_mode + 1;
	const int mix_channels = inst_mix_channels(w, q);                    /* should_mix_channels */
	const int count = mix_channels ? channels : 1;                       /* init_channels_and_buffers, audio_spatializer.cpp:172-179 */
	const float mix_rate = w->cfg.mix_rate;
	const gas_params *prm = &q->params;                                  /* :328 */
	gas_effect_chain fx = q->fx;
	if (kind == GAS_SPATIALIZER_EFFECT && q->effect_gain_binding >= 0 && q->effect_gain_binding < fx.n_effects) {
		fx.effects[q->effect_gain_binding].gain = prm->linear_attenuation; /* example _process_effects, gd_spatializer_instance.gd:125-127 */
	}

	for (int c = 0; c < count; c++) { /* :335-343 */
		memset(sc->mix[c], 0, sizeof(gas_frame) * frames);
		if (bus64) {
			memset(sc->mix64[c], 0, sizeof(frame64) * frames);
		}
	}
	for (int j = 0; j < nv; j++) { /* :353 */
		const gas_voice *v = &voices[idx[j]];
		gas_voice_state *st = &w->vs[v->voice];
		vstate64 *st64 = &w->vs64[v->voice];
		const gas_frame *buf = v->src_row >= 0 ? src + (size_t)v->src_row * frames : sc->zero; /* :367-408 done by the caller */
		if (bus64) {
			for (int i = 0; i < frames; i++) {
				sc->src64[i].l = buf[i].l;
				sc->src64[i].r = buf[i].r;
			}
		}
		const gas_frame *processed = buf;
		const frame64 *processed64 = sc->src64;
		if (!mix_channels) { /* should_process_frames: Mode A and Effect, :411-417 */
			if (kind == GAS_SPATIALIZER_EFFECT) {
				process_frames_effect_f32(&fx, st, mix_rate, sc->process, buf, frames);
				if (bus64) {
					process_frames_effect_f64(&fx, st64, mix_rate, sc->process64, sc->src64, frames);
				}
			} else {
				process_frames_3d_f32(prm, st, mix_rate, sc->process, buf, frames);
				if (bus64) {
					process_frames_3d_f64(prm, st64, mix_rate, sc->process64, sc->src64, frames);
				}
			}
			processed = sc->process;
			processed64 = sc->process64;
		}
		gas_frame peak = { 0, 0 }; /* :419 */
		if (mix_channels) {        /* :421-445 */
			for (int c = 0; c < channels; c++) {
				mix_channel_3d_f32(prm, st, mix_rate, c, sc->temp, processed, frames);
				gas_frame *cb = sc->mix[c];
				for (int i = 0; i < frames; i++) {
					cb[i].l += sc->temp[i].l;
					cb[i].r += sc->temp[i].r;
					float l = fabsf(sc->temp[i].l);
					if (l > peak.l) {
						peak.l = l;
					}
					float r = fabsf(sc->temp[i].r);
					if (r > peak.r) {
						peak.r = r;
					}
				}
				if (bus64) {
					mix_channel_3d_f64(prm, st64, mix_rate, c, sc->temp64, processed64, frames);
					for (int i = 0; i < frames; i++) {
						sc->mix64[c][i].l += sc->temp64[i].l;
						sc->mix64[c][i].r += sc->temp64[i].r;
					}
				}
			}
		} else { /* :446-462 */
			gas_frame *ob = sc->mix[0];
			for (int i = 0; i < frames; i++) {
				ob[i].l += processed[i].l;
				ob[i].r += processed[i].r;
				float l = fabsf(processed[i].l);
				if (l > peak.l) {
					peak.l = l;
				}
				float r = fabsf(processed[i].r);
				if (r > peak.r) {
					peak.r = r;
				}
			}
			if (bus64) {
				for (int i = 0; i < frames; i++) {
					sc->mix64[0][i].l += processed64[i].l;
					sc->mix64[0][i].r += processed64[i].r;
				}
			}
		}
		if (peaks) {
			peaks[idx[j]] = peak;
		}
	}

	/* upstream AudioServer::_mix_step for this instance's proxy playbacks: every
	 * active bus slot is mixed with the previous volume looked up by bus (absent => 0 => fade-in), buses
	 * only present in the previous details are mixed once more towards 0, then prev <- cur. */
	if (bus) {
		for (int k = 0; k < q->cur.n; k++) {
			int b = resolve_bus(&w->cfg, q->cur.bus[k]);
			int pk = details_find(&q->prev, q->cur.bus[k]);
			for (int c = 0; c < channels; c++) {
				float ps[2] = { 0, 0 };
				if (pk >= 0) {
					ps[0] = q->prev.vol[pk][c][0];
					ps[1] = q->prev.vol[pk][c][1];
				}
				const int sc_idx = mix_channels ? c : 0;
				server_mix_step_for_channel_f32(bus + ((size_t)b * channels + c) * frames, sc->mix[sc_idx], ps[0], ps[1],
						q->cur.vol[k][c][0], q->cur.vol[k][c][1], frames);
				if (bus64) {
					server_mix_step_for_channel_f64((frame64 *)bus64 + ((size_t)b * channels + c) * frames, sc->mix64[sc_idx], ps[0], ps[1],
							q->cur.vol[k][c][0], q->cur.vol[k][c][1], frames);
				}
			}
		}
		for (int pk = 0; pk < q->prev.n; pk++) {
			if (details_find(&q->cur, q->prev.bus[pk]) >= 0) {
				continue;
			}
			int b = resolve_bus(&w->cfg, q->prev.bus[pk]);
			for (int c = 0; c < channels; c++) {
				const int sc_idx = mix_channels ? c : 0;
				server_mix_step_for_channel_f32(bus + ((size_t)b * channels + c) * frames, sc->mix[sc_idx],
						q->prev.vol[pk][c][0], q->prev.vol[pk][c][1], 0, 0, frames);
				if (bus64) {
					server_mix_step_for_channel_f64((frame64 *)bus64 + ((size_t)b * channels + c) * frames, sc->mix64[sc_idx],
							q->prev.vol[pk][c][0], q->prev.vol[pk][c][1], 0, 0, frames);
				}
			}
		}
		/* Mode B: proxy c' is a playback of its own whose bus map is masked to pair c' (audio_spatializer.cpp:298-312),
		 * but AudioServer still runs _mix_step_for_channel for EVERY pair of every bus of the map, with volume 0 for the
		 * masked-out pairs.  0 * x only matters when x is not finite — which the module produces itself (Q1: NaN pan
		 * gains) — and then the NaN reaches every pair of the bus, same side.  Adding +-0 is a no-op otherwise, so the
		 * cross terms are only run for pair buffers that hold a non-finite sample.  Pinned by oracle/_ref. */
		if (mix_channels) {
			for (int cp = 0; cp < channels; cp++) {
				int bad = 0;
				for (int i = 0; i < frames && !bad; i++) {
					bad = !isfinite(sc->mix[cp][i].l) || !isfinite(sc->mix[cp][i].r);
				}
				if (!bad) {
					continue;
				}
				for (int pass = 0; pass < 2; pass++) {
					const bus_details *d = pass == 0 ? &q->cur : &q->prev;
					for (int k = 0; k < d->n; k++) {
						if (pass == 1 && details_find(&q->cur, d->bus[k]) >= 0) {
							continue;
						}
						int b = resolve_bus(&w->cfg, d->bus[k]);
						for (int c = 0; c < channels; c++) {
							if (c == cp) {
								continue;
							}
							server_mix_step_for_channel_f32(bus + ((size_t)b * channels + c) * frames, sc->mix[cp], 0, 0, 0, 0, frames);
							if (bus64) {
								server_mix_step_for_channel_f64((frame64 *)bus64 + ((size_t)b * channels + c) * frames, sc->mix64[cp], 0, 0, 0, 0, frames);
							}
						}
					}
				}
			}
		}
	}
	q->prev = q->cur;
}

static int cmp_int(const void *a, const void *b) {
	return *(const int *)a - *(const int *)b;
}

int orc_mix_block(orc_world *w, int n_voices, const gas_voice *voices, const gas_frame *src, int src_rows,
		int frames, gas_frame *bus_out, gas_frame *peaks, double *bus_out64, int threads) {
	if (n_voices < 0 || frames <= 0 || (frames & 1) || frames > w->cfg.max_frames) {
		return GAS_ERR_INVALID;
	}
	for (int i = 0; i < n_voices; i++) {
		if (voices[i].voice < 0 || voices[i].voice >= w->cfg.max_voices || voices[i].instance < 0 ||
				voices[i].instance >= w->cfg.max_instances || voices[i].src_row >= src_rows) {
			return GAS_ERR_INVALID;
		}
	}
	const int channels = w->cfg.speaker_mode + 1;
	const size_t bus_frames = (size_t)w->cfg.num_buses * channels * frames;
	double t0 = now_s();
	if (bus_out) {
		memset(bus_out, 0, bus_frames * sizeof(gas_frame));
	}
	if (bus_out64) {
		memset(bus_out64, 0, bus_frames * 2 * sizeof(double));
	}
	if (peaks) {
		memset(peaks, 0, sizeof(gas_frame) * n_voices);
	}

	/* group voices by instance, keeping list order inside an instance */
	const int ni = w->cfg.max_instances;
	int *count = (int *)calloc(ni + 1, sizeof(int));
	for (int i = 0; i < n_voices; i++) {
		count[voices[i].instance + 1]++;
	}
	for (int i = 0; i < ni; i++) {
		count[i + 1] += count[i];
	}
	int *fill = (int *)malloc(sizeof(int) * (ni + 1));
	memcpy(fill, count, sizeof(int) * (ni + 1));
	int *order = (int *)malloc(sizeof(int) * (n_voices > 0 ? n_voices : 1));
	for (int i = 0; i < n_voices; i++) {
		order[fill[voices[i].instance]++] = i;
	}
	/* inside an instance: most recently started voice first (the order only fixes the float summation order) */
	for (int q = 0; q < ni; q++) {
		for (int a = count[q] + 1; a < count[q + 1]; a++) {
			int v = order[a];
			int b = a - 1;
			while (b >= count[q] && w->start_seq[voices[order[b]].voice] < w->start_seq[voices[v].voice]) {
				order[b + 1] = order[b];
				b--;
			}
			order[b + 1] = v;
		}
	}
	/* instances to step: every active instance (its proxies are mixed by AudioServer every step) */
	int *todo = (int *)malloc(sizeof(int) * ni);
	int ntodo = 0;
	for (int i = 0; i < ni; i++) {
		if (w->inst[i].active) {
			todo[ntodo++] = i;
		}
	}
	qsort(todo, ntodo, sizeof(int), cmp_int);

	if (threads <= 1) {
		scratch sc;
		scratch_alloc(&sc, frames, bus_out64 != NULL);
		for (int t = 0; t < ntodo; t++) {
			int qi = todo[t];
			mix_instance(w, qi, voices, order + count[qi], count[qi + 1] - count[qi], src, frames, bus_out, peaks, bus_out64, &sc);
		}
		scratch_free(&sc);
	} else {
#ifdef _OPENMP
		gas_frame *partials = (gas_frame *)calloc(bus_frames * threads, sizeof(gas_frame));
#pragma omp parallel num_threads(threads)
		{
			int tid = omp_get_thread_num();
			scratch sc;
			scratch_alloc(&sc, frames, 0);
			gas_frame *mine = partials + bus_frames * tid;
#pragma omp for schedule(static)
			for (int t = 0; t < ntodo; t++) {
				int qi = todo[t];
				mix_instance(w, qi, voices, order + count[qi], count[qi + 1] - count[qi], src, frames, bus_out ? mine : NULL, peaks, NULL, &sc);
			}
			scratch_free(&sc);
		}
		if (bus_out) {
			for (int t = 0; t < threads; t++) {
				const gas_frame *p = partials + bus_frames * t;
				for (size_t i = 0; i < bus_frames; i++) {
					bus_out[i].l += p[i].l;
					bus_out[i].r += p[i].r;
				}
			}
		}
		free(partials);
#else
		free(count);
		free(fill);
		free(order);
		free(todo);
		return GAS_ERR_STATE;
#endif
	}
	free(count);
	free(fill);
	free(order);
	free(todo);
	w->last_mix_s = now_s() - t0;
	return GAS_OK;
}

int orc_set_playback_disable_threshold_db(orc_world *w, int n, const int32_t *instances, const float *db) {
	for (int i = 0; i < n; i++) {
		if (instances[i] < 0 || instances[i] >= w->cfg.max_instances) {
			return GAS_ERR_INVALID;
		}
	}
	for (int i = 0; i < n; i++) {
		w->inst_threshold_db[instances[i]] = db[i];
	}
	return GAS_OK;
}

/* AudioSpatializerInstance::_mix_from_playback_list around the per-voice call, audio_spatializer.cpp:353-408 and :464-469:
 * row r of `src` holds the mixed_frames[i] frames AudioStreamPlayback::mix returned for voice i this block (:378).
 *   :355      inactive voices are skipped
 *   :369-373  the lookahead (last 64 frames of the previous mix) goes in front, the new frames behind it: the block that is
 *             processed is buf[0..F), the new lookahead buf[F..F+64)
 *   :380-398  a short mix ends the stream: over the last 64 valid frames  coef *= 0.96; buf[idx] *= coef * (64 - k) / 64,
 *             zeros after them; has_frames is cleared
 *   :405-408  without frames the voice is processed with a zero-filled buffer (filter tails)
 *   :464-469  a voice without frames is deactivated once its block peak <= db_to_linear(playback_disable_threshold_db)
 * status_out[i]: bit 0 = active after the block, bit 1 = has_frames after the block. */
int orc_mix_block_stream(orc_world *w, int n_voices, const gas_voice *voices, const gas_frame *src, int src_rows, int frames,
		const int32_t *mixed_frames, gas_frame *bus_out, gas_frame *peaks, int32_t *status_out, int threads) {
	const int L = GAS_LOOKAHEAD_BUFFER_SIZE;
	if (n_voices < 0 || frames <= 0 || (frames & 1) || frames > w->cfg.max_frames || !mixed_frames) {
		return GAS_ERR_INVALID;
	}
	for (int i = 0; i < n_voices; i++) {
		if (voices[i].voice < 0 || voices[i].voice >= w->cfg.max_voices || voices[i].instance < 0 || voices[i].instance >= w->cfg.max_instances ||
				voices[i].src_row >= src_rows || mixed_frames[i] < 0 || mixed_frames[i] > frames) {
			return GAS_ERR_INVALID;
		}
	}
	gas_frame *stage = (gas_frame *)calloc((size_t)(n_voices > 0 ? n_voices : 1) * frames, sizeof(gas_frame));
	gas_voice *vl = (gas_voice *)calloc(n_voices > 0 ? n_voices : 1, sizeof(gas_voice));
	int *map = (int *)malloc(sizeof(int) * (n_voices > 0 ? n_voices : 1));
	gas_frame *buf = (gas_frame *)malloc(sizeof(gas_frame) * (size_t)(frames + L));
	int nl = 0;
	for (int i = 0; i < n_voices; i++) {
		const int v = voices[i].voice;
		if (!w->v_active[v]) { /* :355 */
			continue;
		}
		gas_voice o = voices[i];
		o.flags |= GAS_VOICE_WANT_PEAK; /* :419: the reference always tracks the peak */
		if (w->v_has_frames[v]) {
			const int mixed = voices[i].src_row >= 0 ? mixed_frames[i] : 0;
			const gas_frame *row = voices[i].src_row >= 0 ? src + (size_t)voices[i].src_row * frames : NULL;
			for (int k = 0; k < L; k++) { /* :371-373 */
				buf[k] = w->lookahead[v][k];
			}
			for (int k = 0; k < frames; k++) { /* :378 — frames the playback did not deliver keep whatever the buffer held */
				if (k < mixed) {
					buf[L + k] = row[k];
				} else {
					buf[L + k].l = buf[L + k].r = 0.0f;
				}
			}
			if (mixed != frames) { /* :380-398 */
				float fadeout_base = (float)0.96;
				float fadeout_coefficient = 1;
				float buffer_size_float = (float)L;
				float buffer_linear_fade_idx = 0.0f;
				int fade_limit = mixed + L;
				for (int idx = mixed; idx < frames; idx++) {
					if (idx < fade_limit) {
						fadeout_coefficient *= fadeout_base;
						float f = fadeout_coefficient * (buffer_size_float - buffer_linear_fade_idx) / buffer_size_float;
						buf[idx].l *= f;
						buf[idx].r *= f;
						buffer_linear_fade_idx += 1.0f;
					} else {
						buf[idx].l *= 0.0f;
						buf[idx].r *= 0.0f;
					}
				}
				w->v_has_frames[v] = 0;
			} else {
				for (int k = 0; k < L; k++) { /* :401-403 */
					w->lookahead[v][k] = buf[frames + k];
				}
			}
			memcpy(stage + (size_t)nl * frames, buf, sizeof(gas_frame) * frames);
			o.src_row = nl;
		} else {
			o.src_row = -1; /* :405-408 */
		}
		map[nl] = i;
		vl[nl++] = o;
	}
	gas_frame *pk = (gas_frame *)calloc(nl > 0 ? nl : 1, sizeof(gas_frame));
	int st = orc_mix_block(w, nl, vl, stage, nl, frames, bus_out, pk, NULL, threads);
	if (status_out) {
		for (int i = 0; i < n_voices; i++) {
			status_out[i] = 0;
		}
	}
	if (peaks) {
		memset(peaks, 0, sizeof(gas_frame) * n_voices);
	}
	if (st == GAS_OK) {
		for (int j = 0; j < nl; j++) {
			const int v = vl[j].voice;
			if (peaks) {
				peaks[map[j]] = pk[j];
			}
			if (!w->v_has_frames[v]) { /* :464-469 */
				float m = pk[j].r > pk[j].l ? pk[j].r : pk[j].l;
				if (m <= orc_db_to_linear_f(w->inst_threshold_db[vl[j].instance])) {
					w->v_active[v] = 0;
				}
			}
			if (status_out) {
				status_out[map[j]] = (w->v_active[v] ? 1 : 0) | (w->v_has_frames[v] ? 2 : 0);
			}
		}
	}
	free(pk);
	free(stage);
	free(vl);
	free(map);
	free(buf);
	return st;
}

int orc_voice_state_export(orc_world *w, int n, const int32_t *voices, gas_voice_state *out) {
	for (int i = 0; i < n; i++) {
		if (voices[i] < 0 || voices[i] >= w->cfg.max_voices) {
			return GAS_ERR_INVALID;
		}
		out[i] = w->vs[voices[i]];
	}
	return GAS_OK;
}

int orc_voice_state_import(orc_world *w, int n, const int32_t *voices, const gas_voice_state *in) {
	for (int i = 0; i < n; i++) {
		if (voices[i] < 0 || voices[i] >= w->cfg.max_voices) {
			return GAS_ERR_INVALID;
		}
	}
	for (int i = 0; i < n; i++) {
		gas_voice_state *d = &w->vs[voices[i]];
		vstate64 *e = &w->vs64[voices[i]];
		*d = in[i];
		for (int c = 0; c < 4; c++) {
			e->prev_mix_volumes[c][0] = d->prev_mix_volumes[c][0];
			e->prev_mix_volumes[c][1] = d->prev_mix_volumes[c][1];
		}
		for (int k = 0; k < 8; k++) {
			const gas_processor_state *p = &d->filter_processors[k];
			proc64 *r = &e->filter_processors[k];
			r->b0 = p->b0;
			r->b1 = p->b1;
			r->b2 = p->b2;
			r->a1 = p->a1;
			r->a2 = p->a2;
			r->ha1 = p->ha1;
			r->ha2 = p->ha2;
			r->hb1 = p->hb1;
			r->hb2 = p->hb2;
		}
		for (int a = 0; a < GAS_MAX_EFFECTS; a++) {
			for (int b = 0; b < 2; b++) {
				for (int c = 0; c < GAS_MAX_FILTER_STAGES; c++) {
					for (int k = 0; k < 4; k++) {
						e->effect_history[a][b][c][k] = d->effect_history[a][b][c][k];
					}
				}
			}
		}
	}
	return GAS_OK;
}

double orc_last_mix_seconds(const orc_world *w) {
	return w->last_mix_s;
}
double orc_last_gain_seconds(const orc_world *w) {
	return w->last_gain_s;
}
int orc_max_threads(void) {
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

size_t orc_sizeof(int32_t id) {
	switch (id) {
		case GAS_STRUCT_FRAME: return sizeof(gas_frame);
		case GAS_STRUCT_EFFECT: return sizeof(gas_effect);
		case GAS_STRUCT_EFFECT_CHAIN: return sizeof(gas_effect_chain);
		case GAS_STRUCT_SPATIALIZER: return sizeof(gas_spatializer);
		case GAS_STRUCT_LISTENER: return sizeof(gas_listener);
		case GAS_STRUCT_AREA: return sizeof(gas_area);
		case GAS_STRUCT_EMITTER: return sizeof(gas_emitter);
		case GAS_STRUCT_PARAMS: return sizeof(gas_params);
		case GAS_STRUCT_VOICE: return sizeof(gas_voice);
		case GAS_STRUCT_PROCESSOR_STATE: return sizeof(gas_processor_state);
		case GAS_STRUCT_VOICE_STATE: return sizeof(gas_voice_state);
		case GAS_STRUCT_CONFIG: return sizeof(gas_config);
		case GAS_STRUCT_VOICE_LIFE: return sizeof(gas_voice_life);
		case GAS_STRUCT_BUS_DESC: return sizeof(gas_bus_desc);
		case GAS_STRUCT_STEP_NEXT: return sizeof(gas_step_next);
		default: return 0;
	}
}

/* ---------------------------------------------------------------------------------------------
 * The resampler in front of the path: what `playback->stream_playback->mix(&buf[64], pitch_scale, n)`
 * (reference audio_spatializer.cpp:375-378) runs for a resampled stream — upstream
 * AudioStreamPlaybackResampled::begin_resample / ::mix (servers/audio/audio_stream.cpp), restated literally from Godot
 * 4.x AS RECALLED: the engine is not in the reference tree, so THIS PART OF THE ORACLE IS NOT PINNED by reference code
 * (parity for it means GPU == this restatement).  The stream behind it is plain PCM the way AudioStreamPlaybackWAV
 * delivers it through _mix_internal: forward, optional loop over the whole stream, silence and `active = false` after
 * the end.
 * ------------------------------------------------------------------------------------------- */
#define ORC_FP_BITS 16
#define ORC_FP_LEN (1 << ORC_FP_BITS)
#define ORC_FP_MASK (ORC_FP_LEN - 1)
#define ORC_INTERNAL_BUFFER_LEN 128
#define ORC_CUBIC_INTERP_HISTORY 4

struct orc_resampler {
	gas_frame internal_buffer[ORC_INTERNAL_BUFFER_LEN + ORC_CUBIC_INTERP_HISTORY];
	unsigned int internal_buffer_end; /* upstream: unsigned; -1 = the buffer holds no end of stream */
	uint64_t mix_offset;
	const gas_frame *pcm;
	int n_frames, loop;
	long cursor; /* next source frame _mix_internal delivers */
	int playing;
	float sample_rate;
};

/* AudioStreamPlaybackWAV-like _mix_internal: up to n frames; after the end of a non-looping stream the rest is silence,
 * the playback stops and the count of real frames is returned */
static int orc_rs_mix_internal(orc_resampler *r, gas_frame *buf, int n) {
	int mixed = 0;
	for (int i = 0; i < n; i++) {
		if (r->cursor >= r->n_frames) {
			if (r->loop && r->n_frames > 0) {
				r->cursor = 0;
			} else {
				r->playing = 0;
				for (; i < n; i++) {
					buf[i].l = buf[i].r = 0.f;
				}
				return mixed;
			}
		}
		buf[i] = r->pcm[r->cursor++];
		mixed++;
	}
	return mixed;
}

orc_resampler *orc_resampler_begin(const gas_frame *pcm, int n_frames, int loop, float sample_rate, int start_frame) {
	orc_resampler *r = (orc_resampler *)calloc(1, sizeof(orc_resampler));
	if (!r) {
		return NULL;
	}
	r->pcm = pcm;
	r->n_frames = n_frames;
	r->loop = loop;
	r->sample_rate = sample_rate;
	r->cursor = start_frame;
	r->playing = 1;
	r->internal_buffer_end = (unsigned int)-1;
	/* begin_resample(): clear the cubic interpolation history, mix the first buffer (its return value is not looked at) */
	for (int i = 0; i < ORC_CUBIC_INTERP_HISTORY; i++) {
		r->internal_buffer[i].l = r->internal_buffer[i].r = 0.f;
	}
	orc_rs_mix_internal(r, r->internal_buffer + ORC_CUBIC_INTERP_HISTORY, ORC_INTERNAL_BUFFER_LEN);
	r->mix_offset = 0;
	return r;
}

void orc_resampler_free(orc_resampler *r) { free(r); }

int orc_resampler_mix(orc_resampler *r, gas_frame *p_buffer, float p_rate_scale, float target_rate, int p_frames) {
	const float playback_speed_scale = 1.0f;
	uint64_t mix_increment = (uint64_t)(((r->sample_rate * p_rate_scale * playback_speed_scale) / (double)target_rate) * (double)ORC_FP_LEN);
	int mixed_frames_total = -1;
	int i;
	for (i = 0; i < p_frames; i++) {
		uint32_t idx = ORC_CUBIC_INTERP_HISTORY + (uint32_t)(r->mix_offset >> ORC_FP_BITS);
		float mu = (r->mix_offset & ORC_FP_MASK) / (float)ORC_FP_LEN;
		gas_frame y0 = r->internal_buffer[idx - 3];
		gas_frame y1 = r->internal_buffer[idx - 2];
		gas_frame y2 = r->internal_buffer[idx - 1];
		gas_frame y3 = r->internal_buffer[idx - 0];
		if (idx >= r->internal_buffer_end && mixed_frames_total == -1) {
			mixed_frames_total = i;
		}
		float mu2 = mu * mu;
		float h11 = mu2 * (mu - 1);
		float z = mu2 - h11;
		float h01 = z - h11;
		float h10 = mu - z;
		/* p_buffer[i] = y1 + (y2 - y1) * h01 + ((y2 - y0) * h10 + (y3 - y1) * h11) * 0.5; */
		{
			float al = (y2.l - y1.l) * h01, ar = (y2.r - y1.r) * h01;
			float bl = ((y2.l - y0.l) * h10 + (y3.l - y1.l) * h11) * 0.5f, br = ((y2.r - y0.r) * h10 + (y3.r - y1.r) * h11) * 0.5f;
			p_buffer[i].l = (y1.l + al) + bl;
			p_buffer[i].r = (y1.r + ar) + br;
		}
		r->mix_offset += mix_increment;
		while ((r->mix_offset >> ORC_FP_BITS) >= ORC_INTERNAL_BUFFER_LEN) {
			for (int k = 0; k < ORC_CUBIC_INTERP_HISTORY; k++) {
				r->internal_buffer[k] = r->internal_buffer[ORC_INTERNAL_BUFFER_LEN + k];
			}
			if (r->playing) {
				int mixed_frames = orc_rs_mix_internal(r, r->internal_buffer + ORC_CUBIC_INTERP_HISTORY, ORC_INTERNAL_BUFFER_LEN);
				if (mixed_frames != ORC_INTERNAL_BUFFER_LEN) {
					r->internal_buffer_end = (unsigned int)mixed_frames;
				} else {
					r->internal_buffer_end = (unsigned int)-1;
				}
			} else {
				for (int j = 0; j < ORC_INTERNAL_BUFFER_LEN; j++) {
					r->internal_buffer[j + ORC_CUBIC_INTERP_HISTORY].l = r->internal_buffer[j + ORC_CUBIC_INTERP_HISTORY].r = 0.f;
				}
			}
			r->mix_offset -= ((uint64_t)ORC_INTERNAL_BUFFER_LEN << ORC_FP_BITS);
		}
	}
	if (mixed_frames_total == -1 && i == p_frames) {
		mixed_frames_total = p_frames;
	}
	return mixed_frames_total;
}

/* ---------------------------------------------------------------------------------------------
 * The bus graph after the mix: upstream AudioServer::_mix_step from "process send" on, restated from
 * Godot 4.x AS RECALLED (not pinned by reference code): buses from the last to the first; volume = db_to_linear(volume_db),
 * 0 when muted (no bus soloed) or not on a soloed chain (some bus soloed); buf *= volume; send buffer += buf.
 * bus: [n_buses][channels][frames] AudioFrames, in place.
 * ------------------------------------------------------------------------------------------- */
void orc_bus_graph(int n_buses, int channels, int frames, const float *volume_db, const int32_t *mute, const int32_t *solo, const int32_t *send_in,
		gas_frame *bus) {
	int send[GAS_MAX_BUSES] = { 0 }, soloed[GAS_MAX_BUSES] = { 0 };
	int solo_mode = 0;
	for (int b = 0; b < n_buses; b++) {
		int t = send_in[b];
		send[b] = (b > 0 && t >= 0 && t < b) ? t : 0; /* an invalid send goes to Master */
		if (solo[b]) {
			solo_mode = 1;
		}
	}
	if (solo_mode) {
		for (int b = 0; b < n_buses; b++) {
			if (solo[b]) {
				int i = b;
				soloed[i] = 1;
				while (i != 0) {
					i = send[i];
					soloed[i] = 1;
				}
			}
		}
	}
	for (int b = n_buses - 1; b >= 0; b--) {
		float volume = orc_db_to_linear_f(volume_db[b]);
		if (solo_mode) {
			if (!soloed[b]) {
				volume = 0.0f;
			}
		} else if (mute[b]) {
			volume = 0.0f;
		}
		for (int k = 0; k < channels; k++) {
			gas_frame *buf = bus + ((size_t)b * channels + k) * frames;
			for (int j = 0; j < frames; j++) {
				buf[j].l *= volume;
				buf[j].r *= volume;
			}
			if (b > 0) {
				gas_frame *target = bus + ((size_t)send[b] * channels + k) * frames;
				for (int j = 0; j < frames; j++) {
					target[j].l += buf[j].l;
					target[j].r += buf[j].r;
				}
			}
		}
	}
}
