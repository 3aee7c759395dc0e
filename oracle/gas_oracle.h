/*
 * gas_oracle.h — CPU oracle: a scalar restatement of the reference's hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under godot-audio-spatializer_b200/ may include, link or call this;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * PARITY PINNED BY REFERENCE CODE (module side) / RECALLED (upstream side):
 *  - every module-side line of the path is pinned bit for bit against the reference's own sources:
 *    oracle/_ref/libgas_ref.so is the .cpp files of /root/reference, UNMODIFIED, compiled against the godot-lite stand-in
 *    headers (oracle/godot_lite/) by `make -C oracle ref` and driven by oracle/ref_harness.cpp;
 *    tests/test_oracle_vs_ref.py compares oracle == _ref on float32 bit patterns (gains, bus maps, bus buffers,
 *    persistent voice state, NaN cases included), tests/test_lifecycle.py does the same for the lookahead / end
 *    fade / deactivation path, and the .npz files under tests/golden/ are outputs of _ref (tests/golden/make_golden.py);
 *  - the upstream-Godot pieces that are not under /root/reference (AudioFilterSW, AudioServer::_mix_step /
 *    _mix_step_for_channel, Math::*, Basis/Transform3D, AudioEffectFilter; godotengine/godot 4.x — the example
 *    project declares feature "4.6", no commit is pinned) are restated from Godot 4.x as recalled (SURVEY.md
 *    Appendix A), twice and independently (C here, C++ in oracle/godot_lite/), and the two restatements are checked
 *    against each other.  That part remains unpinned by executed upstream code: there is no Godot tree here.
 *
 * The oracle uses the same POD records as the C ABI (include/gas.h) so that a parity test drives both
 * with identical call sequences.
 */
#ifndef GAS_ORACLE_H
#define GAS_ORACLE_H

#include "../include/gas.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_world orc_world;

/* ---- scalar pieces (exported for known-answer tests) ------------------------------------------- */
float orc_db_to_linear_f(float db);       /* upstream Math::db_to_linear(float) */
double orc_db_to_linear_d(double db);
float orc_linear_to_db_f(float lin);
double orc_linear_to_db_d(double lin);
/* audio_spatializer_3d.cpp:123-151 */
float orc_get_attenuation_db(const gas_spatializer *s, float volume_db, float max_db, float distance);
/* audio_spatializer_3d.cpp:103-110 */
void orc_calc_output_vol_stereo(const float dir[3], float pan_strength, float out[4][2]);
/* audio_spatializer_3d.cpp:57-98 + :903-938 ; speaker_mode != STEREO */
void orc_calc_output_vol_surround(int speaker_mode, const float dir[3], float tightness, float out[4][2]);
/* audio_spatializer_3d.cpp:112-121 */
void orc_calc_output_vol(int speaker_mode, float global_panning, float panning_strength, const float dir[3], float out[4][2]);
/* SPCAP pieces, audio_spatializer_3d.cpp:47-55, :903-916, :926-938 */
void orc_spcap_effective_speakers(int speaker_count, float eff[7]);
void orc_spcap_calculate(int speaker_count, const float dir[3], float tightness, float volumes[7]);
/* upstream AudioFilterSW::prepare_coefficients; out = {b0,b1,b2,a1,a2} (a's negated) */
void orc_filter_prepare_coefficients(int mode, float cutoff, float resonance, float gain, int stages,
		float sampling_rate, float out[5]);
/* audio_spatializer_3d.cpp:277-489 for one emitter.  `was_further` is the instance's
 * was_further_than_max_distance_last_frame, updated in place. */
void orc_calculate_spatialization(const gas_config *cfg, const gas_spatializer *s, const gas_emitter *e,
		int n_listeners, const gas_listener *listeners, const gas_area *area, int *was_further, gas_params *out);
/* audio_spatializer.cpp:274-324 for one proxy channel.  out_bus/out_vol sized 6 / [6][4][2]. */
int orc_get_bus_map(const gas_params *p, int mix_channels, int channel, int out_bus[6], float out_vol[6][4][2]);
/* audio_spatializer_3d.cpp:491-552 / :554-609 on a single voice state (float32). */
void orc_process_frames_3d(const gas_params *p, gas_voice_state *st, float mix_rate,
		gas_frame *out, const gas_frame *src, int frames);
void orc_mix_channel_3d(const gas_params *p, gas_voice_state *st, float mix_rate, int channel,
		gas_frame *out, const gas_frame *src, int frames);
/* audio_spatializer_effect.cpp:33-77 with AudioEffectFilter instances. */
void orc_process_frames_effect(const gas_effect_chain *chain, gas_voice_state *st, float mix_rate,
		gas_frame *out, const gas_frame *src, int frames);

/* ---- world: mirrors the C ABI call for call ------------------------------------------------------- */
orc_world *orc_create(const gas_config *cfg);
void orc_destroy(orc_world *w);
int orc_set_speaker_mode(orc_world *w, int mode);
int orc_set_mix_rate(orc_world *w, float hz);
int orc_set_global_panning_strength(orc_world *w, float s);
int orc_spatializer_set(orc_world *w, int slot, const gas_spatializer *s);
int orc_instance_init(orc_world *w, int n, const int32_t *instances, const int32_t *spatializers);
int orc_instance_start(orc_world *w, int n, const int32_t *instances);
int orc_instance_stop(orc_world *w, int n, const int32_t *instances);
int orc_voice_init(orc_world *w, int n, const int32_t *voices);
int orc_gain_compute(orc_world *w, int n, const gas_emitter *emitters, int n_listeners, const gas_listener *listeners,
		int n_areas, const gas_area *areas, gas_params *out_params);
int orc_params_set(orc_world *w, int n, const int32_t *instances, const gas_params *params);
int orc_params_get(orc_world *w, int n, const int32_t *instances, gas_params *out);
int orc_effect_params_set(orc_world *w, int n, const int32_t *instances, const gas_effect_chain *chains);
/* Reference loop structure: per instance, per voice, temp buffer then accumulate
 * (audio_spatializer.cpp:326-471), then the AudioServer ramped bus accumulate.  threads <= 1 is the
 * faithful single audio thread; threads > 1 runs instances in parallel with OpenMP (per-thread bus
 * partials summed in thread order) and is used only as the "all host cores" CPU baseline.
 * bus_out64 (optional): float64 shadow of the same block (double ramps, recurrences and sums; own
 * shadow state inside the world) for error budgeting. */
int orc_mix_block(orc_world *w, int n_voices, const gas_voice *voices, const gas_frame *src, int src_rows,
		int frames, gas_frame *bus_out, gas_frame *peaks, double *bus_out64, int threads);
/* Stream form: the voice lifecycle of _mix_from_playback_list (lookahead splice, end-of-stream fade, zero-input tails,
 * deactivation below the threshold), audio_spatializer.cpp:353-408, :464-469.  See gas_mix_block_stream in gas.h. */
int orc_mix_block_stream(orc_world *w, int n_voices, const gas_voice *voices, const gas_frame *src, int src_rows, int frames,
		const int32_t *mixed_frames, gas_frame *bus_out, gas_frame *peaks, int32_t *status_out, int threads);
int orc_set_playback_disable_threshold_db(orc_world *w, int n, const int32_t *instances, const float *db);
int orc_voice_state_export(orc_world *w, int n, const int32_t *voices, gas_voice_state *out);
int orc_voice_state_import(orc_world *w, int n, const int32_t *voices, const gas_voice_state *in);
/* seconds spent inside the last orc_mix_block / orc_gain_compute (steady_clock), for bench.py */
double orc_last_mix_seconds(const orc_world *w);
double orc_last_gain_seconds(const orc_world *w);
int orc_max_threads(void);
size_t orc_sizeof(int32_t struct_id); /* same ids as gas_abi_sizeof */

/* The bus graph after the mix (upstream AudioServer::_mix_step, restated AS RECALLED): volume / mute / solo / send routing in place. */
void orc_bus_graph(int n_buses, int channels, int frames, const float *volume_db, const int32_t *mute, const int32_t *solo, const int32_t *send,
		gas_frame *bus);

/* The resampler in front of the path (upstream AudioStreamPlaybackResampled over plain PCM, restated AS RECALLED: not pinned by
 * reference code, see gas_oracle.c).  One object per playing voice; pcm must outlive it. */
typedef struct orc_resampler orc_resampler;
orc_resampler *orc_resampler_begin(const gas_frame *pcm, int n_frames, int loop, float sample_rate, int start_frame);
int orc_resampler_mix(orc_resampler *r, gas_frame *out, float rate_scale, float target_rate, int frames);
void orc_resampler_free(orc_resampler *r);

#ifdef __cplusplus
}
#endif
#endif
