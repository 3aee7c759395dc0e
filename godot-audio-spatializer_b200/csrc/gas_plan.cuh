// gas_plan.cuh — the per-block planner: turns the current parameters + persistent ramp state into the block's plan
// (classes, weight records, voice records) and advances the ramp state.  One implementation, two hosts: the
// stand-alone kernel k_plan (gas_prologue.cu) and the control warps of the step kernel (gas_mix_stream.cu), which plan
// block b + 1 on the SMs that stream block b.
//
// Work is laid out 2 lanes per voice (lane = side: L, R), each lane looping over the channel pairs:
//   voice part: what process_frames / mix_channel decide before their sample loop (reference
//       audio_spatializer_3d.cpp:499-523, :562-587, :537-551, :608): ramp end points, filter on/off,
//       clear-history, target coefficients — plus the AudioServer side of the instance's proxy playbacks
//       for this mix step (upstream AudioServer::_mix_step): previous volume looked up
//       by bus, buses that disappeared fade to 0.  The voice is classified by what its weights look like
//       and appended to its class list (CTA-level aggregation in shared memory, one global atomic per
//       class per CTA and pass).
//   instance part: prev <- cur of the bus details.  prev is double-buffered by block parity: block b
//       reads inst_prev[b & 1] and writes inst_prev[(b + 1) & 1], so no grid-wide barrier is needed.
//   The planner also zeroes the block's bus buffers / peaks; its last CTA writes the compact class tables of the block,
//   recycles idle class slots, clears the class counts of the next plan slot, advances the block counter and publishes
//   the plan (PlanHdr::seq).
//
// Include only from translation units compiled with -fmad=false (coefficient preparation is double arithmetic
// narrowed to float; the weight products must round like the reference's separate multiplies).
#pragma once

#include "gas_internal.h"
#include "gas_filter.cuh"

namespace gasplan {

constexpr int kTab = 128; // distinct classes a CTA can meet in one pass (>= voices per pass = planner threads / 2)
constexpr int kBigKey = 0x7fffffff;

struct PlanSmem {
	unsigned long long gkey[GAS_MAX_CLASSES]; // snapshot of the global slot table
	unsigned long long gaux[GAS_MAX_CLASSES];
	unsigned long long key[kTab]; // classes met in this pass
	unsigned long long aux[kTab];
	int cnt[kTab], cid[kTab], base[kTab];
	int block;  // index of the block being planned
	int ticket; // finish ticket of this CTA
};

// The threads of one CTA that plan together: the whole CTA of k_plan, or the control warps of the step kernel.
struct PlanGroup {
	int tid;      // 0 .. nthreads - 1
	int nthreads; // multiple of 32, <= 2 * kTab
	int cta, n_cta;
	int bar_id;   // named barrier of the group
	unsigned long long *tl; // experiments: globaltimer stamps of the group's first thread (nullptr normally)
};
static __device__ __forceinline__ void group_stamp(const PlanGroup &G, int slot) {
	if (G.tl) {
		G.tl[slot] = gas_globaltimer();
	}
}
static __device__ __forceinline__ void group_sync(const PlanGroup &G) { gas_bar_sync(G.bar_id, G.nthreads); }

struct PlanArgs {
	DevTables t;
	GlobalCfg g;
	BlockPlan plan;
	int inst_hwm;
	int n_voices;
	const gas_voice *voices;
	int src_rows;
	float4 *bus; // the block's bus buffers, zeroed here
	int bus_f4;
	float2 *peaks; // optional, n_voices entries, zeroed here
	int scaled_classes;
};

static __device__ __forceinline__ int resolve_bus(const GlobalCfg &g, int bus) { return (bus >= 0 && bus < g.num_buses) ? bus : 0; }
static __device__ __forceinline__ int ld_volatile(const int32_t *p) { return *(volatile const int32_t *)p; }
static __device__ __forceinline__ void st_release(int32_t *p, int v) { gas_st_release_gpu_s32(p, v); }
static __device__ __forceinline__ int ld_acquire(const int32_t *p) { return gas_ld_acquire_gpu_s32(p); }

// Sends of an instance with more than two buses on either side (custom parameters only: calculate_spatialization never
// produces more than two): resolved straight into the voice's InstSends record in global memory, this lane's side.  Same
// order as the register path below: every bus of the current details with the previous volume looked up by bus (absent
// => 0 => fade-in), then buses only present in the previous details towards 0; ascending by (bus, appearance).
static __device__ __noinline__ void resolve_sends_wide(const BusDetails *__restrict__ cur, const BusDetails *__restrict__ prev, const GlobalCfg &g, int x,
		bool poison, InstSends *__restrict__ ps) {
	constexpr int K = GAS_MAX_BUSES_PER_PLAYBACK;
	int cn = cur->n, pn = prev->n;
	cn = cn < 0 ? 0 : (cn > K ? K : cn);
	pn = pn < 0 ? 0 : (pn > K ? K : pn);
	int key[2 * K], srcp[2 * K]; // srcp: index into prev for the previous volume (-1 = none)
	int total = 0;
	for (int k = 0; k < cn; k++) {
		const int b = cur->bus[k];
		int jp = -1;
		for (int j = 0; j < pn; j++) {
			if (prev->bus[j] == b) {
				jp = j; // the last match wins, like a lookup that keeps scanning
			}
		}
		key[total] = resolve_bus(g, b) * 16 + k;
		srcp[total] = jp;
		total++;
	}
	for (int j = 0; j < pn; j++) {
		bool only = true;
		for (int k = 0; k < cn; k++) {
			if (cur->bus[k] == prev->bus[j]) {
				only = false;
			}
		}
		if (only) {
			key[total] = resolve_bus(g, prev->bus[j]) * 16 + K + j;
			srcp[total] = j;
			total++;
		}
	}
	uint32_t mask = 0;
	int last = -1;
	const float qnan = __int_as_float(0x7fc00000);
	for (int i = 0; i < total; i++) {
		int best = kBigKey, bi = 0;
		for (int m = 0; m < total; m++) {
			if (key[m] > last && key[m] < best) {
				best = key[m];
				bi = m;
			}
		}
		last = best;
		const int slot16 = best & 15;
		const bool from_cur = slot16 < K;
		if (x == 0) {
			ps->bus[i] = best >> 4;
		}
		for (int c = 0; c < GAS_MAX_CHANNELS_PER_BUS; c++) {
			const float vp = srcp[bi] >= 0 ? prev->vol[srcp[bi]][c][x] : 0.f;
			float vn = from_cur ? cur->vol[slot16][c][x] : 0.f;
			if (poison) {
				vn = qnan;
			}
			ps->vp[i][c][x] = vp;
			ps->vn[i][c][x] = vn;
		}
		mask |= 1u << (best >> 4);
	}
	if (x == 0) {
		ps->n = total;
		ps->mask = mask;
	}
}

// Coefficient preparation is long double-precision code that only voices of the voice-parallel path need: out of line, so
// that the common path (streamed voices) does not have to be fetched around it.
static __device__ __noinline__ void write_filter_target(VoiceRec *rec, float cutoff, float lin_att, float mix_rate) {
	float target[5];
	prepare_coefficients(GAS_FILTER_HIGHSHELF, cutoff, 1.0f, lin_att, 1, mix_rate, target); // audio_spatializer_3d.cpp:504-510
#pragma unroll
	for (int i = 0; i < 5; i++) {
		rec->target[i] = target[i];
	}
}
static __device__ __noinline__ void write_effect_coefficients(VoiceRec *rec, const gas_effect_chain *fx, int n_fx, int fx_binding, float lin_att,
		float mix_rate, int x) {
	for (int ei = x; ei < n_fx; ei += 2) { // effects x, x + 2 on this lane
		gas_effect ef = fx->effects[ei];
		if (fx_binding == ei) {
			ef.gain = lin_att; // example _process_effects (gd_spatializer_instance.gd:125-127)
		}
		const int stages = ef.stages < 1 ? 1 : (ef.stages > GAS_MAX_FILTER_STAGES ? GAS_MAX_FILTER_STAGES : ef.stages);
		float coef[5];
		prepare_coefficients(ef.mode, ef.cutoff_hz, ef.resonance, ef.gain, stages, mix_rate, coef);
		rec->fx_stages[ei] = stages;
#pragma unroll
		for (int i = 0; i < 5; i++) {
			rec->fx_coef[ei][i] = coef[i];
		}
	}
}

// One pass of the voice part: the group's threads take nthreads / 2 consecutive voices starting at j0.
static __device__ __forceinline__ void plan_voices_pass(const PlanGroup &G, PlanSmem &S, const PlanArgs &a, int b, int j0, int wait_total, int own_q,
		const gas_voice *pre) {
	const DevTables &t = a.t;
	const GlobalCfg &g = a.g;
	const BlockPlan &plan = a.plan;
	const int C = g.channels;
	const int maxv = g.max_voices;
	const int slot_p = b & (GAS_PLAN_DEPTH - 1);
	const int parity = b & 1;
	int32_t *cnt_now = plan.cls_count + slot_p * GAS_MAX_CLASSES;
	const BusDetails *prev_rd = t.inst_prev + (size_t)parity * t.max_instances;
	const int x = G.tid & 1;                  // side
	const int j = j0 + (G.tid >> 1);          // call-order index of this lane pair's voice
	const unsigned lane = threadIdx.x & 31u;
	const unsigned gm = 3u << (lane & 30u);   // the two lanes of the voice

	// ---- reset the pass table -----------------------------------------------------------------------------------
	for (int i = G.tid; i < kTab; i += G.nthreads) {
		S.key[i] = 0ULL;
		S.aux[i] = CLS_AUX_NONE;
		S.cnt[i] = 0;
		S.cid[i] = -1;
		S.base[i] = 0;
	}
	for (int i = G.tid; i < GAS_MAX_CLASSES; i += G.nthreads) {
		S.gkey[i] = __ldcg(plan.cls_key + i);
		S.gaux[i] = __ldcg(plan.cls_aux + i);
	}

	// ---- level 0 / level 1 loads -----------------------------------------------------------------------------------
	gas_voice v{};
	v.voice = -1;
	if (pre) {
		v = *pre; // fetched by the caller before its gain tasks
	} else if (j < a.n_voices) {
		v = a.voices[j];
	}
	const bool valid = v.voice >= 0 && v.voice < maxv && v.instance >= 0 && v.instance < g.max_instances;
	const int q = valid ? v.instance : 0;
	const int vslot = valid ? v.voice : 0;
	if (wait_total > 0 && valid && q != own_q) { // (own_q: this lane pair computed that instance's gains itself)
		// fine-grained dependency instead of a grid-wide barrier between the gain tasks and the plan: normally the flag is already
		// set (the same lanes computed this instance's gains a moment ago)
		while (ld_acquire(&t.inst_seq[q]) != b + 1 && ld_acquire(&t.blk[BLK_GAIN_DONE]) < wait_total) {
			__nanosleep(40);
		}
	}
	const int v_active = t.inst_active[q];
	const int imode = t.inst_mode[q];
	const gas_params *prm = &t.inst_params[q];
	const float lin_att = prm->linear_attenuation;
	const float cutoff = prm->attenuation_filter_cutoff_hz;
	const BusDetails *curp = &t.inst_cur[q];
	const BusDetails *prevp = &prev_rd[q];
	const int cur_n = curp->n, prev_n = prevp->n;
	const int cur_bus[2] = { curp->bus[0], curp->bus[1] };
	const int prev_bus[2] = { prevp->bus[0], prevp->bus[1] };
	float mixv[4], cur_vol[2][4], prev_vol[2][4], vp_old[4];
	float *vprev = t.vs_prev + (size_t)vslot * 8;
#pragma unroll
	for (int c = 0; c < 4; c++) {
		mixv[c] = prm->mix_volumes[c][x];
		cur_vol[0][c] = curp->vol[0][c][x];
		cur_vol[1][c] = curp->vol[1][c][x];
		prev_vol[0][c] = prevp->vol[0][c][x];
		prev_vol[1][c] = prevp->vol[1][c][x];
		vp_old[c] = vprev[c * 2 + x];
	}
	group_sync(G); // the pass table is reset
	group_stamp(G, 28);

	// ---- voice part ----------------------------------------------------------------------------------------
	int path = PATH_NONE, mode = MODE_A, n_send = 0, n_group = 0;
	uint32_t cflags = 0, mask = 0, quad = 0, rflags = 0;
	const bool live = valid && v_active != 0;
	if (v.src_row >= a.src_rows) {
		v.src_row = -1;
	}
	int s_bus[4] = { 0, 0, 0, 0 };
	float s_vp[4][4], s_vn[4][4];
#pragma unroll
	for (int i = 0; i < 4; i++) {
#pragma unroll
		for (int c = 0; c < 4; c++) {
			s_vp[i][c] = 0.f;
			s_vn[i][c] = 0.f;
		}
	}
	float m_prev[4] = { 1.f, 1.f, 1.f, 1.f }, m_new[4] = { 1.f, 1.f, 1.f, 1.f }; // this side's mix_channel ramp per pair
	int n_fx = 0, fx_binding = -1;
	bool shared = false, scaled = false, wide = false;
	float sc1 = 0.f, sc2 = 0.f;
	unsigned long long aux = CLS_AUX_NONE; // second word of the class identity

	if (live) {
		mode = imode & 0xff;
		fx_binding = (imode >> 8) - 1;
		wide = cur_n > 2 || prev_n > 2;
		if (!wide) {
			// sends: every bus of the current details with the previous volume looked up by bus (absent => 0 => fade-in), then
			// buses only present in the previous details once more towards 0 (fade-out); ascending by (bus, appearance) so
			// that a class is identified by its bus mask
			const int cn = cur_n < 0 ? 0 : cur_n, pn = prev_n < 0 ? 0 : prev_n;
			int ckey[2], pkey[2];
			float cvp[2][4];
			int total = cn;
#pragma unroll
			for (int k = 0; k < 2; k++) {
				ckey[k] = k < cn ? resolve_bus(g, cur_bus[k]) * 16 + k : kBigKey;
#pragma unroll
				for (int c = 0; c < 4; c++) {
					cvp[k][c] = 0.f;
				}
#pragma unroll
				for (int jj = 0; jj < 2; jj++) {
					if (k < cn && jj < pn && prev_bus[jj] == cur_bus[k]) {
#pragma unroll
						for (int c = 0; c < 4; c++) {
							cvp[k][c] = prev_vol[jj][c]; // the last match wins, like a lookup that keeps scanning
						}
					}
				}
			}
#pragma unroll
			for (int jj = 0; jj < 2; jj++) {
				bool only = jj < pn;
#pragma unroll
				for (int k = 0; k < 2; k++) {
					if (k < cn && cur_bus[k] == prev_bus[jj]) {
						only = false;
					}
				}
				pkey[jj] = only ? resolve_bus(g, prev_bus[jj]) * 16 + 2 + jj : kBigKey;
				total += only ? 1 : 0;
			}
			n_send = total;
			int last = -1;
#pragma unroll
			for (int i = 0; i < 4; i++) {
				if (i < total) {
					int best = kBigKey;
					float bp[4] = { 0.f, 0.f, 0.f, 0.f }, bn[4] = { 0.f, 0.f, 0.f, 0.f };
#pragma unroll
					for (int k = 0; k < 2; k++) {
						if (ckey[k] > last && ckey[k] < best) {
							best = ckey[k];
#pragma unroll
							for (int c = 0; c < 4; c++) {
								bp[c] = cvp[k][c];
								bn[c] = cur_vol[k][c];
							}
						}
					}
#pragma unroll
					for (int jj = 0; jj < 2; jj++) {
						if (pkey[jj] > last && pkey[jj] < best) {
							best = pkey[jj];
#pragma unroll
							for (int c = 0; c < 4; c++) {
								bp[c] = prev_vol[jj][c];
								bn[c] = 0.f;
							}
						}
					}
					s_bus[i] = best >> 4;
#pragma unroll
					for (int c = 0; c < 4; c++) {
						s_vp[i][c] = bp[c];
						s_vn[i][c] = bn[c];
					}
					mask |= 1u << (best >> 4);
					last = best;
				}
			}
		}
		const bool filt = mode != MODE_E && (double)lin_att >= 0.001; // audio_spatializer_3d.cpp:503, :568
		const bool want_peak = (v.flags & GAS_VOICE_WANT_PEAK) != 0;
		rflags = v.flags & 0xffu;
		if (mode == MODE_B) {
			uint32_t zero_bits = 0;
#pragma unroll
			for (int c = 0; c < 4; c++) {
				if (c < C) {
					m_prev[c] = vp_old[c];           // :564
					m_new[c] = mixv[c];              // :565
					vprev[c * 2 + x] = m_new[c];     // :608
					if (m_prev[c] == 0.f) {
						zero_bits |= 1u << c;
					}
				}
			}
			// is_just_started per pair: previous (L, R) exactly (0, 0), :583
			zero_bits &= __shfl_xor_sync(gm, zero_bits, 1);
			rflags |= zero_bits << 8;
		} else if (mode == MODE_A) {
			const float other0 = __shfl_xor_sync(gm, vp_old[0], 1);
			if (vp_old[0] == 0.f && other0 == 0.f) {
				rflags |= 1u << 8; // :518
			}
			// :537-551 — the (L,R) pair holding the first maximum in scan order c0.L, c0.R, c1.L, ...
			// (a running maximum that starts at 0 and only moves on a strict >: NaN volumes are passed over)
			float bv = 0.f;
			int bi = 99;
#pragma unroll
			for (int c = 0; c < 4; c++) {
				if (mixv[c] > bv) {
					bv = mixv[c];
					bi = c * 2 + x;
				}
			}
			{
				const float ov = __shfl_xor_sync(gm, bv, 1);
				const int oi = __shfl_xor_sync(gm, bi, 1);
				if (ov > bv || (ov == bv && oi < bi)) {
					bv = ov;
					bi = oi;
				}
			}
			const int max_index = bv > 0.f ? (bi >> 1) : 0;
			float keep = mixv[0];
#pragma unroll
			for (int c = 1; c < 4; c++) {
				keep = c == max_index ? mixv[c] : keep;
			}
			vprev[x] = keep;
		}
		if (filt) {
			cflags |= CLS_FILT;
		}
		if (mode == MODE_E) {
			n_fx = t.inst_fx[q].n_effects;
			n_fx = n_fx < 0 ? 0 : (n_fx > GAS_MAX_EFFECTS ? GAS_MAX_EFFECTS : n_fx);
		}
		const bool has_dsp = filt || (mode == MODE_E && n_fx > 0);
		bool poison = false;
		if (mode == MODE_B) {
			// Every Mode-B proxy is a playback of its own: AudioServer runs _mix_step_for_channel for every pair of every
			// bus of its map, with volume 0 for the pairs the map masks out (reference audio_spatializer.cpp:298-312).
			// 0 * x only matters when the proxy's buffer is not finite, which the module produces itself (NaN pan gains,
			// SURVEY Q1): the NaN then reaches every pair of every bus the instance sends to, same side.  A NaN ramp end
			// point of any pair therefore poisons this side's send volumes of the voice.
#pragma unroll
			for (int c = 0; c < 4; c++) {
				if (c < C && (m_prev[c] != m_prev[c] || m_new[c] != m_new[c])) {
					poison = true;
				}
			}
			if (poison) {
				const float qnan = __int_as_float(0x7fc00000);
#pragma unroll
				for (int k = 0; k < 4; k++) {
					if (k < n_send) {
#pragma unroll
						for (int c = 0; c < 4; c++) {
							s_vn[k][c] = qnan;
						}
					}
				}
			}
		}
		if (wide) {
			// more than two buses on one side: the voice-parallel kernel's generic class mixes any send layout
			resolve_sends_wide(curp, prevp, g, x, poison, &plan.sends[(size_t)slot_p * maxv + j]);
			path = PATH_VOICE;
			cflags &= CLS_FILT;
		} else {
			// weight polynomial per (send, pair, side): w(t) = A + B t + Cq t^2 with t = i/F, from
			// (vn*t + (1-t)*vp) of the AudioServer ramp times (m_new*t + (1-t)*m_prev) of mix_channel.
			bool streamed = false;
			// A silent source (the reference's zero-filled playback_buffer, audio_spatializer.cpp:405-408) contributes exactly nothing —
			// unless one of its volumes is not finite (NaN pan gains): 0 * NaN is NaN and reaches the buses like any other
			// sample.  Such a voice goes through the voice-parallel kernel, which multiplies the zeros out.
			int silent_poison = 0;
			if (v.src_row < 0) {
#pragma unroll
				for (int k = 0; k < 4; k++) {
#pragma unroll
					for (int c = 0; c < 4; c++) {
						if (k < n_send && c < C) {
							const float z = (s_vn[k][c] - s_vn[k][c]) + (s_vp[k][c] - s_vp[k][c]) + (m_new[c] - m_new[c]) + (m_prev[c] - m_prev[c]);
							if (z != 0.f) { // x - x is 0 for every finite x
								silent_poison = 1;
							}
						}
					}
				}
				silent_poison |= __shfl_xor_sync(gm, silent_poison, 1);
			}
			if (!has_dsp && !want_peak && n_send >= 1 && !silent_poison) {
				int q_any = 0, differs = 0;
#pragma unroll
				for (int k = 0; k < 4; k++) {
#pragma unroll
					for (int c = 0; c < 4; c++) {
						if (k < n_send && c < C) {
							const float dn = s_vn[k][c] - s_vp[k][c];
							const float dm = m_new[c] - m_prev[c];
							if (dn * dm != 0.f) {
								q_any = 1;
							}
							if (s_vn[k][c] != s_vn[0][c] || s_vp[k][c] != s_vp[0][c]) {
								differs = 1;
							}
						}
					}
				}
				q_any |= __shfl_xor_sync(gm, q_any, 1);
				differs |= __shfl_xor_sync(gm, differs, 1);
				shared = n_send >= 2 && !differs;
				// Scaled sends: every further send is send 0 times ONE scalar (same for both ramp end points, all pairs, both
				// sides) — what a reverb send with uniformity 0 is (reverb_vol = direct * area_send, reference
				// audio_spatializer_3d.cpp:192-196, so bus_vol / mix_vol is area_send to the last bit or two on every pair).
				// Such a voice needs one row group: the flush adds the sums to bus 0 as they are and to the other buses times
				// the class's scales.  The scalar of a send is taken from the first element (pair-major, L before R) with a
				// non-zero base volume; an element accepts it if it reproduces its own volumes within 3e-7 relative (2-3 ulp:
				// far inside the 1e-5 tolerance).
				if (n_send >= 2 && n_send <= 3 && !shared && a.scaled_classes) {
					int first = 99;
					float r1 = 0.f, r2 = 0.f;
#pragma unroll
					for (int c = 3; c >= 0; c--) {
						if (c < C && (s_vn[0][c] != 0.f || s_vp[0][c] != 0.f)) {
							first = c * 2 + x;
							const float bn = s_vn[0][c], bp = s_vp[0][c];
							const bool use_n = bn != 0.f;
							r1 = use_n ? s_vn[1][c] / bn : s_vp[1][c] / bp;
							r2 = n_send > 2 ? (use_n ? s_vn[2][c] / bn : s_vp[2][c] / bp) : 0.f;
						}
					}
					const int ofirst = __shfl_xor_sync(gm, first, 1);
					const float or1 = __shfl_xor_sync(gm, r1, 1), or2 = __shfl_xor_sync(gm, r2, 1);
					int ok = (first < 99 || ofirst < 99) ? 1 : 0;
					if (ofirst < first) {
						r1 = or1;
						r2 = or2;
					}
					sc1 = r1;
					sc2 = r2;
					if (ok) {
						const float tol = 3e-7f;
#pragma unroll
						for (int c = 0; c < 4; c++) {
							if (c < C) {
								const float bn = s_vn[0][c], bp = s_vp[0][c];
								bool o = fabsf(s_vn[1][c] - sc1 * bn) <= tol * fabsf(s_vn[1][c]) && fabsf(s_vp[1][c] - sc1 * bp) <= tol * fabsf(s_vp[1][c]);
								if (n_send > 2) {
									o = o && fabsf(s_vn[2][c] - sc2 * bn) <= tol * fabsf(s_vn[2][c]) && fabsf(s_vp[2][c] - sc2 * bp) <= tol * fabsf(s_vp[2][c]);
								}
								ok = (ok && o) ? 1 : 0;
							}
						}
						ok = (ok && (sc1 == sc1) && (sc2 == sc2) && fabsf(sc1) < 3.0e38f && fabsf(sc2) < 3.0e38f) ? 1 : 0;
					}
					ok &= __shfl_xor_sync(gm, ok, 1);
					scaled = ok != 0;
				}
				n_group = (shared || scaled) ? 1 : n_send;
				if (scaled) {
					// only send 0's rows are stored: its own ramp product decides the degree
					int q0 = 0;
#pragma unroll
					for (int c = 0; c < 4; c++) {
						if (c < C && (s_vn[0][c] - s_vp[0][c]) * (m_new[c] - m_prev[c]) != 0.f) {
							q0 = 1;
						}
					}
					q_any = q0 | __shfl_xor_sync(gm, q0, 1);
				}
				quad = q_any ? 1u : 0u;
				if (n_group * (2 + (int)quad) <= GAS_K2_MAX_ROWS) {
					streamed = true;
					path = PATH_STREAM;
					if (shared) {
						cflags |= CLS_SHARED;
					}
					if (scaled) {
						cflags |= CLS_SCALED;
						aux = (unsigned long long)__float_as_uint(sc1) | ((unsigned long long)__float_as_uint(sc2) << 32);
					}
					if (v.src_row < 0) {
						path = PATH_NONE; // silent source, no DSP state to advance: contributes exactly nothing
					}
				}
			}
			if (!streamed) {
				// needs the voice-parallel kernel unless there is neither DSP state to advance, nor a peak
				// to report, nor a bus to reach
				if (has_dsp || want_peak || n_send > 0) {
					path = PATH_VOICE;
					cflags &= CLS_FILT;
					n_group = n_send;
					quad = 0;
				}
			}
		}
	}

	if (G.tl) {
		G.tl[8] = (unsigned long long)(__float_as_uint(mixv[0] + cur_vol[0][0] + prev_vol[0][0] + vp_old[0]) & 1u); // (forces the loads to have landed)
	}
	group_stamp(G, 29);
	// ---- class lookup: once per class per CTA in shared memory, then one global atomic per class ------------
	// A class is (key, aux).  The key of a scaled class carries a 20-bit hash of its aux word in its spare bits, so that
	// the compare-and-swap that claims a slot sees (almost always) the whole identity; the aux words are compared once
	// they are published (after the barrier in the CTA table; after a short wait in the global table).
	unsigned long long key = (path != PATH_NONE && !wide) ? cls_key(path, mode, cflags, n_send, mask, quad) : 0ULL;
	if (key != 0ULL && aux != CLS_AUX_NONE) {
		key |= ((aux * 0x9E3779B97F4A7C15ULL) >> 44) << 44;
	}
	int slot = wide && path != PATH_NONE ? -2 : -1, lpos = 0;
	if (key != 0ULL && x == 0) {
		for (int i = 0; i < kTab; i++) {
			unsigned long long k = *(volatile unsigned long long *)&S.key[i];
			if (k == 0ULL) {
				k = atomicCAS(&S.key[i], 0ULL, key);
				if (k == 0ULL) {
					k = key;
					S.aux[i] = aux;
				}
			}
			if (k == key) {
				slot = i;
				break;
			}
		}
	}
	group_sync(G);
	if (slot >= 0 && S.aux[slot] != aux) {
		// two scaled classes whose aux words hash alike met in one pass (one in 2^20 pairs): this voice takes the generic
		// class of the voice-parallel kernel instead, which mixes any voice
		slot = -2;
	}
	if (slot >= 0) {
		lpos = atomicAdd(&S.cnt[slot], 1);
	}
	int gpos = 0; // position in the generic class when slot == -2
	const int generic_cid = GAS_CLS_DYNAMIC + mode * 2 + ((cflags & CLS_FILT) ? 1 : 0);
	if (slot == -2 && x == 0) {
		gpos = atomicAdd(&cnt_now[generic_cid], 1);
	}
	group_sync(G);
	group_stamp(G, 30);
	if (G.tid < kTab && S.key[G.tid] != 0ULL && S.cnt[G.tid] > 0) {
		// Slots are stable across blocks: in the steady state the class is already in the snapshot and the only
		// global operation is the add that reserves this CTA's range of the class list.
		const unsigned long long k = S.key[G.tid];
		const unsigned long long ka = S.aux[G.tid];
		int cid = -1;
		for (int i = 0; i < GAS_CLS_DYNAMIC; i++) {
			if (S.gkey[i] == k && S.gaux[i] == ka) {
				cid = i;
				break;
			}
		}
		if (cid < 0) { // first appearance of the class: claim a free slot, or find the one another CTA just claimed
			for (int i = 0; i < GAS_CLS_DYNAMIC && cid < 0; i++) {
				unsigned long long o = S.gkey[i];
				if (o != 0ULL && o != k) {
					continue;
				}
				o = atomicCAS(&plan.cls_key[i], 0ULL, k);
				if (o == 0ULL) {
					*(volatile unsigned long long *)&plan.cls_aux[i] = ka; // the claimer publishes the aux word
					__threadfence();
					cid = i;
				} else if (o == k) {
					// another CTA owns the slot under the same key: the same class if its aux word matches.  Its claimer
					// publishes the aux word right after its compare-and-swap; give it a moment, then look elsewhere (a class
					// may end up in two slots during the block it first appears in: the kernels treat them as two classes).
					unsigned long long aw = *(volatile unsigned long long *)&plan.cls_aux[i];
					if (ka != CLS_AUX_NONE) {
						for (int spin = 0; spin < 256 && aw == CLS_AUX_NONE; spin++) {
							aw = *(volatile unsigned long long *)&plan.cls_aux[i];
						}
					}
					if (aw == ka) {
						cid = i;
					}
				}
			}
			if (cid < 0) {
				// more distinct classes than slots: the voices go to the generic class of their mode
				*plan.overflow = 1;
				const int kmode = (int)((k >> 2) & 3u);
				const int kfilt = ((k >> 4) & CLS_FILT) ? 1 : 0;
				cid = GAS_CLS_DYNAMIC + kmode * 2 + kfilt;
			}
		}
		S.base[G.tid] = atomicAdd(&cnt_now[cid], S.cnt[G.tid]);
		S.cid[G.tid] = cid;
	}
	group_sync(G);
	group_stamp(G, 31);
	slot = __shfl_sync(gm, slot, lane & 30u);
	lpos = __shfl_sync(gm, lpos, lane & 30u);
	gpos = __shfl_sync(gm, gpos, lane & 30u);
	const int cid = slot >= 0 ? S.cid[slot] : (slot == -2 ? generic_cid : -1);
	if (cid >= 0) {
		const int pos = slot >= 0 ? S.base[slot] + lpos : gpos;
		if (x == 0) {
			plan_list(plan, slot_p, cid, maxv)[pos] = make_int2(j, v.src_row);
		}
		if (cid >= GAS_CLS_DYNAMIC) {
			path = PATH_VOICE; // generic class: the voice-parallel kernel mixes any voice
		}
		if (path == PATH_STREAM) {
			// weight record: [group][pair]{A_L, A_R, B_L, B_R}, then [group][pair]{C_L, C_R} for classes with a t^2 row
			const int nf = cls_row_floats(n_group, (int)quad, C);
			float *dst = plan_rows(plan, b, cid, maxv) + (size_t)pos * nf;
			float *dstq = dst + n_group * C * 4;
#pragma unroll
			for (int k = 0; k < GAS_K2_MAX_ROWS / 2; k++) {
				if (k < n_group) {
#pragma unroll
					for (int c = 0; c < 4; c++) {
						if (c < C) {
							const float mp = m_prev[c], dm = m_new[c] - m_prev[c];
							const float np = s_vp[k][c], dn = s_vn[k][c] - np;
							dst[(k * C + c) * 4 + x] = np * mp;
							dst[(k * C + c) * 4 + 2 + x] = np * dm + dn * mp;
							if (quad) {
								dstq[(k * C + c) * 2 + x] = dn * dm;
							}
						}
					}
				}
			}
		} else {
			VoiceRec *rec = &plan.rec[(size_t)slot_p * maxv + j];
			InstSends *ps = &plan.sends[(size_t)slot_p * maxv + j];
#pragma unroll
			for (int c = 0; c < 4; c++) {
				rec->m_prev[c][x] = m_prev[c];
				rec->m_new[c][x] = m_new[c];
			}
			if (x == 0) {
				rec->voice = v.voice;
				rec->instance = v.instance;
				rec->src_row = v.src_row;
				rec->flags = rflags;
				rec->n_fx = n_fx;
				if (cflags & CLS_FILT) {
					write_filter_target(rec, cutoff, lin_att, g.mix_rate);
				}
				if (!wide) {
					ps->n = n_send;
					ps->mask = mask;
				}
			}
			if (n_fx > 0) {
				write_effect_coefficients(rec, &t.inst_fx[q], n_fx, fx_binding, lin_att, g.mix_rate, x);
			}
			if (!wide) {
#pragma unroll
				for (int k = 0; k < 4; k++) {
					if (k < n_send) {
						if (x == 0) {
							ps->bus[k] = s_bus[k];
						}
#pragma unroll
						for (int c = 0; c < 4; c++) {
							ps->vp[k][c][x] = s_vp[k][c];
							ps->vn[k][c][x] = s_vn[k][c];
						}
					}
				}
			}
		}
	}
}

// The whole plan of block b = blk[BLK_P] by n_cta cooperating groups.  wait_total > 0: the gains of this block are being
// computed by the control warps of this launch (step kernel): a voice waits until its instance's parameters are in place
// (DevTables::inst_seq) or until all wait_total warps have reported (BLK_GAIN_DONE).
static __device__ __forceinline__ int plan_first_voice(const PlanGroup &G) { return G.cta * (G.nthreads >> 1) + (G.tid >> 1); }
static __device__ __forceinline__ void plan_block(const PlanGroup &G, PlanSmem &S, const PlanArgs &a, int wait_total, int own_q = -1,
		const gas_voice *pre = nullptr) {
	const DevTables &t = a.t;
	const BlockPlan &plan = a.plan;
	if (G.tid == 0) {
		// every CTA reads the block index before it takes its finish ticket; the last ticket holder advances it
		S.block = ld_volatile(&t.blk[BLK_P]);
	}
	group_sync(G);
	const int b = S.block;
	const int slot_p = b & (GAS_PLAN_DEPTH - 1);
	const int parity = b & 1;
	const int gtid = G.cta * G.nthreads + G.tid;
	const int gthreads = G.n_cta * G.nthreads;

	// ---- housekeeping stores --------------------------------------------------------------------------------------
	for (int i = gtid; i < a.bus_f4; i += gthreads) {
		a.bus[i] = make_float4(0.f, 0.f, 0.f, 0.f);
	}
	if (a.peaks) {
		for (int i = gtid; i < a.n_voices; i += gthreads) {
			a.peaks[i] = make_float2(0.f, 0.f);
		}
	}

	group_stamp(G, 19);
	// ---- voice part, nthreads / 2 voices per CTA and pass -------------------------------------------------------------
	const int vpc = G.nthreads >> 1;
	const int per_pass = G.n_cta * vpc;
	const int passes = (a.n_voices + per_pass - 1) / per_pass;
	for (int p = 0; p < passes; p++) {
		plan_voices_pass(G, S, a, b, (p * G.n_cta + G.cta) * vpc, wait_total, own_q, p == 0 ? pre : nullptr);
		group_sync(G); // the pass table is free again
	}

	group_stamp(G, 20);
	// ---- instance part: prev <- cur (2 lanes per instance), after every instance's gains are in place -------------------------
	{
		const BusDetails *curs = t.inst_cur;
		BusDetails *prev_wr = t.inst_prev + (size_t)(parity ^ 1) * t.max_instances;
		const int x = G.tid & 1;
		for (int q = gtid >> 1; q < a.inst_hwm; q += gthreads >> 1) {
			if (wait_total > 0 && q != own_q) {
				// this block's gains of instance q are in place (its flag), or no gain task of the launch is outstanding at all (an
				// instance without an emitter in this block).  Per instance instead of one wait for every warp of the launch: when
				// instance, emitter and voice indices coincide (q == own_q, the common layout) a CTA waits for no other CTA at all.
				while (ld_acquire(&t.inst_seq[q]) != b + 1 && ld_acquire(&t.blk[BLK_GAIN_DONE]) < wait_total) {
					__nanosleep(40);
				}
			}
			// one round of loads: the first two buses (calculate_spatialization never produces more), the rest only if present
			const BusDetails *cs = &curs[q];
			BusDetails *pw = &prev_wr[q];
			const int act = t.inst_active[q];
			int cn = cs->n;
			const int b0 = cs->bus[0], b1 = cs->bus[1];
			float v0[4], v1[4];
#pragma unroll
			for (int c = 0; c < 4; c++) {
				v0[c] = cs->vol[0][c][x];
				v1[c] = cs->vol[1][c][x];
			}
			if (!act) {
				continue;
			}
			cn = cn < 0 ? 0 : (cn > GAS_MAX_BUSES_PER_PLAYBACK ? GAS_MAX_BUSES_PER_PLAYBACK : cn);
			if (x == 0) {
				pw->n = cn;
				if (cn > 0) {
					pw->bus[0] = b0;
				}
				if (cn > 1) {
					pw->bus[1] = b1;
				}
			}
#pragma unroll
			for (int c = 0; c < 4; c++) {
				if (cn > 0) {
					pw->vol[0][c][x] = v0[c];
				}
				if (cn > 1) {
					pw->vol[1][c][x] = v1[c];
				}
			}
			for (int k = 2; k < cn; k++) {
				if (x == 0) {
					pw->bus[k] = cs->bus[k];
				}
#pragma unroll
				for (int c = 0; c < 4; c++) {
					pw->vol[k][c][x] = cs->vol[k][c][x];
				}
			}
		}
	}
	group_stamp(G, 19);
	// ---- finish: every warp reports (release: its lanes' stores first); the last one publishes the plan, then tidies up -------------
	__syncwarp();
	int last = 0;
	if ((G.tid & 31) == 0) {
		const int old = gas_atom_add_acq_rel_gpu_s32(&t.blk[BLK_P_TICKET], 1);
		last = old == G.n_cta * (G.nthreads >> 5) - 1;
	}
	last = __shfl_sync(0xffffffffu, last, 0);
	group_stamp(G, 21);
	if (!last) {
		return;
	}
	const int lane = G.tid & 31;
	PlanHdr *hdr = &plan.hdr[slot_p];
	// Published: the class slots and this block's counts, lists and records are final.  The step kernel reads the slot table and
	// the counts directly; what follows (compact table for the voice-parallel kernel, slot recycling, clearing the next counts,
	// the block counter) is only needed by launches that start after this one has ended.
	if (lane == 0) {
		st_release(&hdr->seq, b + 1);
	}
	const int32_t *cnt_now = plan.cls_count + slot_p * GAS_MAX_CLASSES;
	int32_t *cnt_next = plan.cls_count + ((b + 1) & (GAS_PLAN_DEPTH - 1)) * GAS_MAX_CLASSES;
	int base_s = 0, base_v = 0;
	constexpr int R = GAS_MAX_CLASSES / 32;
	unsigned long long keys[R], auxs[R];
	int counts[R], idles[R];
#pragma unroll
	for (int r = 0; r < R; r++) { // one round of loads
		const int i = r * 32 + lane;
		keys[r] = __ldcg(plan.cls_key + i);
		auxs[r] = __ldcg(plan.cls_aux + i);
		counts[r] = __ldcg(cnt_now + i);
		idles[r] = __ldcg(plan.cls_idle + i);
	}
#pragma unroll
	for (int r = 0; r < R; r++) {
		const int i = r * 32 + lane;
		const unsigned long long key = keys[r], auxw = auxs[r];
		const int count = counts[r], idle = idles[r];
		const bool used = key != 0ULL && count > 0;
		const bool on_s = used && (int)(key & 3u) == PATH_STREAM;
		const bool on_v = used && (int)(key & 3u) == PATH_VOICE;
		const unsigned ms = __ballot_sync(0xffffffffu, on_s);
		const unsigned mv = __ballot_sync(0xffffffffu, on_v);
		if (used) {
			ClassInfo ci = cls_decode(key, count);
			ci.slot = i;
			if (ci.flags & CLS_SCALED) {
				ci.scale[0] = __uint_as_float((unsigned)(auxw & 0xffffffffu));
				ci.scale[1] = __uint_as_float((unsigned)(auxw >> 32));
			}
			if (on_s) {
				hdr->cls[base_s + __popc(ms & ((1u << lane) - 1u))] = ci;
			} else if (on_v) {
				hdr->vcls[base_v + __popc(mv & ((1u << lane) - 1u))] = ci;
			}
		}
		base_s += __popc(ms);
		base_v += __popc(mv);
		// Slot recycling: a slot whose class stayed empty for GAS_CLS_IDLE_BLOCKS blocks is handed back.  Nothing reads a
		// slot with a zero count, and the next planner starts after this one has published.
		if (i < GAS_CLS_DYNAMIC && key != 0ULL) {
			const int age = count > 0 ? 0 : idle + 1;
			if (age >= GAS_CLS_IDLE_BLOCKS) {
				plan.cls_aux[i] = CLS_AUX_NONE;
				plan.cls_key[i] = 0ULL;
				plan.cls_idle[i] = 0;
			} else if (age != idle) {
				plan.cls_idle[i] = age;
			}
		}
		cnt_next[i] = 0;
	}
	if (lane == 0) {
		hdr->n_cls = base_s;
		hdr->n_vcls = base_v;
		t.blk[BLK_P_TICKET] = 0;
		t.blk[BLK_GAIN_DONE] = 0;
		*(volatile int32_t *)&t.blk[BLK_P] = b + 1;
	}
	group_stamp(G, 22);
}

} // namespace gasplan
