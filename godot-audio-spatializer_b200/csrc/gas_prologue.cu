// gas_prologue.cu — per-block prologue (one kernel): turns the current parameters + persistent ramp state
// into the block's plan (classes, weight rows, voice records) and advances the ramp state.
//
// Work is laid out 8 lanes per voice (lane = channel pair * 2 + side), 16 voices per CTA, so that every
// table access of a voice is one 32-byte segment:
//   voice part: what process_frames / mix_channel decide before their sample loop (reference
//       audio_spatializer_3d.cpp:499-523, :562-587, :537-551, :608): ramp end points, filter on/off,
//       clear-history, target coefficients — plus the AudioServer side of the instance's proxy playbacks
//       for this mix step (upstream AudioServer::_mix_step): previous volume looked up
//       by bus, buses that disappeared fade to 0.  The voice is classified by what its weights look like
//       and appended to its class list (CTA-level aggregation in shared memory, one global atomic per
//       class per CTA).
//   instance part: prev <- cur of the bus details.  prev is double-buffered by block parity: this block
//       reads inst_prev[p] and writes inst_prev[1-p], so no grid-wide barrier is needed.
//   The kernel also zeroes the bus buffers / peaks, clears the class table of the NEXT block, and its last
//   CTA advances the block counter.
//
// Compiled with -fmad=false (coefficient preparation is double arithmetic narrowed to float).
#include "gas_internal.h"
#include "gas_filter.cuh"

#include <stdlib.h>

namespace {

constexpr int kLanes = 8;          // lanes per voice
// 128-thread CTAs at ~120 registers: measured best inside the step on B200 (64 voices per CTA capped at
// 64 registers spilled and cost +1.5 us; 8 voices per CTA pays +1 us for the extra class-table atomics)
constexpr int kVoicesPerCta = 16;
constexpr int kCtaThreads = kLanes * kVoicesPerCta;
constexpr int kBigKey = 0x7fffffff;

__device__ __forceinline__ int resolve_bus(const GlobalCfg &g, int bus) {
	return (bus >= 0 && bus < g.num_buses) ? bus : 0;
}

// The bus details of one instance as this lane sees them: counts and bus ids (same in the 8 lanes of a
// voice) plus this lane's (pair, side) element of every volume.  Loaded unconditionally (all 6 slots) so
// that every load of the kernel is in flight at once; entries beyond n are ignored.
struct LaneDetails {
	int n;
	int bus[GAS_MAX_BUSES_PER_PLAYBACK];
	float vol[GAS_MAX_BUSES_PER_PLAYBACK];
};

__device__ __forceinline__ void details_load_lane(LaneDetails &d, const BusDetails *__restrict__ src, int c, int x) {
	d.n = src->n;
#pragma unroll
	for (int k = 0; k < GAS_MAX_BUSES_PER_PLAYBACK; k++) {
		d.bus[k] = src->bus[k];
		d.vol[k] = src->vol[k][c][x];
	}
}

// This lane's element (pair c, side x) of the sends of one instance for this block: every bus of the
// current details with the previous volume looked up by bus (absent => 0 => fade-in), then buses only
// present in the previous details once more towards 0 (fade-out); ascending by (bus, appearance) so that
// a class is identified by its bus mask.  All loops are fully unrolled: everything stays in registers.
struct LaneSends {
	int n;
	uint32_t mask;
	int bus[GAS_MAX_SENDS];
	float vp[GAS_MAX_SENDS];
	float vn[GAS_MAX_SENDS];
};

// K = number of detail slots examined on each side: the kernel uses K = 2 when neither side has more than two
// buses (almost always) and the full 6 otherwise; the result is the same, the unrolled code is 9x shorter.
template <int K>
__device__ __forceinline__ void resolve_sends_lane(const LaneDetails &cur, const LaneDetails &prev, const GlobalCfg &g, LaneSends &s) {
	int cn = cur.n, pn = prev.n;
	cn = cn < 0 ? 0 : (cn > K ? K : cn);
	pn = pn < 0 ? 0 : (pn > K ? K : pn);
	int ckey[K], pkey[K];
	float cvp[K];
	int total = cn;
#pragma unroll
	for (int k = 0; k < K; k++) {
		ckey[k] = k < cn ? resolve_bus(g, cur.bus[k]) * 16 + k : kBigKey;
		cvp[k] = 0.f;
#pragma unroll
		for (int j = 0; j < K; j++) {
			if (k < cn && j < pn && prev.bus[j] == cur.bus[k]) {
				cvp[k] = prev.vol[j]; // the last match wins, like a lookup that keeps scanning
			}
		}
	}
#pragma unroll
	for (int j = 0; j < K; j++) {
		bool only = j < pn;
#pragma unroll
		for (int k = 0; k < K; k++) {
			if (k < cn && cur.bus[k] == prev.bus[j]) {
				only = false;
			}
		}
		pkey[j] = only ? resolve_bus(g, prev.bus[j]) * 16 + K + j : kBigKey;
		total += only ? 1 : 0;
	}
	s.n = total;
	s.mask = 0;
	int last = -1;
#pragma unroll
	for (int i = 0; i < GAS_MAX_SENDS; i++) {
		s.bus[i] = 0;
		s.vp[i] = 0.f;
		s.vn[i] = 0.f;
		if (i < 2 * K && i < total) { // uniform over the 8 lanes of a voice
			int best = kBigKey;
			float bp = 0.f, bn = 0.f;
#pragma unroll
			for (int k = 0; k < K; k++) {
				if (ckey[k] > last && ckey[k] < best) {
					best = ckey[k];
					bp = cvp[k];
					bn = cur.vol[k];
				}
			}
#pragma unroll
			for (int j = 0; j < K; j++) {
				if (pkey[j] > last && pkey[j] < best) {
					best = pkey[j];
					bp = prev.vol[j];
					bn = 0.f;
				}
			}
			s.bus[i] = best >> 4;
			s.vp[i] = bp;
			s.vn[i] = bn;
			s.mask |= 1u << (best >> 4);
			last = best;
		}
	}
}

// reductions over the 8 lanes of one voice (gm = those lanes' bits in the warp)
__device__ __forceinline__ int group_or(unsigned gm, int v) {
	v |= __shfl_xor_sync(gm, v, 1);
	v |= __shfl_xor_sync(gm, v, 2);
	v |= __shfl_xor_sync(gm, v, 4);
	return v;
}

// MINB: CTAs per SM the register allocation is capped for.  4 (122 registers, no spills) makes the grid of a 16384-voice
// block 1.7 waves; 7 (72 registers, a few hundred bytes of spills in the rare six-bus path) makes it one.
template <int MINB>
__global__ void __launch_bounds__(kCtaThreads, MINB) k_prologue(DevTables t, GlobalCfg g, BlockPlan plan, int inst_hwm, int n_voices,
		const gas_voice *__restrict__ voices, int src_rows, float4 *__restrict__ bus, int bus_f4, float4 *__restrict__ rep, int rep_f4,
		float2 *__restrict__ peaks, int g_scaled_classes) {
	__shared__ unsigned long long s_key[kVoicesPerCta];  // classes met in this CTA
	__shared__ int s_cnt[kVoicesPerCta], s_cid[kVoicesPerCta], s_base[kVoicesPerCta];
	__shared__ unsigned long long s_gkey[GAS_MAX_CLASSES]; // snapshot of the global slot table
	__shared__ unsigned long long s_gaux[GAS_MAX_CLASSES];
	__shared__ unsigned long long s_aux[kVoicesPerCta];    // aux word of the classes met in this CTA
	__shared__ int s_parity;

	const int tid = blockIdx.x * blockDim.x + threadIdx.x;
	const int nthreads = gridDim.x * blockDim.x;
	const int grp = threadIdx.x >> 3;       // voice slot inside the CTA
	const int l = threadIdx.x & 7;          // lane of the voice
	const int c = l >> 1, x = l & 1;        // channel pair, side
	const unsigned gm = 0xffu << (threadIdx.x & 24);
	const int C = g.channels;
	const int maxv = g.max_voices;
	const int j = blockIdx.x * kVoicesPerCta + grp;
	int blk_n = 0, ticket = -1;

	GAS_GRID_DEP_WAIT(); // programmatic dependent launch: the previous block's kernels are complete after this
	GAS_GRID_DEP_LAUNCH();

	// ---- level 0: everything that needs no other load -------------------------------------------------------
	gas_voice v{};
	v.voice = -1;
	if (j < n_voices) {
		v = voices[j];
	}
	if (threadIdx.x == 0) {
		// Block counter: read it, then take this CTA's ticket.  The last CTA to take one advances the counter
		// for the kernels that follow; every CTA has read the old value by then (its read precedes its ticket).
		blk_n = *(volatile int32_t *)&t.blk[0];
		s_parity = blk_n & 1;
		ticket = atomicAdd(&t.blk[1], 1); // consumed at the very end: nothing waits for this round trip
	}
	for (int i = threadIdx.x; i < GAS_MAX_CLASSES; i += kCtaThreads) {
		s_gkey[i] = plan.cls_key[i];
		s_gaux[i] = plan.cls_aux[i];
	}
	if (threadIdx.x < kVoicesPerCta) {
		s_key[threadIdx.x] = 0ULL;
		s_aux[threadIdx.x] = CLS_AUX_NONE;
		s_cnt[threadIdx.x] = 0;
		s_cid[threadIdx.x] = -1;
		s_base[threadIdx.x] = 0;
	}
	const int qi = blockIdx.x * kVoicesPerCta + grp; // instance part: one instance per 8-lane group
	int i_active = 0;
	LaneDetails icur;
	icur.n = 0;
	if (qi < inst_hwm) {
		i_active = t.inst_active[qi];
		details_load_lane(icur, &t.inst_cur[qi], c, x);
	}
	__syncthreads();
	const int parity = s_parity;
	int32_t *cnt_now = plan.cls_count + parity * GAS_MAX_CLASSES;
	const BusDetails *prev_rd = t.inst_prev + (size_t)parity * t.max_instances;
	BusDetails *prev_wr = t.inst_prev + (size_t)(parity ^ 1) * t.max_instances;

	// ---- level 1: everything behind the voice record ------------------------------------------------------------
	const bool valid = v.voice >= 0 && v.voice < maxv && v.instance >= 0 && v.instance < g.max_instances;
	const int q = valid ? v.instance : 0;
	const int vslot = valid ? v.voice : 0;
	const int v_active = t.inst_active[q];
	const int imode = t.inst_mode[q];
	const gas_params *prm = &t.inst_params[q];
	const float lin_att = prm->linear_attenuation;
	const float cutoff = prm->attenuation_filter_cutoff_hz;
	const float mixv = prm->mix_volumes[c][x];
	LaneDetails cur, prev;
	details_load_lane(cur, &t.inst_cur[q], c, x);
	details_load_lane(prev, &prev_rd[q], c, x);
	float *vprev = t.vs_prev + (size_t)vslot * 8;
	const float vp_l = vprev[l], vp_0 = vprev[0], vp_1 = vprev[1];

	// ---- housekeeping stores --------------------------------------------------------------------------------------
	for (int i = tid; i < bus_f4; i += nthreads) {
		bus[i] = make_float4(0.f, 0.f, 0.f, 0.f);
	}
	for (int i = tid; i < rep_f4; i += nthreads) {
		rep[i] = make_float4(0.f, 0.f, 0.f, 0.f);
	}
	if (peaks) {
		for (int i = tid; i < n_voices; i += nthreads) {
			peaks[i] = make_float2(0.f, 0.f);
		}
	}
	if (tid < GAS_MAX_CLASSES) { // class counts of the next block
		plan.cls_count[(parity ^ 1) * GAS_MAX_CLASSES + tid] = 0;
	}

	// ---- instance part: prev <- cur ---------------------------------------------------------------------
	if (qi < inst_hwm && i_active) {
		BusDetails *pw = &prev_wr[qi];
		int cn = icur.n;
		cn = cn < 0 ? 0 : (cn > GAS_MAX_BUSES_PER_PLAYBACK ? GAS_MAX_BUSES_PER_PLAYBACK : cn);
		if (l == 0) {
			pw->n = cn;
		}
#pragma unroll
		for (int k = 0; k < GAS_MAX_BUSES_PER_PLAYBACK; k++) {
			if (k < cn) {
				if (l == 0) {
					pw->bus[k] = icur.bus[k];
				}
				pw->vol[k][c][x] = icur.vol[k];
			}
		}
	}
	for (int q2 = qi + gridDim.x * kVoicesPerCta; q2 < inst_hwm; q2 += gridDim.x * kVoicesPerCta) { // only if the grid is smaller than the table
		if (!t.inst_active[q2]) {
			continue;
		}
		LaneDetails d;
		details_load_lane(d, &t.inst_cur[q2], c, x);
		BusDetails *pw = &prev_wr[q2];
		const int cn = d.n < 0 ? 0 : (d.n > GAS_MAX_BUSES_PER_PLAYBACK ? GAS_MAX_BUSES_PER_PLAYBACK : d.n);
		if (l == 0) {
			pw->n = cn;
		}
#pragma unroll
		for (int k = 0; k < GAS_MAX_BUSES_PER_PLAYBACK; k++) {
			if (k < cn) {
				if (l == 0) {
					pw->bus[k] = d.bus[k];
				}
				pw->vol[k][c][x] = d.vol[k];
			}
		}
	}

	// ---- voice part ----------------------------------------------------------------------------------------
	int path = PATH_NONE, mode = MODE_A, n_send = 0, n_group = 0, n_rows = 0;
	uint32_t cflags = 0, mask = 0, quad = 0, rflags = 0;
	const bool live = valid && v_active != 0;
	if (v.src_row >= src_rows) {
		v.src_row = -1;
	}
	LaneSends snd;
	snd.n = 0;
	float m_prev = 1.f, m_new = 1.f; // this lane's (pair, side) element of the mix_channel ramp
	float target[5] = { 0.f, 0.f, 0.f, 0.f, 0.f };
	int n_fx = 0, fx_stage = 1;
	float fx_coef[5] = { 0.f, 0.f, 0.f, 0.f, 0.f };
	bool shared = false;
	unsigned long long aux = CLS_AUX_NONE; // second word of the class identity

	if (live) {
		mode = imode & 0xff;
		const int fx_binding = (imode >> 8) - 1;
		if (cur.n <= 2 && prev.n <= 2) {
			resolve_sends_lane<2>(cur, prev, g, snd);
		} else {
			resolve_sends_lane<GAS_MAX_BUSES_PER_PLAYBACK>(cur, prev, g, snd);
		}
		n_send = snd.n;
		mask = snd.mask;
		const bool filt = mode != MODE_E && (double)lin_att >= 0.001; // audio_spatializer_3d.cpp:503, :568
		const bool want_peak = (v.flags & GAS_VOICE_WANT_PEAK) != 0;
		rflags = v.flags & 0xffu;
		if (mode == MODE_B) {
			if (c < C) {
				m_prev = vp_l;     // :564
				m_new = mixv;      // :565
				vprev[l] = m_new;  // :608
			}
			// is_just_started per pair: previous (L, R) exactly (0, 0), :583
			const int zero = (c < C && m_prev == 0.f) ? 1 : 0;
			const int both = zero & __shfl_xor_sync(gm, zero, 1);
			rflags |= (uint32_t)group_or(gm, (both && x == 0) ? (1 << (8 + c)) : 0);
		} else if (mode == MODE_A) {
			if (vp_0 == 0.f && vp_1 == 0.f) {
				rflags |= 1u << 8; // :518
			}
			// :537-551 — the (L,R) pair holding the first maximum in scan order c0.L, c0.R, c1.L, ...
			float bv = mixv;
			int bi = l;
#pragma unroll
			for (int d = 1; d < 8; d <<= 1) {
				const float ov = __shfl_xor_sync(gm, bv, d);
				const int oi = __shfl_xor_sync(gm, bi, d);
				if (ov > bv || (ov == bv && oi < bi)) {
					bv = ov;
					bi = oi;
				}
			}
			const int max_index = bv > 0.f ? (bi >> 1) : 0;
			const float keep = __shfl_sync(gm, mixv, (threadIdx.x & 24) + max_index * 2 + x);
			if (l < 2) {
				vprev[l] = keep;
			}
		}
		if (filt) {
			cflags |= CLS_FILT;
			if (l == 0) {
				prepare_coefficients(GAS_FILTER_HIGHSHELF, cutoff, 1.0f, lin_att, 1, g.mix_rate, target); // :504-510
			}
		}
		if (mode == MODE_E) {
			const gas_effect_chain *fx = &t.inst_fx[q];
			n_fx = fx->n_effects;
			n_fx = n_fx < 0 ? 0 : (n_fx > GAS_MAX_EFFECTS ? GAS_MAX_EFFECTS : n_fx);
			if (l < n_fx) { // one effect per lane
				gas_effect ef = fx->effects[l];
				if (fx_binding == l) {
					ef.gain = lin_att; // example _process_effects (gd_spatializer_instance.gd:125-127)
				}
				fx_stage = ef.stages < 1 ? 1 : (ef.stages > GAS_MAX_FILTER_STAGES ? GAS_MAX_FILTER_STAGES : ef.stages);
				prepare_coefficients(ef.mode, ef.cutoff_hz, ef.resonance, ef.gain, fx_stage, g.mix_rate, fx_coef);
			}
		}
		const bool has_dsp = filt || (mode == MODE_E && n_fx > 0);
		if (mode == MODE_B) {
			// Every Mode-B proxy is a playback of its own: AudioServer runs _mix_step_for_channel for every pair of every
			// bus of its map, with volume 0 for the pairs the map masks out (reference audio_spatializer.cpp:298-312).
			// 0 * x only matters when the proxy's buffer is not finite, which the module produces itself (NaN pan gains,
			// SURVEY Q1): the NaN then reaches every pair of every bus the instance sends to, same side.  A NaN ramp end
			// point of any pair therefore poisons this side's send volumes of the voice.
			int bad = (c < C && (m_prev != m_prev || m_new != m_new)) ? 1 : 0;
			bad |= __shfl_xor_sync(gm, bad, 2);
			bad |= __shfl_xor_sync(gm, bad, 4);
			if (bad) {
				const float qnan = __int_as_float(0x7fc00000);
#pragma unroll
				for (int k = 0; k < GAS_MAX_SENDS; k++) {
					if (k < n_send) {
						snd.vn[k] = qnan;
					}
				}
			}
		}

		// weight polynomial per (send, pair, side): w(t) = A + B t + Cq t^2 with t = i/F, from
		// (vn*t + (1-t)*vp) of the AudioServer ramp times (m_new*t + (1-t)*m_prev) of mix_channel.
		bool streamed = false;
		if (!has_dsp && !want_peak && n_send >= 1) {
			const float dm = m_new - m_prev;
			int q_bits = 0, differs = 0;
#pragma unroll
			for (int k = 0; k < GAS_MAX_SENDS; k++) {
				if (k < n_send && c < C) {
					const float dn = snd.vn[k] - snd.vp[k];
					if (dn * dm != 0.f) {
						q_bits |= 1 << k;
					}
					if (snd.vn[k] != snd.vn[0] || snd.vp[k] != snd.vp[0]) {
						differs = 1;
					}
				}
			}
			q_bits = group_or(gm, q_bits);
			differs = group_or(gm, differs);
			shared = n_send >= 2 && !differs;
			// Scaled sends: every further send is send 0 times ONE scalar (same for both ramp end points, all pairs, both
			// sides) — what a reverb send with uniformity 0 is (reverb_vol = direct * area_send, reference
			// audio_spatializer_3d.cpp:192-196, so bus_vol / mix_vol is area_send to the last bit or two on every pair).
			// Such a voice needs one row group: the flush adds the sums to bus 0 as they are and to the other buses times
			// the class's scales.  The scalar of a send is taken from the first lane with a non-zero base volume; a lane
			// accepts it if it reproduces its own volumes within 3e-7 relative (2-3 ulp: far inside the 1e-5 tolerance).
			bool scaled = false;
			float sc1 = 0.f, sc2 = 0.f;
			if (n_send >= 2 && n_send <= 3 && !shared && g_scaled_classes) {
				const bool base_nz = c < C && (snd.vn[0] != 0.f || snd.vp[0] != 0.f);
				const unsigned nzm = __ballot_sync(gm, base_nz) & gm;
				int ok = nzm != 0u;
				if (ok) {
					const int src_lane = __ffs(nzm) - 1;
					const float bn = snd.vn[0], bp = snd.vp[0];
					const bool use_n = bn != 0.f;
					const float r1 = use_n ? snd.vn[1] / bn : snd.vp[1] / bp;
					const float r2 = n_send > 2 ? (use_n ? snd.vn[2] / bn : snd.vp[2] / bp) : 0.f;
					sc1 = __shfl_sync(gm, r1, src_lane);
					sc2 = __shfl_sync(gm, r2, src_lane);
					if (c < C) {
						const float tol = 3e-7f;
						ok = fabsf(snd.vn[1] - sc1 * bn) <= tol * fabsf(snd.vn[1]) && fabsf(snd.vp[1] - sc1 * bp) <= tol * fabsf(snd.vp[1]);
						if (n_send > 2) {
							ok = ok && fabsf(snd.vn[2] - sc2 * bn) <= tol * fabsf(snd.vn[2]) && fabsf(snd.vp[2] - sc2 * bp) <= tol * fabsf(snd.vp[2]);
						}
						ok = ok && (sc1 == sc1) && (sc2 == sc2) && fabsf(sc1) < 3.0e38f && fabsf(sc2) < 3.0e38f;
					}
				}
				ok = !group_or(gm, ok ? 0 : 1);
				scaled = ok != 0;
			}
			n_group = (shared || scaled) ? 1 : n_send;
			quad = (shared || scaled) ? ((scaled ? (q_bits & 1) : q_bits) ? 1u : 0u) : (uint32_t)q_bits;
			n_rows = 2 * n_group + __popc(quad);
			if (n_rows <= GAS_K2_MAX_ROWS) {
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
				n_rows = 0;
				quad = 0;
			}
		}
	}

	// ---- class lookup: once per class per CTA in shared memory, then one global atomic per class ------------
	// A class is (key, aux).  The key of a scaled class carries a 20-bit hash of its aux word in its spare bits, so that
	// the compare-and-swap that claims a slot sees (almost always) the whole identity; the aux words are compared once
	// they are published (after the barrier in the CTA table; after a short wait in the global table).
	unsigned long long key = path != PATH_NONE ? cls_key(path, mode, cflags, n_send, mask, quad) : 0ULL;
	if (key != 0ULL && aux != CLS_AUX_NONE) {
		key |= ((aux * 0x9E3779B97F4A7C15ULL) >> 44) << 44;
	}
	int slot = -1, lpos = 0;
	if (key != 0ULL && l == 0) {
		for (int i = 0; i < kVoicesPerCta; i++) {
			unsigned long long k = *(volatile unsigned long long *)&s_key[i];
			if (k == 0ULL) {
				k = atomicCAS(&s_key[i], 0ULL, key);
				if (k == 0ULL) {
					k = key;
					s_aux[i] = aux;
				}
			}
			if (k == key) {
				slot = i;
				break;
			}
		}
	}
	__syncthreads();
	if (slot >= 0 && s_aux[slot] != aux) {
		// two scaled classes whose aux words hash alike met in one CTA (one in 2^20 pairs): this voice takes the generic
		// class of the voice-parallel kernel instead, which mixes any voice
		slot = -2;
	}
	if (slot >= 0) {
		lpos = atomicAdd(&s_cnt[slot], 1);
	}
	int gpos = 0; // position in the generic class when slot == -2
	const int generic_cid = GAS_CLS_DYNAMIC + mode * 2 + ((cflags & CLS_FILT) ? 1 : 0);
	if (slot == -2) {
		gpos = atomicAdd(&cnt_now[generic_cid], 1);
	}
	__syncthreads();
	if (threadIdx.x < kVoicesPerCta && s_key[threadIdx.x] != 0ULL && s_cnt[threadIdx.x] > 0) {
		// Slots are stable across blocks: in the steady state the class is already in the snapshot and the only
		// global operation is the add that reserves this CTA's range of the class list.
		const unsigned long long k = s_key[threadIdx.x];
		const unsigned long long ka = s_aux[threadIdx.x];
		int cid = -1;
		for (int i = 0; i < GAS_CLS_DYNAMIC; i++) {
			if (s_gkey[i] == k && s_gaux[i] == ka) {
				cid = i;
				break;
			}
		}
		if (cid < 0) { // first appearance of the class: claim a free slot, or find the one another CTA just claimed
			for (int i = 0; i < GAS_CLS_DYNAMIC && cid < 0; i++) {
				unsigned long long o = s_gkey[i];
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
					unsigned long long a = *(volatile unsigned long long *)&plan.cls_aux[i];
					if (ka != CLS_AUX_NONE) {
						for (int spin = 0; spin < 256 && a == CLS_AUX_NONE; spin++) {
							a = *(volatile unsigned long long *)&plan.cls_aux[i];
						}
					}
					if (a == ka) {
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
		s_base[threadIdx.x] = atomicAdd(&cnt_now[cid], s_cnt[threadIdx.x]);
		s_cid[threadIdx.x] = cid;
	}
	__syncthreads();
	slot = __shfl_sync(gm, slot, threadIdx.x & 24);
	lpos = __shfl_sync(gm, lpos, threadIdx.x & 24);
	gpos = __shfl_sync(gm, gpos, threadIdx.x & 24);
	const int cid = slot >= 0 ? s_cid[slot] : (slot == -2 ? generic_cid : -1);
	if (cid >= 0) {
		const int pos = slot >= 0 ? s_base[slot] + lpos : gpos;
		if (l == 0) {
			plan.list[(size_t)cid * maxv + pos] = make_int2(j, v.src_row);
		}
		if (cid >= GAS_CLS_DYNAMIC) {
			path = PATH_VOICE; // generic class: the voice-parallel kernel mixes any voice
		}
		if (path == PATH_STREAM) {
			if (c < C) {
				const int nf = n_rows * C * 2;
				float *dst = plan.k2_rows + (size_t)cid * maxv * GAS_K2_ROW_FLOATS + (size_t)pos * nf + c * 2 + x;
				const float mp = m_prev, dm = m_new - m_prev;
				int r = 0;
#pragma unroll
				for (int k = 0; k < GAS_K2_MAX_ROWS / 2; k++) {
					if (k < n_group) {
						const float np = snd.vp[k], dn = snd.vn[k] - np;
						dst[(size_t)(r + 0) * C * 2] = np * mp;
						dst[(size_t)(r + 1) * C * 2] = np * dm + dn * mp;
						if ((quad >> k) & 1u) {
							dst[(size_t)(r + 2) * C * 2] = dn * dm;
							r += 3;
						} else {
							r += 2;
						}
					}
				}
			}
		} else {
			VoiceRec *rec = &plan.rec[j];
			InstSends *ps = &plan.sends[j];
			rec->m_prev[c][x] = m_prev;
			rec->m_new[c][x] = m_new;
			if (l == 0) {
				rec->voice = v.voice;
				rec->instance = v.instance;
				rec->src_row = v.src_row;
				rec->flags = rflags;
				rec->n_fx = n_fx;
#pragma unroll
				for (int i = 0; i < 5; i++) {
					rec->target[i] = target[i];
				}
				ps->n = n_send;
				ps->mask = mask;
			}
			if (l < n_fx) {
				rec->fx_stages[l] = fx_stage;
#pragma unroll
				for (int i = 0; i < 5; i++) {
					rec->fx_coef[l][i] = fx_coef[i];
				}
			}
#pragma unroll
			for (int k = 0; k < GAS_MAX_SENDS; k++) {
				if (k < n_send) {
					if (l == 0) {
						ps->bus[k] = snd.bus[k];
					}
					ps->vp[k][c][x] = snd.vp[k];
					ps->vn[k][c][x] = snd.vn[k];
				}
			}
		}
	}
	if (threadIdx.x == 0 && ticket == (int)gridDim.x - 1) {
		t.blk[1] = 0;
		t.blk[0] = blk_n + 1;
	}
}

} // namespace

cudaError_t launch_prologue(gas_ctx *ctx, int n_voices, const gas_voice *d_voices, int src_rows, int frames, gas_frame *d_bus,
		gas_frame *d_peaks, cudaStream_t st) {
	const int bus_f4 = gas_bus_f4(ctx, frames);
	const int rep_f4 = ctx->replicas > 1 ? ctx->replicas * bus_f4 : 0;
	int work = ctx->inst_hwm > n_voices ? ctx->inst_hwm : n_voices;
	work = work > 1 ? work : 1;
	const int blocks = (work + kVoicesPerCta - 1) / kVoicesPerCta;
	static int minb = -1;
	if (minb < 0) {
		const char *e = getenv("GAS_PROLOGUE_MINB");
		minb = e ? atoi(e) : 7; // one wave for a 16384-voice block: measured 0.7-2 us faster per step than the 122-register variant
	}
	cudaError_t e;
#define GAS_PRO_LAUNCH(M_)                                                                                                                       \
	e = gas_launch(k_prologue<M_>, dim3(blocks), dim3(kCtaThreads), 0, st, (ctx->pdl & 1) != 0, ctx->t, ctx->g, ctx->plan, ctx->inst_hwm, n_voices, \
			d_voices, src_rows, (float4 *)d_bus, bus_f4, (float4 *)ctx->d_rep, rep_f4, (float2 *)d_peaks, ctx->scaled_classes ? 1 : 0)
	switch (minb) {
		case 6: GAS_PRO_LAUNCH(6); break;
		case 7: GAS_PRO_LAUNCH(7); break;
		case 8: GAS_PRO_LAUNCH(8); break;
		default: GAS_PRO_LAUNCH(4); break;
	}
#undef GAS_PRO_LAUNCH
	ctx->launches++;
	return e;
}
