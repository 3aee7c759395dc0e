// gas_prologue.cu — per-block prologue (one fused kernel): turns the current parameters + persistent ramp
// state into the block's plan (classes, weight rows, voice records) and advances the ramp state.
//
//   instance part (thread q < inst_hwm): the AudioServer side of a mix step for the instance's proxy
//       playbacks — previous volume looked up by bus, buses that disappeared fade to 0, then prev <- cur
//       (upstream AudioServer::_mix_step).  prev is double-buffered by block parity:
//       this block reads inst_prev[p] and writes inst_prev[1-p], so the voice part of other threads can
//       read the old value without a grid-wide barrier.
//   voice part (thread j < n_voices): what process_frames / mix_channel decide before their sample loop
//       (reference audio_spatializer_3d.cpp:499-523, :562-587, :537-551, :608): ramp end points, filter
//       on/off, clear-history, target coefficients; classifies the voice and appends it to its class list.
//   The kernel also zeroes the bus buffers / peaks, clears the class table of the NEXT block, and its
//   last CTA advances the block counter.
//
// Compiled with -fmad=false (coefficient preparation is double arithmetic narrowed to float).
#include "gas_internal.h"

namespace {

__device__ __forceinline__ int resolve_bus(const GlobalCfg &g, int bus) {
	return (bus >= 0 && bus < g.num_buses) ? bus : 0;
}

// Sends of one instance for this block: every bus of the current details with the previous volume
// looked up by bus (absent => 0 => fade-in), then buses only present in the previous details once more
// towards 0 (fade-out); ascending by bus so that a class is identified by its bus mask.
__device__ void resolve_sends(const BusDetails &cur, const BusDetails &prev, const GlobalCfg &g, InstSends &s) {
	s.n = 0;
	s.mask = 0;
	for (int k = 0; k < cur.n; k++) {
		int pk = -1;
		for (int j = 0; j < prev.n; j++) {
			if (prev.bus[j] == cur.bus[k]) {
				pk = j;
			}
		}
		const int slot = s.n++;
		s.bus[slot] = resolve_bus(g, cur.bus[k]);
		for (int c = 0; c < 4; c++) {
			for (int x = 0; x < 2; x++) {
				s.vp[slot][c][x] = pk >= 0 ? prev.vol[pk][c][x] : 0.f;
				s.vn[slot][c][x] = cur.vol[k][c][x];
			}
		}
	}
	for (int j = 0; j < prev.n; j++) {
		bool still = false;
		for (int k = 0; k < cur.n; k++) {
			still |= (cur.bus[k] == prev.bus[j]);
		}
		if (still) {
			continue;
		}
		const int slot = s.n++;
		s.bus[slot] = resolve_bus(g, prev.bus[j]);
		for (int c = 0; c < 4; c++) {
			for (int x = 0; x < 2; x++) {
				s.vp[slot][c][x] = prev.vol[j][c][x];
				s.vn[slot][c][x] = 0.f;
			}
		}
	}
	for (int a = 1; a < s.n; a++) { // insertion sort, <= 12 entries
		for (int b = a; b > 0 && s.bus[b - 1] > s.bus[b]; b--) {
			const int tb = s.bus[b];
			s.bus[b] = s.bus[b - 1];
			s.bus[b - 1] = tb;
			for (int c = 0; c < 4; c++) {
				for (int x = 0; x < 2; x++) {
					const float tp = s.vp[b][c][x], tn = s.vn[b][c][x];
					s.vp[b][c][x] = s.vp[b - 1][c][x];
					s.vn[b][c][x] = s.vn[b - 1][c][x];
					s.vp[b - 1][c][x] = tp;
					s.vn[b - 1][c][x] = tn;
				}
			}
		}
	}
	for (int k = 0; k < s.n; k++) {
		s.mask |= 1u << s.bus[k];
	}
}

// only the first `n` entries of a BusDetails are meaningful: load just those
__device__ __forceinline__ void details_load(BusDetails &d, const BusDetails *src) {
	d.n = src->n;
	for (int k = 0; k < d.n && k < GAS_MAX_BUSES_PER_PLAYBACK; k++) {
		d.bus[k] = src->bus[k];
		for (int c = 0; c < 4; c++) {
			d.vol[k][c][0] = src->vol[k][c][0];
			d.vol[k][c][1] = src->vol[k][c][1];
		}
	}
}

// upstream AudioFilterSW::prepare_coefficients: double arithmetic, every coefficient
// narrowed to float on store and once more after the division by a0; feedback terms stored negated.
__device__ void prepare_coefficients(int mode, float cutoff, float resonance, float gain, int stages, float sampling_rate, float out[5]) {
	int sr_limit = (int)((sampling_rate / 2) + 512);
	double final_cutoff = (cutoff > sr_limit) ? (double)sr_limit : (double)cutoff;
	if (final_cutoff < 1) {
		final_cutoff = 1;
	}
	const double TAU = 6.2831853071795864769252867666;
	double omega = TAU * final_cutoff / (double)sampling_rate;
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
	float b0 = 0.f, b1 = 0.f, b2 = 0.f, a1 = 0.f, a2 = 0.f;
	switch (mode) {
		case GAS_FILTER_LOWPASS:
			b0 = (float)((1.0 - cos_v) / 2.0);
			b1 = (float)(1.0 - cos_v);
			b2 = (float)((1.0 - cos_v) / 2.0);
			a1 = (float)(-2.0 * cos_v);
			a2 = (float)(1.0 - alpha);
			break;
		case GAS_FILTER_HIGHPASS:
			b0 = (float)((1.0 + cos_v) / 2.0);
			b1 = (float)(-(1.0 + cos_v));
			b2 = (float)((1.0 + cos_v) / 2.0);
			a1 = (float)(-2.0 * cos_v);
			a2 = (float)(1.0 - alpha);
			break;
		case GAS_FILTER_BANDPASS:
			b0 = (float)(alpha * sqrt(Q + 1));
			b1 = 0.f;
			b2 = (float)(-alpha * sqrt(Q + 1));
			a1 = (float)(-2.0 * cos_v);
			a2 = (float)(1.0 - alpha);
			break;
		case GAS_FILTER_NOTCH:
			b0 = 1.f;
			b1 = (float)(-2.0 * cos_v);
			b2 = 1.f;
			a1 = (float)(-2.0 * cos_v);
			a2 = (float)(1.0 - alpha);
			break;
		case GAS_FILTER_PEAK:
			b0 = (float)(1.0 + alpha * tmpgain);
			b1 = (float)(-2.0 * cos_v);
			b2 = (float)(1.0 - alpha * tmpgain);
			a1 = (float)(-2 * cos_v);
			a2 = (float)(1 - alpha / tmpgain);
			break;
		case GAS_FILTER_BANDLIMIT: {
			double hicutoff = resonance;
			double centercutoff = ((double)cutoff + (double)resonance) / 2.0;
			double bandwidth = (log(centercutoff) - log(hicutoff)) / log(2.0);
			omega = TAU * centercutoff / (double)sampling_rate;
			alpha = sin(omega) * sinh(log(2.0) / 2 * bandwidth * omega / sin(omega));
			a0 = 1 + alpha;
			b0 = (float)alpha;
			b1 = 0.f;
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
		default: { // HIGHSHELF
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

__device__ int class_find_or_insert(ClassInfo *cls, unsigned long long key, int *overflow) {
	for (int i = 0; i < GAS_MAX_CLASSES; i++) {
		unsigned long long k = *(volatile unsigned long long *)&cls[i].key;
		if (k == key) {
			return i;
		}
		if (k == 0ULL) {
			k = atomicCAS(&cls[i].key, 0ULL, key);
			if (k == 0ULL || k == key) {
				return i;
			}
		}
	}
	*overflow = 1;
	return -1;
}


__global__ void __launch_bounds__(128) k_prologue(DevTables t, GlobalCfg g, BlockPlan plan, int inst_hwm, int n_voices,
		const gas_voice *__restrict__ voices, int src_rows, float4 *__restrict__ bus, int bus_f4, float2 *__restrict__ peaks) {
	const int tid = blockIdx.x * blockDim.x + threadIdx.x;
	const int nthreads = gridDim.x * blockDim.x;
	const unsigned lane = threadIdx.x & 31u;
	const int C = g.channels;
	const int maxv = g.max_voices;
	const int parity = t.blk[0] & 1;
	ClassInfo *cls = plan.cls + parity * GAS_MAX_CLASSES;
	const BusDetails *prev_rd = t.inst_prev + (size_t)parity * t.max_instances;
	BusDetails *prev_wr = t.inst_prev + (size_t)(parity ^ 1) * t.max_instances;

	// ---- housekeeping -------------------------------------------------------------------------------
	for (int i = tid; i < bus_f4; i += nthreads) {
		bus[i] = make_float4(0.f, 0.f, 0.f, 0.f);
	}
	if (peaks) {
		for (int i = tid; i < n_voices; i += nthreads) {
			peaks[i] = make_float2(0.f, 0.f);
		}
	}
	if (tid < GAS_MAX_CLASSES) { // class table of the next block
		ClassInfo z{};
		plan.cls[(parity ^ 1) * GAS_MAX_CLASSES + tid] = z;
	}

	// ---- instance part ----------------------------------------------------------------------------------
	for (int q = tid; q < inst_hwm; q += nthreads) {
		if (!t.inst_active[q]) {
			continue;
		}
		BusDetails cur, prev;
		details_load(cur, &t.inst_cur[q]);
		details_load(prev, &prev_rd[q]);
		InstSends s;
		resolve_sends(cur, prev, g, s);
		InstSends *dst = &t.inst_sends[q]; // read by the voice-parallel kernel
		dst->n = s.n;
		dst->mask = s.mask;
		for (int k = 0; k < s.n; k++) {
			dst->bus[k] = s.bus[k];
			for (int c = 0; c < 4; c++) {
				for (int x = 0; x < 2; x++) {
					dst->vp[k][c][x] = s.vp[k][c][x];
					dst->vn[k][c][x] = s.vn[k][c][x];
				}
			}
		}
		BusDetails *pw = &prev_wr[q]; // prev <- cur
		pw->n = cur.n;
		for (int k = 0; k < cur.n; k++) {
			pw->bus[k] = cur.bus[k];
			for (int c = 0; c < 4; c++) {
				pw->vol[k][c][0] = cur.vol[k][c][0];
				pw->vol[k][c][1] = cur.vol[k][c][1];
			}
		}
	}

	// ---- voice part ----------------------------------------------------------------------------------------
	const int j = tid;
	int path = PATH_NONE, mode = MODE_A, n_send = 0, n_group = 0, n_rows = 0;
	uint32_t cflags = 0, mask = 0;
	float rows[GAS_K2_ROW_FLOATS];
	VoiceRec rec;
	gas_voice v{};
	bool live = false;

	if (j < n_voices) {
		v = voices[j];
		live = v.voice >= 0 && v.voice < maxv && v.instance >= 0 && v.instance < g.max_instances && t.inst_active[v.instance] != 0;
		if (v.src_row >= src_rows) {
			v.src_row = -1;
		}
	}
	if (live) {
		const int q = v.instance;
		const int imode = t.inst_mode[q];
		const gas_params *prm = &t.inst_params[q];
		mode = imode & 0xff;
		const int fx_binding = (imode >> 8) - 1;
		InstSends snd;
		{
			BusDetails cur, prev;
			details_load(cur, &t.inst_cur[q]);
			details_load(prev, &prev_rd[q]);
			resolve_sends(cur, prev, g, snd);
		}
		n_send = snd.n;
		mask = snd.mask;
		const float lin_att = prm->linear_attenuation;
		const bool filt = mode != MODE_E && (double)lin_att >= 0.001; // audio_spatializer_3d.cpp:503, :568
		const bool want_peak = (v.flags & GAS_VOICE_WANT_PEAK) != 0;

		rec.voice = v.voice;
		rec.instance = q;
		rec.src_row = v.src_row;
		rec.flags = v.flags & 0xffu;
		rec.n_fx = 0;
		float *vprev = t.vs_prev + (size_t)v.voice * 8;
		for (int c = 0; c < 4; c++) {
			rec.m_prev[c][0] = rec.m_prev[c][1] = 1.f;
			rec.m_new[c][0] = rec.m_new[c][1] = 1.f;
		}
		if (mode == MODE_B) {
			for (int c = 0; c < C; c++) {
				rec.m_prev[c][0] = vprev[c * 2 + 0]; // :564
				rec.m_prev[c][1] = vprev[c * 2 + 1];
				rec.m_new[c][0] = prm->mix_volumes[c][0]; // :565
				rec.m_new[c][1] = prm->mix_volumes[c][1];
				if (rec.m_prev[c][0] == 0.f && rec.m_prev[c][1] == 0.f) {
					rec.flags |= 1u << (8 + c); // is_just_started, :583
				}
				vprev[c * 2 + 0] = rec.m_new[c][0]; // :608
				vprev[c * 2 + 1] = rec.m_new[c][1];
			}
		} else if (mode == MODE_A) {
			if (vprev[0] == 0.f && vprev[1] == 0.f) {
				rec.flags |= 1u << 8; // :518
			}
			float max_volume = 0.f; // :537-551
			int max_index = 0;
			for (int c = 0; c < 4; c++) {
				if (prm->mix_volumes[c][0] > max_volume) {
					max_volume = prm->mix_volumes[c][0];
					max_index = c;
				}
				if (prm->mix_volumes[c][1] > max_volume) {
					max_volume = prm->mix_volumes[c][1];
					max_index = c;
				}
			}
			vprev[0] = prm->mix_volumes[max_index][0];
			vprev[1] = prm->mix_volumes[max_index][1];
		}
		if (filt) {
			cflags |= CLS_FILT;
			prepare_coefficients(GAS_FILTER_HIGHSHELF, prm->attenuation_filter_cutoff_hz, 1.0f, lin_att, 1, g.mix_rate, rec.target); // :504-510
		}
		if (mode == MODE_E) {
			const gas_effect_chain *fx = &t.inst_fx[q];
			int nfx = fx->n_effects;
			nfx = nfx < 0 ? 0 : (nfx > GAS_MAX_EFFECTS ? GAS_MAX_EFFECTS : nfx);
			rec.n_fx = nfx;
			for (int e = 0; e < nfx; e++) {
				gas_effect ef = fx->effects[e];
				if (fx_binding == e) {
					ef.gain = lin_att; // example _process_effects (gd_spatializer_instance.gd:125-127)
				}
				int st = ef.stages < 1 ? 1 : (ef.stages > GAS_MAX_FILTER_STAGES ? GAS_MAX_FILTER_STAGES : ef.stages);
				rec.fx_stages[e] = st;
				prepare_coefficients(ef.mode, ef.cutoff_hz, ef.resonance, ef.gain, st, g.mix_rate, rec.fx_coef[e]);
			}
		}
		const bool has_dsp = filt || (mode == MODE_E && rec.n_fx > 0);

		// weight polynomial per (send, pair, side): w(t) = A + B t + Cq t^2 with t = i/F, from
		// (vn*t + (1-t)*vp) of the AudioServer ramp times (m_new*t + (1-t)*m_prev) of mix_channel.
		bool lin = true, shared = n_send >= 2, streamed = false;
		if (!has_dsp && !want_peak && n_send >= 1) {
			for (int k = 0; k < n_send; k++) {
				for (int c = 0; c < C; c++) {
					for (int x = 0; x < 2; x++) {
						const float dn = snd.vn[k][c][x] - snd.vp[k][c][x];
						const float dm = rec.m_new[c][x] - rec.m_prev[c][x];
						if (dn * dm != 0.f) {
							lin = false;
						}
						if (snd.vn[k][c][x] != snd.vn[0][c][x] || snd.vp[k][c][x] != snd.vp[0][c][x]) {
							shared = false;
						}
					}
				}
			}
			n_group = shared ? 1 : n_send;
			const int P = lin ? 2 : 3;
			n_rows = n_group * P;
			if (n_rows <= GAS_K2_MAX_ROWS) {
				streamed = true;
				path = PATH_STREAM;
				if (lin) {
					cflags |= CLS_LIN;
				}
				if (shared) {
					cflags |= CLS_SHARED;
				}
				for (int k = 0; k < n_group; k++) {
					for (int c = 0; c < C; c++) {
						for (int x = 0; x < 2; x++) {
							const float np = snd.vp[k][c][x], dn = snd.vn[k][c][x] - np;
							const float mp = rec.m_prev[c][x], dm = rec.m_new[c][x] - mp;
							rows[((k * P + 0) * C + c) * 2 + x] = np * mp;
							rows[((k * P + 1) * C + c) * 2 + x] = np * dm + dn * mp;
							if (!lin) {
								rows[((k * P + 2) * C + c) * 2 + x] = dn * dm;
							}
						}
					}
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
			}
		}
	}

	// class lookup once per distinct key per warp, then a warp-aggregated append
	unsigned long long key = 0ULL;
	if (path != PATH_NONE) {
		key = (unsigned long long)path | ((unsigned long long)mode << 2) | ((unsigned long long)cflags << 4) |
				((unsigned long long)n_send << 8) | ((unsigned long long)mask << 16);
	}
	const unsigned peers = __match_any_sync(0xffffffffu, key);
	const int leader = __ffs(peers) - 1;
	int cid = -1, base = 0;
	if (key != 0ULL && (int)lane == leader) {
		cid = class_find_or_insert(cls, key, plan.overflow);
		if (cid >= 0) {
			ClassInfo *ci = &cls[cid];
			ci->path = path;
			ci->mode = mode;
			ci->flags = cflags;
			ci->mask = mask;
			ci->n_send = n_send;
			ci->n_group = n_group;
			ci->n_rows = n_rows;
			base = atomicAdd(&ci->count, __popc(peers));
		}
	}
	cid = __shfl_sync(0xffffffffu, cid, leader);
	base = __shfl_sync(0xffffffffu, base, leader);
	if (key != 0ULL && cid >= 0) {
		const int pos = base + __popc(peers & ((1u << lane) - 1u));
		if (path == PATH_STREAM) {
			plan.k2_src[(size_t)cid * maxv + pos] = v.src_row;
			const int nf = n_rows * C * 2;
			float *dst = plan.k2_rows + (size_t)cid * maxv * GAS_K2_ROW_FLOATS + (size_t)pos * nf;
			for (int i = 0; i < nf; i++) {
				dst[i] = rows[i];
			}
		} else {
			plan.k3_list[(size_t)cid * maxv + pos] = j;
			plan.rec[j] = rec;
		}
	}

	// ---- the last CTA to finish advances the block counter (every CTA has read it by then) ----------------
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		const int ticket = atomicAdd(&t.blk[1], 1);
		if (ticket == (int)gridDim.x - 1) {
			t.blk[1] = 0;
			t.blk[0] = t.blk[0] + 1;
		}
	}
}

} // namespace

cudaError_t launch_prologue(gas_ctx *ctx, int n_voices, const gas_voice *d_voices, int src_rows, int frames, gas_frame *d_bus,
		gas_frame *d_peaks, cudaStream_t st) {
	const int bus_f4 = ctx->g.num_buses * ctx->g.channels * frames / 2;
	int work = ctx->inst_hwm > n_voices ? ctx->inst_hwm : n_voices;
	work = work > GAS_MAX_CLASSES ? work : GAS_MAX_CLASSES;
	int blocks = (work + 127) / 128;
	k_prologue<<<blocks, 128, 0, st>>>(ctx->t, ctx->g, ctx->plan, ctx->inst_hwm, n_voices, d_voices, src_rows, (float4 *)d_bus, bus_f4,
			(float2 *)d_peaks);
	ctx->launches++;
	return cudaGetLastError();
}
