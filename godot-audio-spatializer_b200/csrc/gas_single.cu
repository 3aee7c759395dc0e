// gas_single.cu — the reference's per-call virtuals on one voice (reference audio_spatializer.h:146,148 and the raw-pointer
// GDVIRTUAL mirror :103-112): process_frames (audio_spatializer_3d.cpp:491-552, audio_spatializer_effect.cpp:33-77) and
// mix_channel (audio_spatializer_3d.cpp:554-609), with the instance's current parameters and the voice's playback data,
// exactly as the reference runs them: out is overwritten, the playback data advance.
//
// The batched path never calls these (it does the same work for every voice of the block at once); they are the entry
// points a subclass that overrides one of the two virtuals calls for the built-in behaviour, and the way to run a single
// playback outside a mix step.  One warp, lane 0 = left, lane 1 = right: a 512-frame block is ~10 us of serial
// recurrence — the price of a per-call interface, which is why the mix path is not built on it.
// Compiled with -fmad=false: every operation rounds like the reference's scalar loop.
#include "gas_internal.h"
#include "gas_filter.cuh"

namespace {

struct Proc { // upstream AudioFilterSW::Processor
	float b0, b1, b2, a1, a2, ha1, ha2, hb1, hb2;
};

__device__ __forceinline__ float process_one(Proc &p, float x) {
	const float y = x * p.b0 + p.hb1 * p.b1 + p.hb2 * p.b2 + p.ha1 * p.a1 + p.ha2 * p.a2;
	p.ha2 = p.ha1;
	p.hb2 = p.hb1;
	p.hb1 = x;
	p.ha1 = y;
	return y;
}

__device__ __forceinline__ Proc proc_load(const gas_processor_state &s) {
	return Proc{ s.b0, s.b1, s.b2, s.a1, s.a2, s.ha1, s.ha2, s.hb1, s.hb2 };
}
__device__ __forceinline__ void proc_store(gas_processor_state &s, const Proc &p) {
	s.b0 = p.b0;
	s.b1 = p.b1;
	s.b2 = p.b2;
	s.a1 = p.a1;
	s.a2 = p.a2;
	s.ha1 = p.ha1;
	s.ha2 = p.ha2;
	s.hb1 = p.hb1;
	s.hb2 = p.hb2;
}

// channel < 0: process_frames; channel >= 0: mix_channel for that pair
__global__ void k_single_voice(DevTables t, GlobalCfg g, int q, int v, int channel, gas_frame *__restrict__ out, const gas_frame *__restrict__ src, int F) {
	const int side = threadIdx.x; // 0 = left, 1 = right
	if (side > 1) {
		return;
	}
	const gas_params *prm = &t.inst_params[q];
	const int imode = t.inst_mode[q];
	const int mode = imode & 0xff;
	const float *srcf = reinterpret_cast<const float *>(src);
	float *outf = reinterpret_cast<float *>(out);
	float *vprev = t.vs_prev + (size_t)v * 8;
	if (channel < 0 && mode == MODE_E) {
		// AudioSpatializerInstanceEffect::process_frames: the chain of AudioEffectFilter instances, coefficients not interpolated
		const gas_effect_chain *fx = &t.inst_fx[q];
		const int binding = (imode >> 8) - 1;
		int n_fx = fx->n_effects;
		n_fx = n_fx < 0 ? 0 : (n_fx > GAS_MAX_EFFECTS ? GAS_MAX_EFFECTS : n_fx);
		for (int i = 0; i < F; i++) {
			outf[i * 2 + side] = srcf[i * 2 + side];
		}
		for (int j = 0; j < n_fx; j++) {
			gas_effect ef = fx->effects[j];
			if (binding == j) {
				ef.gain = prm->linear_attenuation; // example _process_effects (gd_spatializer_instance.gd:125-127)
			}
			const int stages = ef.stages < 1 ? 1 : (ef.stages > GAS_MAX_FILTER_STAGES ? GAS_MAX_FILTER_STAGES : ef.stages);
			float cf[5];
			prepare_coefficients(ef.mode, ef.cutoff_hz, ef.resonance, ef.gain, stages, g.mix_rate, cf);
			float *hist = t.vs_fx + (((size_t)v * GAS_MAX_EFFECTS + j) * 2 + side) * GAS_MAX_FILTER_STAGES * 4;
			Proc p[GAS_MAX_FILTER_STAGES];
			for (int s = 0; s < stages; s++) {
				p[s] = Proc{ cf[0], cf[1], cf[2], cf[3], cf[4], hist[s * 4 + 0], hist[s * 4 + 1], hist[s * 4 + 2], hist[s * 4 + 3] };
			}
			for (int i = 0; i < F; i++) {
				float f = outf[i * 2 + side];
				for (int s = 0; s < stages; s++) {
					f = process_one(p[s], f);
				}
				outf[i * 2 + side] = f;
			}
			for (int s = 0; s < stages; s++) {
				hist[s * 4 + 0] = p[s].ha1;
				hist[s * 4 + 1] = p[s].ha2;
				hist[s * 4 + 2] = p[s].hb1;
				hist[s * 4 + 3] = p[s].hb2;
			}
		}
		return;
	}
	const int c = channel < 0 ? 0 : channel;
	const float vs = vprev[c * 2 + side], vs_other = vprev[c * 2 + (side ^ 1)]; // :500 / :564
	const float vf = prm->mix_volumes[c][side];                                 // :565
	const float highshelf_gain = prm->linear_attenuation;
	const bool ramp = channel >= 0;
	if ((double)highshelf_gain >= 0.001) { // :503 / :568
		float target[5];
		prepare_coefficients(GAS_FILTER_HIGHSHELF, prm->attenuation_filter_cutoff_hz, 1.0f, highshelf_gain, 1, g.mix_rate, target);
		gas_processor_state *ps = t.vs_proc + (size_t)v * 8 + c * 2 + side; // index pair * 2 + (left ? 0 : 1), :887-894
		Proc p = proc_load(*ps);
		if (vs == 0.f && vs_other == 0.f) { // is_just_started: clear history (:518-521, :583-586)
			p.ha1 = p.ha2 = p.hb1 = p.hb2 = 0.f;
		}
		const float inc[5] = { (target[0] - p.b0) / F, (target[1] - p.b1) / F, (target[2] - p.b2) / F, (target[3] - p.a1) / F, (target[4] - p.a2) / F };
		for (int i = 0; i < F; i++) {
			float x = srcf[i * 2 + side];
			if (ramp) {
				const float tt = (float)i / F;                  // :591
				x = (vf * tt + (1 - tt) * vs) * x;              // :592-593
			}
			outf[i * 2 + side] = process_one(p, x);            // process_one_interp ...
			p.b0 += inc[0];
			p.b1 += inc[1];
			p.b2 += inc[2];
			p.a1 += inc[3];
			p.a2 += inc[4];
		}
		proc_store(*ps, p);
	} else {
		for (int i = 0; i < F; i++) {
			float x = srcf[i * 2 + side];
			if (ramp) {
				const float tt = (float)i / F;
				x = (vf * tt + (1 - tt) * vs) * x; // :603
			}
			outf[i * 2 + side] = x;
		}
	}
	// previous volume bookkeeping: mix_channel keeps volumes[channel] (:608); process_frames keeps the pair holding the
	// largest component, scanned c0.L, c0.R, c1.L, ... (:537-551)
	__syncwarp(0x3u);
	if (ramp) {
		vprev[c * 2 + side] = vf;
	} else {
		float max_volume = 0.f;
		int max_index = 0;
		for (int i = 0; i < GAS_MAX_CHANNELS_PER_BUS; i++) {
			if (prm->mix_volumes[i][0] > max_volume) {
				max_volume = prm->mix_volumes[i][0];
				max_index = i;
			}
			if (prm->mix_volumes[i][1] > max_volume) {
				max_volume = prm->mix_volumes[i][1];
				max_index = i;
			}
		}
		vprev[side] = prm->mix_volumes[max_index][side];
	}
}

} // namespace

cudaError_t launch_single_voice(gas_ctx *ctx, int instance, int voice, int channel, gas_frame *d_out, const gas_frame *d_src, int frames, cudaStream_t st) {
	k_single_voice<<<1, 32, 0, st>>>(ctx->t, ctx->g, instance, voice, channel, d_out, d_src, frames);
	ctx->launches++;
	return cudaGetLastError();
}
