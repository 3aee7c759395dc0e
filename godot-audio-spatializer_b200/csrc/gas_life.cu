// gas_life.cu — voice lifecycle around the per-voice mix: what AudioSpatializerInstance::_mix_from_playback_list does
// before and after it calls process_frames / mix_channel (reference audio_spatializer.cpp:353-408, :464-469).
//
// Per voice slot the context keeps the rest of SpatialPlaybackListNode (reference audio_spatializer.h:55-66): `active`,
// `has_frames` and the 64-frame lookahead.  The stream form of the mix (gas_mix_block_stream*) takes, per voice, the frames
// AudioStreamPlayback::mix returned this block and
//   k_life_stage  skips inactive voices (:355), splices the lookahead in front of the new frames (:369-378), keeps the last
//                 64 frames as the next lookahead (:401-403) or — when the playback came up short — fades the last 64 valid
//                 frames with 0.96^k * (64 - k) / 64, zeroes the rest and clears has_frames (:380-398); a voice without
//                 frames is mixed with silence (:405-408).  Output: a staged block per voice + a rewritten voice list, which
//                 then go through the ordinary block path (prologue, streaming / voice-parallel kernels).
//   k_life_post   deactivates a voice without frames whose block peak is at or below the instance's
//                 playback_disable_threshold_db (:464-469) and reports every voice's state.
// One warp per voice in k_life_stage: the copy is coalesced 8-byte elements; the fade is evaluated per element from
// the closed form of the reference's running product (same float operations in the same order, see fade_factor).
#include "gas_internal.h"

namespace {

constexpr int L = GAS_LOOKAHEAD_BUFFER_SIZE;

// The reference's running product  coef *= 0.96f  after k + 1 steps, then  coef * (64 - k) / 64  (:388-389): the k-th
// faded frame.  Evaluated by repeated multiplication so that every intermediate rounds exactly like the serial loop.
__device__ __forceinline__ float fade_factor(int k) {
	const float base = (float)0.96;
	float coef = 1.0f;
	for (int i = 0; i <= k; i++) {
		coef = __fmul_rn(coef, base);
	}
	const float size = (float)L;
	return __fdiv_rn(__fmul_rn(coef, size - (float)k), size);
}

__global__ void __launch_bounds__(256) k_life_stage(DevTables t, int n_voices, const gas_voice *__restrict__ voices, const int32_t *__restrict__ mixed_frames,
		const gas_frame *__restrict__ src, int src_rows, int src_stride, int F, gas_voice *__restrict__ out_voices, gas_frame *__restrict__ stage,
		int stage_stride) {
	const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (warp >= n_voices) {
		return;
	}
	gas_voice v = voices[warp];
	const bool valid = v.voice >= 0 && v.voice < t.max_voices;
	uint32_t life = valid ? t.vs_life[v.voice] : 0u;
	gas_voice o = v;
	if (!(life & 1u)) { // inactive: the reference does not even look at the playback (:355)
		o.voice = -1;
		o.instance = -1;
		o.src_row = -1;
		if (lane == 0) {
			out_voices[warp] = o;
		}
		return;
	}
	// The reference tracks the block peak of every voice (:419) but reads it only once has_frames is clear (:464): only the voices that
	// end in this block or are already in their tail ask for one.  Everything else keeps the streaming path.
	if (life & 2u) {
		const bool has_row = v.src_row >= 0 && v.src_row < src_rows;
		int mixed = has_row ? mixed_frames[warp] : 0;
		mixed = mixed < 0 ? 0 : (mixed > F ? F : mixed);
		const float2 *row = has_row ? reinterpret_cast<const float2 *>(src + (size_t)v.src_row * src_stride) : nullptr;
		float2 *look = reinterpret_cast<float2 *>(t.vs_look + (size_t)v.voice * L);
		float2 *dst = reinterpret_cast<float2 *>(stage + (size_t)warp * stage_stride);
		const int fade_limit = mixed + L;
		// buf[idx], idx in [0, F): lookahead for idx < 64, else the new frame idx - 64 (zero beyond what was delivered)
		for (int idx = lane; idx < F; idx += 32) {
			float2 x;
			if (idx < L) {
				x = look[idx];
			} else {
				const int k = idx - L;
				x = k < mixed ? __ldg(row + k) : make_float2(0.f, 0.f);
			}
			if (mixed != F && idx >= mixed) {
				if (idx < fade_limit) {
					const float f = fade_factor(idx - mixed);
					x.x = __fmul_rn(x.x, f);
					x.y = __fmul_rn(x.y, f);
				} else {
					x.x = __fmul_rn(x.x, 0.0f);
					x.y = __fmul_rn(x.y, 0.0f);
				}
			}
			dst[idx] = x;
		}
		__syncwarp();
		if (mixed == F) { // the last 64 frames of buf[0 .. F + 64) become the next lookahead (:401-403)
			for (int k = lane; k < L; k += 32) {
				const int idx = F + k; // buf index
				float2 x;
				if (idx < L) {
					x = look[idx]; // only when F < 64: part of the old lookahead moves up
				} else {
					x = __ldg(row + (idx - L));
				}
				__syncwarp();
				// (F < 64 would read and write overlapping lookahead elements; frames are even and >= 2, blocks shorter than
				// the lookahead keep the serial semantics through the register copy above)
				look[k] = x;
			}
		} else {
			life &= ~2u; // no more frames to mix (:397)
			o.flags |= GAS_VOICE_WANT_PEAK;
		}
		o.src_row = warp;
	} else {
		o.src_row = -1; // zero-filled playback buffer (:405-408)
		o.flags |= GAS_VOICE_WANT_PEAK;
	}
	if (lane == 0) {
		t.vs_life[v.voice] = life;
		out_voices[warp] = o;
	}
}

__global__ void k_life_post(DevTables t, int n_voices, const gas_voice *__restrict__ voices, const float2 *__restrict__ peaks,
		int32_t *__restrict__ status_out) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_voices) {
		return;
	}
	const gas_voice v = voices[i];
	int status = 0;
	if (v.voice >= 0 && v.voice < t.max_voices) {
		uint32_t life = t.vs_life[v.voice];
		if ((life & 1u) && !(life & 2u) && v.instance >= 0 && v.instance < t.max_instances) {
			const float2 pk = peaks[i];
			const float m = pk.y > pk.x ? pk.y : pk.x; // MAX(peak.right, peak.left)
			if (m <= t.inst_threshold[v.instance]) {    // :465
				life &= ~1u;
				t.vs_life[v.voice] = life;
			}
		}
		status = (int)(life & 3u);
	}
	if (status_out) {
		status_out[i] = status;
	}
}

__global__ void k_threshold_set(DevTables t, int n, const int32_t *__restrict__ ids, const float *__restrict__ lin) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) {
		t.inst_threshold[ids[i]] = lin[i];
	}
}

__global__ void k_life_export(DevTables t, int n, const int32_t *__restrict__ ids, gas_voice_life *__restrict__ out) {
	const int i = blockIdx.x;
	if (i >= n) {
		return;
	}
	const int v = ids[i];
	for (int k = threadIdx.x; k < L; k += blockDim.x) {
		out[i].lookahead[k] = t.vs_look[(size_t)v * L + k];
	}
	if (threadIdx.x == 0) {
		out[i].flags = t.vs_life[v];
	}
}

__global__ void k_life_import(DevTables t, int n, const int32_t *__restrict__ ids, const gas_voice_life *__restrict__ in) {
	const int i = blockIdx.x;
	if (i >= n) {
		return;
	}
	const int v = ids[i];
	for (int k = threadIdx.x; k < L; k += blockDim.x) {
		t.vs_look[(size_t)v * L + k] = in[i].lookahead[k];
	}
	if (threadIdx.x == 0) {
		t.vs_life[v] = in[i].flags & 3u;
	}
}

} // namespace

cudaError_t launch_life_stage(gas_ctx *ctx, int n_voices, const gas_voice *d_voices, const int32_t *d_mixed, const gas_frame *d_src, int src_rows,
		int src_stride, int frames, gas_voice *d_out_voices, gas_frame *d_stage, int stage_stride, cudaStream_t st) {
	if (n_voices <= 0) {
		return cudaSuccess;
	}
	const int threads = 256;
	const int blocks = (int)(((long long)n_voices * 32 + threads - 1) / threads);
	k_life_stage<<<blocks, threads, 0, st>>>(ctx->t, n_voices, d_voices, d_mixed, d_src, src_rows, src_stride, frames, d_out_voices, d_stage, stage_stride);
	ctx->launches++;
	return cudaGetLastError();
}

cudaError_t launch_life_post(gas_ctx *ctx, int n_voices, const gas_voice *d_voices, const gas_frame *d_peaks, int32_t *d_status, cudaStream_t st) {
	if (n_voices <= 0) {
		return cudaSuccess;
	}
	k_life_post<<<(n_voices + 127) / 128, 128, 0, st>>>(ctx->t, n_voices, d_voices, (const float2 *)d_peaks, d_status);
	ctx->launches++;
	return cudaGetLastError();
}

cudaError_t launch_threshold_set(gas_ctx *ctx, int n, const int32_t *d_ids, const float *d_lin, cudaStream_t st) {
	if (n <= 0) {
		return cudaSuccess;
	}
	k_threshold_set<<<(n + 127) / 128, 128, 0, st>>>(ctx->t, n, d_ids, d_lin);
	ctx->launches++;
	return cudaGetLastError();
}

cudaError_t launch_life_export(gas_ctx *ctx, int n, const int32_t *d_ids, gas_voice_life *d_out, cudaStream_t st) {
	if (n <= 0) {
		return cudaSuccess;
	}
	k_life_export<<<n, 64, 0, st>>>(ctx->t, n, d_ids, d_out);
	ctx->launches++;
	return cudaGetLastError();
}

cudaError_t launch_life_import(gas_ctx *ctx, int n, const int32_t *d_ids, const gas_voice_life *d_in, cudaStream_t st) {
	if (n <= 0) {
		return cudaSuccess;
	}
	k_life_import<<<n, 64, 0, st>>>(ctx->t, n, d_ids, d_in);
	ctx->launches++;
	return cudaGetLastError();
}
