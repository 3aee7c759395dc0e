// gas_resample.cu — the resampler in front of the path: device-resident PCM sources and what
// `playback->stream_playback->mix(&buf[LOOKAHEAD_BUFFER_SIZE], pitch_scale, p_buffer_size)` (reference
// audio_spatializer.cpp:375-378) produces for them, for every voice of a block at once.
//
// Upstream AudioStreamPlaybackResampled::mix (Godot 4.x, as recalled — the engine is not in the reference tree) walks a
// 16.16 fixed-point offset through a 128-frame internal buffer with a 4-frame history and interpolates cubically between
// the frames two and one positions back.  Seen from the stream that is a closed form: output i of a call reads
//     S[g - 3 .. g],  g = (P + i * increment) >> 16,  mu = frac / 65536
// where P is the playback's fixed-point position since begin_resample and S the PCM stream (0 before the start, 0 after the
// end, wrapped when looping) — no recurrence, so a block is frame-parallel.  The end-of-stream rule is upstream's, quirk
// included: a refill that returns fewer than 128 frames sets internal_buffer_end to that count (never reset once the stream has
// stopped), and the first output whose buffer index (4 + g % 128) reaches it ends the count of good frames.
//
// The rows go into the stream form of the mix (gas_mix_block_stream_device: lookahead splice, end fade, deactivation), so a
// block needs no source frames from the host: only the emitters travel.  Compiled with -fmad=false: the interpolation
// rounds like the scalar loop.
#include "gas_internal.h"

namespace {

constexpr int kThreads = 128;
constexpr int kFpBits = 16;
constexpr int kBufLen = 128; // INTERNAL_BUFFER_LEN
constexpr int kHistory = 4;  // CUBIC_INTERP_HISTORY

__device__ __forceinline__ gas_frame stream_at(const SourceDesc &s, long long start, long long x) {
	// frame x of the stream as begin_resample / _mix_internal deliver it: history before the start is cleared
	gas_frame z;
	z.l = z.r = 0.f;
	if (x < 0) {
		return z;
	}
	long long a = start + x;
	if (a >= s.n_frames) {
		if (!s.loop) {
			return z;
		}
		a %= s.n_frames;
	}
	return s.pcm[a];
}

__global__ void __launch_bounds__(kThreads) k_resample(DevTables t, GlobalCfg g, const SourceDesc *__restrict__ sources, int max_sources, int n_voices,
		const gas_voice *__restrict__ voices, int frames, gas_frame *__restrict__ rows, int row_stride, int src_rows, int32_t *__restrict__ mixed_out) {
	__shared__ int s_min;
	const int j = blockIdx.x;
	if (j >= n_voices) {
		return;
	}
	const gas_voice v = voices[j];
	const bool valid = v.voice >= 0 && v.voice < g.max_voices && v.instance >= 0 && v.instance < g.max_instances;
	const int src = valid ? t.vs_src[v.voice] : -1;
	const bool have = src >= 0 && src < max_sources && v.src_row >= 0 && v.src_row < src_rows;
	if (threadIdx.x == 0) {
		s_min = frames;
	}
	__syncthreads();
	if (!have) {
		if (threadIdx.x == 0 && mixed_out) {
			mixed_out[j] = 0; // nothing delivered: the stream form ends such a voice
		}
		if (v.src_row >= 0 && v.src_row < src_rows) {
			gas_frame z;
			z.l = z.r = 0.f;
			for (int i = threadIdx.x; i < frames; i += kThreads) {
				rows[(size_t)v.src_row * row_stride + i] = z;
			}
		}
		return;
	}
	const SourceDesc s = sources[src];
	const float pitch = t.inst_params[v.instance].pitch_scale; // parameters->get_pitch_scale(), :375
	// uint64_t mix_increment = uint64_t(((get_stream_sampling_rate() * p_rate_scale * playback_speed_scale) / double(target_rate)) * double(FP_LEN));
	const float rate = s.sample_rate * pitch * 1.0f;
	const double incd = ((double)rate / (double)g.mix_rate) * 65536.0;
	const unsigned long long inc = incd > 0.0 ? (unsigned long long)incd : 0ULL;
	const unsigned long long p0 = t.vs_pos[v.voice];
	const long long start = t.vs_start[v.voice];
	// end of stream (non-looping): the refill of buffer k_end returns end_val < 128 frames
	const long long n_rel = (long long)s.n_frames - start;
	const long long k_end = n_rel >> 7;
	const unsigned end_val = (unsigned)(n_rel - (k_end << 7));
	int first_bad = frames;
	gas_frame *row = rows + (size_t)v.src_row * row_stride;
	for (int i = threadIdx.x; i < frames; i += kThreads) {
		const unsigned long long p = p0 + (unsigned long long)i * inc;
		const long long gi = (long long)(p >> kFpBits);
		const float mu = (float)(unsigned)(p & 0xffffu) / 65536.0f;
		const gas_frame y0 = stream_at(s, start, gi - 3), y1 = stream_at(s, start, gi - 2), y2 = stream_at(s, start, gi - 1), y3 = stream_at(s, start, gi);
		if (!s.loop && (gi >> 7) >= k_end && (unsigned)kHistory + (unsigned)(gi & (kBufLen - 1)) >= end_val) {
			first_bad = min(first_bad, i);
		}
		const float mu2 = mu * mu;
		const float h11 = mu2 * (mu - 1.f);
		const float z = mu2 - h11;
		const float h01 = z - h11;
		const float h10 = mu - z;
		gas_frame o;
		o.l = (y1.l + (y2.l - y1.l) * h01) + ((y2.l - y0.l) * h10 + (y3.l - y1.l) * h11) * 0.5f;
		o.r = (y1.r + (y2.r - y1.r) * h01) + ((y2.r - y0.r) * h10 + (y3.r - y1.r) * h11) * 0.5f;
		row[i] = o;
	}
	if (first_bad < frames) {
		atomicMin(&s_min, first_bad);
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		if (mixed_out) {
			mixed_out[j] = s_min;
		}
		t.vs_pos[v.voice] = p0 + (unsigned long long)frames * inc;
	}
}

__global__ void k_voice_play(DevTables t, int n, const int32_t *__restrict__ voices, const int32_t *__restrict__ sources, const int32_t *__restrict__ starts) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) {
		return;
	}
	const int v = voices[i];
	t.vs_src[v] = sources[i];
	t.vs_start[v] = starts ? starts[i] : 0;
	t.vs_pos[v] = 0ULL; // begin_resample(): mix_offset = 0, history cleared
}

} // namespace

cudaError_t launch_resample(gas_ctx *ctx, int n_voices, const gas_voice *d_voices, int frames, gas_frame *d_rows, int row_stride, int src_rows,
		int32_t *d_mixed, cudaStream_t st) {
	if (n_voices <= 0) {
		return cudaSuccess;
	}
	k_resample<<<n_voices, kThreads, 0, st>>>(ctx->t, ctx->g, ctx->d_sources, ctx->max_sources, n_voices, d_voices, frames, d_rows, row_stride, src_rows,
			d_mixed);
	ctx->launches++;
	return cudaGetLastError();
}

cudaError_t launch_voice_play(gas_ctx *ctx, int n, const int32_t *d_voices, const int32_t *d_sources, const int32_t *d_starts, cudaStream_t st) {
	if (n <= 0) {
		return cudaSuccess;
	}
	k_voice_play<<<(n + 127) / 128, 128, 0, st>>>(ctx->t, n, d_voices, d_sources, d_starts);
	ctx->launches++;
	return cudaGetLastError();
}
