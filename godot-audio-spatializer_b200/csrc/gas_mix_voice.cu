// gas_mix_voice.cu — K3: the voice-parallel mix kernel (sm_100a).
//
// Voices whose block contains a per-sample recurrence — the attenuation high-shelf biquad of
// process_frames / mix_channel (reference audio_spatializer_3d.cpp:503-529, :568-597, upstream
// AudioFilterSW::Processor) or an AudioSpatializerEffect filter chain (reference
// audio_spatializer_effect.cpp:33-77) — and voices that must report a block peak
// (reference audio_spatializer.cpp:419-461) run here, serial in time with the filter state in registers for
// the whole block.  The parallel axis is the *stream*, not the voice.  Two forms:
//   filter-tile path (Mode A and effect chains of the ordinary classes): the whole CTA works on a batch of voices; one
//        lane per (voice, side) runs nothing but the biquads and hands the processed stream to a frame-parallel
//        contraction through shared memory, 64 frames at a time (see "filter-tile path" below);
//   per-warp path (Mode B: one lane per (voice, pair, side), 4 voices per warp; generic classes: one voice per warp):
//        after the per-stream work the AudioServer ramp (upstream _mix_step_for_channel) is applied per lane, the
//        voices of the warp are summed with shuffles and the result is added to the CTA's accumulation tile with
//        explicit red.shared.  Sends are processed two at a time; a voice with more than two sends (bus transitions) is
//        run in several passes from the same initial state — every pass recomputes bit-identical samples, only the last
//        one stores state and peaks.
// Both add into a CTA-wide accumulation tile in shared memory (the whole bus layout of the block); the CTA adds its tile to
// the bus buffers once, with one vector reduction per 16 bytes.
#include "gas_internal.h"

namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kMaxTileBytes = 96 * 1024; // CTA-wide accumulation tile (whole bus layout of the block) when it fits
constexpr unsigned kFull = 0xffffffffu;

// add into the CTA's accumulation tile (explicit shared-space reduction: a generic-address atomicAdd on shared
// memory is an order of magnitude slower) or, without a tile, into the bus buffers
__device__ __forceinline__ void acc_add(float *bus, gas_smem_addr tile_s, size_t o, float v) {
	if (tile_s) {
		gas_red_shared_add_f32(tile_s + (gas_smem_addr)o * 4u, v);
	} else {
		atomicAdd(bus + o, v);
	}
}

struct Biquad {
	float ha1, ha2, hb1, hb2;
};

// upstream AudioFilterSW::Processor::process_one: y = x*b0 + hb1*b1 + hb2*b2 + ha1*a1 + ha2*a2
__device__ __forceinline__ float biquad_step(Biquad &h, float x, float b0, float b1, float b2, float a1, float a2) {
	float y = x * b0 + h.hb1 * b1 + h.hb2 * b2 + h.ha1 * a1 + h.ha2 * a2;
	h.ha2 = h.ha1;
	h.hb2 = h.hb1;
	h.hb1 = x;
	h.ha1 = y;
	return y;
}

struct ChunkArgs {
	const VoiceRec *rec;
	const InstSends *sends; // by call-order index
	const int2 *list;     // class list {call-order index, source row}
	int count;            // voices in the class
	int chunk;            // which 32-voice chunk
	uint32_t cls_flags;
	uint32_t mask;
	int n_send;
};

// Transposing reduction over the high lane bits (16, 8, ... for LEVELS levels) of a batch of N per-lane values.
// At every level the lanes with the bit set keep the upper half of what is left and receive the lower lanes'
// upper half (and vice versa), so all shuffles of a level are independent of each other; once a single value is
// left the remaining levels are plain butterfly sums.  On return v[0 .. max(1, N >> LEVELS)) hold totals over the
// reduced lanes of the original values base .. base + count, and `owner` tells the one lane of every group of
// identical results that should use them.
template <int W, int LVL, int LEVELS, int N>
struct ReduceHi {
	static __device__ __forceinline__ void run(float (&v)[N], int lane, int &base, bool &owner) {
		constexpr int m = 16 >> LVL;
		if (W > 1) {
			constexpr int H = W / 2 > 0 ? W / 2 : 1;
			const bool upper = (lane & m) != 0;
#pragma unroll
			for (int i = 0; i < H; i++) {
				const float send = upper ? v[i] : v[i + H];
				const float keep = upper ? v[i + H] : v[i];
				v[i] = keep + __shfl_xor_sync(kFull, send, m);
			}
			base += upper ? H : 0;
			ReduceHi<H, LVL + 1, LEVELS, N>::run(v, lane, base, owner);
		} else {
			v[0] += __shfl_xor_sync(kFull, v[0], m);
			owner = owner && (lane & m) == 0;
			ReduceHi<1, LVL + 1, LEVELS, N>::run(v, lane, base, owner);
		}
	}
};
template <int W, int LEVELS, int N>
struct ReduceHi<W, LEVELS, LEVELS, N> {
	static __device__ __forceinline__ void run(float (&)[N], int, int &, bool &) {}
};
template <int N, int LEVELS>
__device__ __forceinline__ void reduce_hi(float (&v)[N], int lane, int &base, bool &owner) {
	base = 0;
	owner = true;
	ReduceHi<N, 0, LEVELS, N>::run(v, lane, base, owner);
}

// Mode B (with or without the attenuation filter: up to 2C biquads per voice).  A lane is one (voice, pair, side) stream — 8
// lanes per voice, 4 voices per warp — so a voice's biquads run side by side instead of one after the other in a
// single lane, and 8x more warps hide the recurrence latency.  Frames are fetched 8 at a time (one float2 per lane
// of the voice = 64 contiguous bytes) and handed round with shuffles; t = i / F is computed once per frame by the
// lane that fetched it.  The 4 voices of the warp are summed with two shuffles per send before the add into the
// CTA's accumulation tile (or the bus buffers).
template <int C, int NS>
__device__ void voice_pass_b(const DevTables &t, const ChunkArgs &a, int send0, bool last_pass, const gas_frame *__restrict__ src,
		int src_stride, int F, float *__restrict__ bus, gas_smem_addr tile, float2 *__restrict__ peaks) {
	const int lane = threadIdx.x & 31;
	const int g = lane >> 3, l = lane & 7, c = l >> 1, side = l & 1;
	const int gbase = lane & 24;
	const int pos = a.chunk * 4 + g;
	const bool has_voice = pos < a.count;
	const bool active = has_voice && c < C;
	const int j = has_voice ? a.list[pos].x : -1;
	const VoiceRec *rec = &a.rec[has_voice ? j : 0];
	const int voice = has_voice ? rec->voice : 0;
	const int src_row = has_voice ? rec->src_row : -1;
	const uint32_t flags = has_voice ? rec->flags : 0u;
	const float m_prev = active ? rec->m_prev[c][side] : 0.f;
	const float m_new = active ? rec->m_new[c][side] : 0.f;
	float target[5];
#pragma unroll
	for (int q = 0; q < 5; q++) {
		target[q] = active ? rec->target[q] : 0.f;
	}
	float np[NS], nn[NS];
	int bus_of[NS];
	{
		const InstSends *snd = &a.sends[has_voice ? j : 0];
		uint32_t m = a.mask;
		for (int s = 0; s < send0; s++) {
			m &= m - 1;
		}
#pragma unroll
		for (int s = 0; s < NS; s++) {
			bus_of[s] = m ? (__ffs(m) - 1) : 0;
			m &= m - 1;
			const bool ok = active && (send0 + s) < a.n_send;
			np[s] = ok ? snd->vp[send0 + s][c][side] : 0.f;
			nn[s] = ok ? snd->vn[send0 + s][c][side] : 0.f;
		}
	}
	const bool filt = (a.cls_flags & CLS_FILT) != 0;
	gas_processor_state *ps = t.vs_proc + (size_t)voice * 8 + l; // processor index pair * 2 + (left ? 0 : 1)
	gas_processor_state st{};
	if (active && filt) {
		st = *ps;
	}
	const bool clear = (flags >> (8 + c)) & 1u; // is_just_started => clear history (:583-586)
	Biquad h;
	h.ha1 = clear ? 0.f : st.ha1;
	h.ha2 = clear ? 0.f : st.ha2;
	h.hb1 = clear ? 0.f : st.hb1;
	h.hb2 = clear ? 0.f : st.hb2;
	float cf[5] = { st.b0, st.b1, st.b2, st.a1, st.a2 }, inc[5];
#pragma unroll
	for (int q = 0; q < 5; q++) { // update_coeffs(F): per-sample increment towards the target
		inc[q] = (target[q] - cf[q]) / (float)F;
	}
	float pk = 0.f;
	const float2 *row = src_row >= 0 ? reinterpret_cast<const float2 *>(src + (size_t)src_row * src_stride) : nullptr;
	float2 nx = (row && l < F) ? __ldg(row + l) : make_float2(0.f, 0.f); // next trip's frames, fetched a trip ahead
	for (int i0 = 0; i0 < F; i0 += 8) {
		const int mi = i0 + l;
		const float2 mx = nx;
		nx = (row && mi + 8 < F) ? __ldg(row + mi + 8) : make_float2(0.f, 0.f);
		const float mt = (float)mi / (float)F; // :591
		float v[8 * NS]; // [frame][send]
#pragma unroll
		for (int k = 0; k < 8; k++) {
			const float xl = __shfl_sync(kFull, mx.x, gbase + k);
			const float xr = __shfl_sync(kFull, mx.y, gbase + k);
			const float tt = __shfl_sync(kFull, mt, gbase + k);
			const float omt = 1.0f - tt;
			const float vol = m_new * tt + omt * m_prev; // :592
			float y = vol * (side ? xr : xl);             // :593
			if (i0 + k < F) {
				if (filt) {
					y = biquad_step(h, y, cf[0], cf[1], cf[2], cf[3], cf[4]); // :594-595
#pragma unroll
					for (int q = 0; q < 5; q++) { // process_one_interp: coeffs += incr
						cf[q] += inc[q];
					}
				}
				pk = fmaxf(pk, fabsf(y));
			} else {
				y = 0.f;
			}
#pragma unroll
			for (int s = 0; s < NS; s++) {
				v[k * NS + s] = (nn[s] * tt + omt * np[s]) * y; // AudioServer ramp of this send (upstream _mix_step_for_channel)
			}
		}
		// the 4 voices of the warp: one transposing reduction over lane bits 4 and 3 for the whole trip
		int base;
		bool owner;
		reduce_hi<8 * NS, 2>(v, lane, base, owner);
		if (owner && c < C) {
#pragma unroll
			for (int r = 0; r < (8 * NS) / 4; r++) {
				const int idx = base + r, k = idx / NS, s = idx % NS;
				if ((send0 + s) < a.n_send && i0 + k < F) {
					acc_add(bus, tile, ((size_t)(bus_of[s] * C + c) * F + i0 + k) * 2 + side, v[r]);
				}
			}
		}
	}
	if (last_pass) {
		if (active && filt) {
			st.b0 = cf[0];
			st.b1 = cf[1];
			st.b2 = cf[2];
			st.a1 = cf[3];
			st.a2 = cf[4];
			st.ha1 = h.ha1;
			st.ha2 = h.ha2;
			st.hb1 = h.hb1;
			st.hb2 = h.hb2;
			*ps = st;
		}
		// block peak: max over the voice's pairs, per side (audio_spatializer.cpp:436-443)
		pk = fmaxf(pk, __shfl_xor_sync(kFull, pk, 2));
		pk = fmaxf(pk, __shfl_xor_sync(kFull, pk, 4));
		if (has_voice && (flags & GAS_VOICE_WANT_PEAK) && peaks && l < 2) {
			reinterpret_cast<float *>(peaks + j)[l] = pk;
		}
	}
}

template <int C>
__device__ void voice_chunk_b(const DevTables &t, const ChunkArgs &a, const gas_frame *__restrict__ src, int src_stride, int F,
		float *__restrict__ bus, gas_smem_addr tile, float2 *__restrict__ peaks) {
	const int n = a.n_send;
	if (n == 0) {
		voice_pass_b<C, 1>(t, a, 0, true, src, src_stride, F, bus, tile, peaks); // state / peak only: every send test fails
		return;
	}
	for (int s0 = 0; s0 < n; s0 += 2) {
		const bool last = s0 + 2 >= n;
		if (n - s0 >= 2) {
			voice_pass_b<C, 2>(t, a, s0, last, src, src_stride, F, bus, tile, peaks);
		} else {
			voice_pass_b<C, 1>(t, a, s0, last, src, src_stride, F, bus, tile, peaks);
		}
	}
}

template <int N>
struct Pow2Ceil8 {
	static constexpr int value = N <= 1 ? 1 : (N <= 2 ? 2 : (N <= 4 ? 4 : 8));
};

// Mode A (one high-shelf pair per voice, in front of every ramp) and AudioSpatializerEffect chains (cascaded
// constant-coefficient biquads): one lane per (voice, side), 16 voices per warp.  Each lane produces one processed
// stream y and NS * C weighted contributions per frame; the contributions of 8 frames are batched and the 16 voices
// summed with one transposing shuffle reduction over lane bits 4..1 (bit 0 is the side and is not reduced).
template <int MODE, int C, int NS>
__device__ void voice_pass_s(const DevTables &t, const ChunkArgs &a, int send0, bool emit, bool last_pass, const gas_frame *__restrict__ src,
		int src_stride, int F, float *__restrict__ bus, gas_smem_addr tile, float2 *__restrict__ peaks) {
	const int lane = threadIdx.x & 31;
	const int side = lane & 1;
	const int pos = a.chunk * 16 + (lane >> 1);
	const bool active = pos < a.count;
	const int j = active ? a.list[pos].x : -1;
	const VoiceRec *rec = &a.rec[active ? j : 0];
	const int voice = active ? rec->voice : 0;
	const int src_row = active ? rec->src_row : -1;
	const uint32_t flags = active ? rec->flags : 0u;
	const bool filt = MODE == MODE_A && (a.cls_flags & CLS_FILT) != 0;
	const int n_fx = (MODE == MODE_E && active) ? rec->n_fx : 0;

	float np[NS][C], nn[NS][C];
	int bus_of[NS];
	{
		const InstSends *snd = &a.sends[active ? j : 0];
		uint32_t m = a.mask;
		for (int s = 0; s < send0; s++) {
			m &= m - 1;
		}
#pragma unroll
		for (int s = 0; s < NS; s++) {
			bus_of[s] = m ? (__ffs(m) - 1) : 0;
			m &= m - 1;
			const bool ok = active && emit && (send0 + s) < a.n_send;
#pragma unroll
			for (int c = 0; c < C; c++) {
				np[s][c] = ok ? snd->vp[send0 + s][c][side] : 0.f;
				nn[s][c] = ok ? snd->vn[send0 + s][c][side] : 0.f;
			}
		}
	}
	// MODE_A: the interpolated high-shelf processor of this side (index 0 left, 1 right, :524-529)
	gas_processor_state *ps = t.vs_proc + (size_t)voice * 8 + side;
	gas_processor_state st{};
	if (active && filt) {
		st = *ps;
	}
	const bool clear = (flags >> 8) & 1u; // is_just_started => clear history (:518-521)
	Biquad h;
	h.ha1 = clear ? 0.f : st.ha1;
	h.ha2 = clear ? 0.f : st.ha2;
	h.hb1 = clear ? 0.f : st.hb1;
	h.hb2 = clear ? 0.f : st.hb2;
	float cf[5] = { st.b0, st.b1, st.b2, st.a1, st.a2 }, inc[5];
#pragma unroll
	for (int q = 0; q < 5; q++) {
		inc[q] = ((active ? rec->target[q] : 0.f) - cf[q]) / (float)F; // update_coeffs(F)
	}
	// MODE_E: histories of this side's cascaded stages, [effect][stage]{ha1,ha2,hb1,hb2} (dynamic shape: local memory)
	float fxh[GAS_MAX_EFFECTS][GAS_MAX_FILTER_STAGES][4];
	float *fxs = t.vs_fx + (size_t)voice * (GAS_MAX_EFFECTS * 2 * GAS_MAX_FILTER_STAGES * 4);
	if (MODE == MODE_E) {
		for (int e = 0; e < GAS_MAX_EFFECTS; e++) {
			for (int q = 0; q < GAS_MAX_FILTER_STAGES; q++) {
				for (int k = 0; k < 4; k++) {
					fxh[e][q][k] = e < n_fx ? fxs[((e * 2 + side) * GAS_MAX_FILTER_STAGES + q) * 4 + k] : 0.f;
				}
			}
		}
	}
	float pk = 0.f;
	const float4 *row = src_row >= 0 ? reinterpret_cast<const float4 *>(src + (size_t)src_row * src_stride) : nullptr;
	constexpr int NVRAW = NS * C;
	constexpr int NV = Pow2Ceil8<NVRAW>::value;
	float4 xb[4], xn[4]; // 8 frames = 64 bytes of this voice's row (both lanes of the voice fetch the same bytes); xn: next trip
#pragma unroll
	for (int u = 0; u < 4; u++) {
		xn[u] = (row && 2 * u < F) ? __ldg(row + u) : make_float4(0.f, 0.f, 0.f, 0.f);
	}
	for (int i0 = 0; i0 < F; i0 += 8) {
#pragma unroll
		for (int u = 0; u < 4; u++) {
			xb[u] = xn[u];
			xn[u] = (row && (i0 + 8 + 2 * u) < F) ? __ldg(row + ((i0 + 8) >> 1) + u) : make_float4(0.f, 0.f, 0.f, 0.f);
		}
		const float mt = (float)(i0 + (lane & 7)) / (float)F; // t of frame i0 + (lane & 7), handed round below (:591)
		float v[8 * NV]; // [frame][send * C + pair]
#pragma unroll
		for (int k = 0; k < 8; k++) {
			const float4 x4 = xb[k >> 1];
			float y = (k & 1) ? (side ? x4.w : x4.z) : (side ? x4.y : x4.x);
			const float tt = __shfl_sync(kFull, mt, (lane & 24) + k);
			const float omt = 1.0f - tt;
			if (i0 + k < F) {
				if (MODE == MODE_A) {
					if (filt) {
						y = biquad_step(h, y, cf[0], cf[1], cf[2], cf[3], cf[4]);
#pragma unroll
						for (int q = 0; q < 5; q++) {
							cf[q] += inc[q];
						}
					}
				} else { // cascaded constant-coefficient biquads per effect (upstream AudioEffectFilter::process)
					for (int e = 0; e < n_fx; e++) {
						const float b0 = rec->fx_coef[e][0], b1 = rec->fx_coef[e][1], b2 = rec->fx_coef[e][2], a1 = rec->fx_coef[e][3], a2 = rec->fx_coef[e][4];
						const int stages = rec->fx_stages[e];
						for (int q = 0; q < stages; q++) {
							float *hh = fxh[e][q];
							const float pre = y;
							y = y * b0 + hh[2] * b1 + hh[3] * b2 + hh[0] * a1 + hh[1] * a2;
							hh[1] = hh[0];
							hh[3] = hh[2];
							hh[2] = pre;
							hh[0] = y;
						}
					}
				}
				pk = fmaxf(pk, fabsf(y)); // block peak (audio_spatializer.cpp:436-443, :453-460)
			} else {
				y = 0.f;
			}
#pragma unroll
			for (int q = 0; q < NV; q++) {
				v[k * NV + q] = 0.f;
			}
#pragma unroll
			for (int s2 = 0; s2 < NS; s2++) {
#pragma unroll
				for (int c = 0; c < C; c++) {
					v[k * NV + s2 * C + c] = (nn[s2][c] * tt + omt * np[s2][c]) * y; // AudioServer ramp per send / pair
				}
			}
		}
		if (emit) {
			int base;
			bool owner;
			reduce_hi<8 * NV, 4>(v, lane, base, owner);
			constexpr int kLeft = (8 * NV) / 16 > 0 ? (8 * NV) / 16 : 1;
			if (owner) {
#pragma unroll
				for (int r = 0; r < kLeft; r++) {
					const int idx = base + r, k = idx / NV, q = idx % NV;
					const int s2 = q / C, c = q % C;
					if (q < NVRAW && (send0 + s2) < a.n_send && i0 + k < F) {
						acc_add(bus, tile, ((size_t)(bus_of[s2] * C + c) * F + i0 + k) * 2 + side, v[r]);
					}
				}
			}
		}
	}
	if (last_pass && active) {
		if (MODE == MODE_A && filt) {
			st.b0 = cf[0];
			st.b1 = cf[1];
			st.b2 = cf[2];
			st.a1 = cf[3];
			st.a2 = cf[4];
			st.ha1 = h.ha1;
			st.ha2 = h.ha2;
			st.hb1 = h.hb1;
			st.hb2 = h.hb2;
			*ps = st;
		}
		if (MODE == MODE_E) {
			for (int e = 0; e < n_fx; e++) {
				for (int q = 0; q < GAS_MAX_FILTER_STAGES; q++) {
					for (int k = 0; k < 4; k++) {
						fxs[((e * 2 + side) * GAS_MAX_FILTER_STAGES + q) * 4 + k] = fxh[e][q][k];
					}
				}
			}
		}
		if ((flags & GAS_VOICE_WANT_PEAK) && peaks) {
			reinterpret_cast<float *>(peaks + j)[side] = pk;
		}
	}
}

template <int MODE, int C>
__device__ void voice_chunk_s(const DevTables &t, const ChunkArgs &a, const gas_frame *__restrict__ src, int src_stride, int F,
		float *__restrict__ bus, gas_smem_addr tile, float2 *__restrict__ peaks) {
	const int n = a.n_send;
	if (n == 0) {
		voice_pass_s<MODE, C, 1>(t, a, 0, false, true, src, src_stride, F, bus, tile, peaks);
		return;
	}
	for (int s0 = 0; s0 < n; s0 += 2) {
		const bool last = s0 + 2 >= n;
		if (n - s0 >= 2) {
			voice_pass_s<MODE, C, 2>(t, a, s0, true, last, src, src_stride, F, bus, tile, peaks);
		} else {
			voice_pass_s<MODE, C, 1>(t, a, s0, true, last, src, src_stride, F, bus, tile, peaks);
		}
	}
}

// ---- filter-tile path: the ordinary (non-generic) classes of every mode -----------------------------------------------------
// The recurrence and the cross-voice sum are two different shapes of work, so they get two different thread mappings inside
// one CTA, handing over through shared memory, a tile of kFtFrames frames at a time:
//   source tiles    the voices' rows of the NEXT tile are copied global -> shared memory asynchronously (cp.async, one frame per
//                   thread and copy) while the current tile is worked on: a tile of work (> 1 us) covers the DRAM latency,
//                   which a register prefetch one 8-frame trip ahead (the per-warp form) does not.
//   filter phase    one lane per stream — (voice, side) in Mode A and effect chains, (voice, pair, side) in Mode B: nothing but
//                   the biquads (state in registers for the whole block), the Mode B mix_channel ramp and the block peak, on the
//                   tile in shared memory (Mode A / effects: in place; Mode B: from the x tile into the stream's own row of the
//                   y tile).  The lanes carry no bus ramps, no cross-lane reduction and no atomics (the per-warp form above spends
//                   three quarters of its instructions on those, inside the serial loop, and its red.shared.add.f32 compiles to a
//                   compare-and-swap loop).
//   contraction     all 256 threads: bus[b][c][i] += (p + t (n - p)) y[i] summed over the voices of the unit, weights
//                   {p_L, p_R, n_L - p_L, n_R - p_R} staged once per unit, read as 16-byte broadcasts, two packed FMAs (FFMA2) per
//                   (voice, row), accumulators in registers, added to the CTA's bus tile without atomics.  Mode B: a thread owns
//                   (frame, row group); Mode A / effects: the thread groups split the voices (ft_contract_split).  The ramp is
//                   evaluated as p + t (n - p) instead of the reference's n t + (1 - t) p: within 2 ulp of it, far inside the
//                   1e-5 tolerance.
// A unit is a batch of voices of one class (up to kFtVoices streams per side); the batch size is chosen so that the units of a
// block just cover the grid (the filter phase takes F serial steps however many lanes it has: more, smaller units are faster).
constexpr int kFtVoices = 56;                              // voices per unit (2 lanes each in the filter phase): 16384 voices on 2 x 148 CTAs
constexpr int kFtFrames = 64;                              // frames per tile
constexpr int kFtYStride = kFtFrames * 2 + 2;              // floats per voice row of a tile: +2 puts the 32 (voice, side) lanes of a warp on 32
                                                           // different banks in the filter phase (rows are then 8-byte aligned only, hence
                                                           // the 8-byte async copies)
constexpr int kFtRows = 16;                                // weight rows (send x pair) of a class: 4 sends (2 current + 2 fading out) x 4 pairs
constexpr int kFtYFloats = kFtVoices * kFtYStride;         // 7280 per tile buffer; two buffers: the tile being worked on and the next one in flight
constexpr int kFtWFloats = kFtVoices * kFtRows * 4;        // 3584
constexpr int kFtSmemBytes = (2 * kFtYFloats + kFtWFloats + kFtFrames) * 4; // tile buffers | weight tile | t of the tile's frames
constexpr int kFtFastStages = 4;                           // effect chains with up to this many biquads per side keep them in registers
static_assert((kFtYFloats * 4) % 16 == 0 && (kFtYStride * 4) % 8 == 0, "tile rows are written 8 bytes, the weight tile behind them is read 16 bytes at a time");
static_assert(kWarpsPerCta * 32 == 4 * kFtFrames, "contraction mapping: 4 thread groups (row groups / voice groups) x kFtFrames frames");

// (the planner sends voices with more than two buses on a side to the generic class, so an ordinary class has at most 4 sends;
// anything wider than the weight tile stays on the per-warp path)
template <int C>
__device__ __forceinline__ bool ft_class(const ClassInfo &ci, int ft_on) {
	return ft_on && !(ci.flags & CLS_GENERIC) && ci.n_send <= 4 && ci.n_send * C <= kFtRows;
}
// voices per unit of a class, given the block's stream budget per unit (Mode B: C streams per voice and side)
template <int C>
__device__ __forceinline__ int ft_vpu_of(const ClassInfo &ci, int budget) {
	return ci.mode == MODE_B ? max(1, min(28, budget / C)) : budget;
}

__device__ __forceinline__ int nth_set_bit(uint32_t m, int n) {
	for (int s = 0; s < n; s++) {
		m &= m - 1;
	}
	return m ? (__ffs(m) - 1) : 0;
}

// source frames [i0, i0 + kFtFrames) of the unit's voices -> tile buffer, asynchronously, one frame (8 bytes) per copy: consecutive
// threads take consecutive frames of a row (silent voices: zeros)
__device__ __forceinline__ void ft_prefetch(const gas_frame *__restrict__ src, int src_stride, const int2 *__restrict__ list, int nv, int i0, int F,
		float *buf) {
	const int n_frames = min(kFtFrames, F - i0);
	for (int idx = threadIdx.x; idx < nv * kFtFrames; idx += blockDim.x) {
		const int v = idx / kFtFrames, fr = idx % kFtFrames;
		if (fr < n_frames) {
			const int srow = list[v].y;
			float *dst = buf + v * kFtYStride + fr * 2;
			if (srow >= 0) {
				gas_cp_async_8(dst, src + (size_t)srow * src_stride + i0 + fr);
			} else {
				*reinterpret_cast<float2 *>(dst) = make_float2(0.f, 0.f);
			}
		}
	}
}

// weights of the unit's voices, rows [0, nrows) -> shared memory
template <int C>
__device__ __forceinline__ void ft_stage_weights(const InstSends *__restrict__ sends, const int2 *__restrict__ list, int nv, int nrows, float4 *s_w) {
	for (int idx = threadIdx.x; idx < nv * nrows; idx += blockDim.x) {
		const int v = idx / nrows, r = idx % nrows;
		const int s = r / C, c = r % C;
		const InstSends *snd = &sends[list[v].x];
		const float p0 = snd->vp[s][c][0], p1 = snd->vp[s][c][1];
		const float n0 = snd->vn[s][c][0], n1 = snd->vn[s][c][1];
		s_w[idx] = make_float4(p0, p1, n0 - p0, n1 - p1);
	}
}

// rows [0, nrows) x frames [i0, i0 + kFtFrames) of the unit: sum over its voices, added to the bus tile / buffers.
// A thread owns frame i0 + (tid % 64) of rows rg, rg + 4, ... (rg = tid / 64): N of them, a template parameter so that the loop over
// the voices carries exactly N weight loads and 2 N packed FMAs (the row count is uniform over the warp).
// PER_PAIR (Mode B): the voice's C pairs are C streams of their own, row (send, pair) sums stream `pair` of every voice.
template <int C, bool PER_PAIR, int N>
__device__ __forceinline__ void ft_contract_n(const float *s_y, const float4 *s_w, int nv, int nrows, const int (&rowoff)[kFtRows / 4], int i, int F,
		float *__restrict__ bus, float *s_tile) {
	const int f = threadIdx.x & (kFtFrames - 1), rg = threadIdx.x / kFtFrames;
	const float tt = (float)i / (float)F; // upstream _mix_step_for_channel: t = i / F
	const float2 t2 = make_float2(tt, tt);
	float2 acc[N];
#pragma unroll
	for (int q = 0; q < N; q++) {
		acc[q] = make_float2(0.f, 0.f);
	}
	const float *yp = s_y + f * 2;
	const float4 *wp = s_w + rg;
	if (!PER_PAIR) {
#pragma unroll 4
		for (int v = 0; v < nv; v++) {
			const float2 y = *reinterpret_cast<const float2 *>(yp + v * kFtYStride);
#pragma unroll
			for (int q = 0; q < N; q++) {
				const float4 w4 = wp[v * nrows + q * 4];
				const float2 w = gas_ffma2(t2, make_float2(w4.z, w4.w), make_float2(w4.x, w4.y));
				acc[q] = gas_ffma2(w, y, acc[q]);
			}
		}
	} else {
		int ycol[N]; // pair of row rg + 4 q = stream of the voice it sums
#pragma unroll
		for (int q = 0; q < N; q++) {
			ycol[q] = ((rg + q * 4) % C) * kFtYStride;
		}
#pragma unroll 4
		for (int v = 0; v < nv; v++) {
#pragma unroll
			for (int q = 0; q < N; q++) {
				const float2 y = *reinterpret_cast<const float2 *>(yp + v * (C * kFtYStride) + ycol[q]);
				const float4 w4 = wp[v * nrows + q * 4];
				const float2 w = gas_ffma2(t2, make_float2(w4.z, w4.w), make_float2(w4.x, w4.y));
				acc[q] = gas_ffma2(w, y, acc[q]);
			}
		}
	}
#pragma unroll
	for (int q = 0; q < N; q++) {
		const size_t o = (size_t)rowoff[q] + (size_t)i * 2;
		if (s_tile) {
			float2 *d = reinterpret_cast<float2 *>(s_tile + o); // one owner per (row, frame): no atomics
			float2 cur = *d;
			cur.x += acc[q].x;
			cur.y += acc[q].y;
			*d = cur;
		} else {
			atomicAdd(bus + o, acc[q].x);
			atomicAdd(bus + o + 1, acc[q].y);
		}
	}
}

template <int C, bool PER_PAIR>
__device__ __forceinline__ void ft_contract(const float *s_y, const float4 *s_w, int nv, int nrows, const int (&rowoff)[kFtRows / 4], int i0, int F,
		float *__restrict__ bus, float *s_tile) {
	const int f = threadIdx.x & (kFtFrames - 1), rg = threadIdx.x / kFtFrames;
	const int mine = (nrows - rg + 3) >> 2; // rows rg, rg + 4, ...: 0 .. kFtRows / 4
	const int i = i0 + f;
	if (i >= F) {
		return;
	}
	static_assert(kFtRows / 4 == 4, "one case per row count");
	switch (mine) {
		case 1:
			ft_contract_n<C, PER_PAIR, 1>(s_y, s_w, nv, nrows, rowoff, i, F, bus, s_tile);
			break;
		case 2:
			ft_contract_n<C, PER_PAIR, 2>(s_y, s_w, nv, nrows, rowoff, i, F, bus, s_tile);
			break;
		case 3:
			ft_contract_n<C, PER_PAIR, 3>(s_y, s_w, nv, nrows, rowoff, i, F, bus, s_tile);
			break;
		case 4:
			ft_contract_n<C, PER_PAIR, 4>(s_y, s_w, nv, nrows, rowoff, i, F, bus, s_tile);
			break;
		default:
			break; // no row of this thread's group in the class
	}
}

// Mode A / effect chains: every row of the class sums the SAME stream y_v, so four thread groups that own different rows would each
// read the whole y tile (shared-memory bandwidth, not issue slots, is what the contraction runs out of).  Here the four groups
// split the VOICES instead (group g: voices g, g + 4, ...) and every thread carries all NS * C rows of its frame: the y tile is
// read once.  The four partial sums of a (row, frame) meet in the bus tile one group after the other, a barrier apart.  Every
// thread of the CTA calls this (it contains barriers).
template <int C, int NS>
__device__ __forceinline__ void ft_contract_split(const float *s_y, const float4 *s_w, int nv, const int (&busoff)[4], int i0, int F,
		float *__restrict__ bus, float *s_tile) {
	constexpr int R = NS * C;
	const int f = threadIdx.x & (kFtFrames - 1), g = threadIdx.x / kFtFrames;
	const int i = i0 + f;
	const bool in_block = i < F;
	float2 acc[R];
#pragma unroll
	for (int r = 0; r < R; r++) {
		acc[r] = make_float2(0.f, 0.f);
	}
	if (in_block) {
		const float tt = (float)i / (float)F; // upstream _mix_step_for_channel: t = i / F
		const float2 t2 = make_float2(tt, tt);
		const float *yp = s_y + f * 2;
#pragma unroll 2
		for (int v = g; v < nv; v += 4) {
			const float2 y = *reinterpret_cast<const float2 *>(yp + v * kFtYStride);
			const float4 *wp = s_w + v * R;
#pragma unroll
			for (int r = 0; r < R; r++) {
				const float4 w4 = wp[r];
				const float2 w = gas_ffma2(t2, make_float2(w4.z, w4.w), make_float2(w4.x, w4.y));
				acc[r] = gas_ffma2(w, y, acc[r]);
			}
		}
	}
#pragma unroll
	for (int ph = 0; ph < 4; ph++) {
		if (ph == g && in_block) {
#pragma unroll
			for (int r = 0; r < R; r++) {
				const size_t o = (size_t)busoff[r / C] + ((size_t)(r % C) * F + i) * 2;
				if (s_tile) {
					float2 *d = reinterpret_cast<float2 *>(s_tile + o); // this group's turn: one owner per (row, frame)
					float2 cur = *d;
					cur.x += acc[r].x;
					cur.y += acc[r].y;
					*d = cur;
				} else {
					atomicAdd(bus + o, acc[r].x);
					atomicAdd(bus + o + 1, acc[r].y);
				}
			}
		}
		if (ph < 3) {
			__syncthreads();
		}
	}
}

// FT_COPY: Mode A below the filter threshold (peaks only): y = x.  FT_E_FAST1 / 2 / (4): effect chains whose longest member in the
// unit has 1 / 2 / up to 4 biquads per side (the slots beyond a lane's own chain are predicated off, so fewer slots = fewer issue slots)
// FT_E_U1 / U2 / U4: every voice of the unit has exactly 1 / 2 / 4 biquads per side (the usual case: one chain per spatializer), so the
// slots carry no predicate at all.
enum : int { FT_A = 0, FT_E_FAST = 1, FT_E_SLOW = 2, FT_COPY = 3, FT_E_FAST1 = 4, FT_E_FAST2 = 5, FT_E_U1 = 6, FT_E_U2 = 7, FT_E_U4 = 8 };
__host__ __device__ constexpr bool ft_uniform(int var) { return var == FT_E_U1 || var == FT_E_U2 || var == FT_E_U4; }
__host__ __device__ constexpr bool ft_is_fast(int var) { return var == FT_E_FAST || var == FT_E_FAST1 || var == FT_E_FAST2 || ft_uniform(var); }
__host__ __device__ constexpr int ft_slots(int var) {
	return (var == FT_E_FAST1 || var == FT_E_U1) ? 1 : ((var == FT_E_FAST2 || var == FT_E_U2) ? 2 : kFtFastStages);
}

// One unit: voices list[0 .. nv) of class `ci`.  Every thread of the CTA takes part (the barriers are CTA-wide).
template <int VAR, int C>
__device__ __noinline__ void ft_unit(const DevTables &t, const ClassInfo &ci, const VoiceRec *__restrict__ recs, const InstSends *__restrict__ sends,
		const int2 *__restrict__ list, int nv, const gas_frame *__restrict__ src, int src_stride, int F, float *__restrict__ bus, float *s_tile,
		float2 *__restrict__ peaks, float *s_y, float4 *s_w) {
	// the first source tile is requested before anything else: its DRAM round trip runs beside the dependent loads of the set-up
	// (list -> voice record -> processor state, list -> sends -> weight tile)
	ft_prefetch(src, src_stride, list, nv, 0, F, s_y);
	const int vl = threadIdx.x >> 1, side = threadIdx.x & 1; // filter phase: lane = (voice, side)
	const bool active = vl < nv;
	const int j = active ? list[vl].x : 0;
	const VoiceRec *rec = &recs[j];
	const int voice = active ? rec->voice : 0;
	const uint32_t flags = active ? rec->flags : 0u;
	const bool filt = VAR == FT_A;

	// FT_A: the interpolated high-shelf processor of this side (index 0 left, 1 right, reference audio_spatializer_3d.cpp:524-529)
	gas_processor_state *ps = t.vs_proc + (size_t)voice * 8 + side;
	gas_processor_state st{};
	Biquad h{};
	float cf[5] = {}, inc[5] = {};
	if (VAR == FT_A) {
		if (active && filt) {
			st = *ps;
		}
		const bool clear = (flags >> 8) & 1u; // is_just_started => clear history (:518-521)
		h.ha1 = clear ? 0.f : st.ha1;
		h.ha2 = clear ? 0.f : st.ha2;
		h.hb1 = clear ? 0.f : st.hb1;
		h.hb2 = clear ? 0.f : st.hb2;
		cf[0] = st.b0;
		cf[1] = st.b1;
		cf[2] = st.b2;
		cf[3] = st.a1;
		cf[4] = st.a2;
#pragma unroll
		for (int q = 0; q < 5; q++) {
			inc[q] = ((active ? rec->target[q] : 0.f) - cf[q]) / (float)F; // update_coeffs(F)
		}
	}
	// effect chains: cascaded constant-coefficient biquads (upstream AudioEffectFilter::process), histories
	// [effect][side][stage]{ha1,ha2,hb1,hb2} in vs_fx
	float *fxs = t.vs_fx + (size_t)voice * (GAS_MAX_EFFECTS * 2 * GAS_MAX_FILTER_STAGES * 4);
	const int n_fx = ((ft_is_fast(VAR) || VAR == FT_E_SLOW) && active) ? rec->n_fx : 0;
	constexpr int kSlots = ft_slots(VAR);
	// FT_E_FAST: the chain flattened to at most kFtFastStages biquads, coefficients and histories in registers
	int n_slot = 0;
	int e_of[kSlots] = {}, q_of[kSlots] = {};
	float sc[kSlots][5] = {}, sh[kSlots][4] = {};
	// FT_E_SLOW: any chain (dynamic shape: local memory)
	float fxh[VAR == FT_E_SLOW ? GAS_MAX_EFFECTS : 1][VAR == FT_E_SLOW ? GAS_MAX_FILTER_STAGES : 1][4];
	if (ft_is_fast(VAR)) {
#pragma unroll
		for (int e = 0; e < GAS_MAX_EFFECTS; e++) {
			const int stages = e < n_fx ? rec->fx_stages[e] : 0;
#pragma unroll
			for (int q = 0; q < GAS_MAX_FILTER_STAGES; q++) {
				const bool on = q < stages;
#pragma unroll
				for (int s = 0; s < kSlots; s++) {
					if (on && n_slot == s) {
						e_of[s] = e;
						q_of[s] = q;
					}
				}
				n_slot += on ? 1 : 0;
			}
		}
#pragma unroll
		for (int s = 0; s < kSlots; s++) {
			if (s < n_slot) {
#pragma unroll
				for (int k = 0; k < 5; k++) {
					sc[s][k] = rec->fx_coef[e_of[s]][k];
				}
#pragma unroll
				for (int k = 0; k < 4; k++) {
					sh[s][k] = fxs[((e_of[s] * 2 + side) * GAS_MAX_FILTER_STAGES + q_of[s]) * 4 + k];
				}
			}
		}
	}
	if (VAR == FT_E_SLOW) {
		for (int e = 0; e < GAS_MAX_EFFECTS; e++) {
			for (int q = 0; q < GAS_MAX_FILTER_STAGES; q++) {
				for (int k = 0; k < 4; k++) {
					fxh[VAR == FT_E_SLOW ? e : 0][VAR == FT_E_SLOW ? q : 0][k] = e < n_fx ? fxs[((e * 2 + side) * GAS_MAX_FILTER_STAGES + q) * 4 + k] : 0.f;
				}
			}
		}
	}

	const int R = ci.n_send * C; // <= kFtRows (ft_class)
	ft_stage_weights<C>(sends, list, nv, R, s_w); // visible to the contraction after the barrier behind the first filter phase

	float pk = 0.f;
	// one frame of this lane's stream: the biquads and the block peak (audio_spatializer.cpp:436-443, :453-460)
	auto step = [&](float y) -> float {
		if (VAR == FT_A) {
			y = biquad_step(h, y, cf[0], cf[1], cf[2], cf[3], cf[4]);
#pragma unroll
			for (int q = 0; q < 5; q++) { // process_one_interp: coeffs += incr
				cf[q] += inc[q];
			}
		} else if (ft_is_fast(VAR)) {
#pragma unroll
			for (int s = 0; s < kSlots; s++) {
				if (ft_uniform(VAR) || s < n_slot) {
					const float pre = y;
					y = y * sc[s][0] + sh[s][2] * sc[s][1] + sh[s][3] * sc[s][2] + sh[s][0] * sc[s][3] + sh[s][1] * sc[s][4];
					sh[s][1] = sh[s][0];
					sh[s][3] = sh[s][2];
					sh[s][2] = pre;
					sh[s][0] = y;
				}
			}
		} else if (VAR == FT_E_SLOW) {
			for (int e = 0; e < n_fx; e++) {
				const float b0 = rec->fx_coef[e][0], b1 = rec->fx_coef[e][1], b2 = rec->fx_coef[e][2], a1 = rec->fx_coef[e][3], a2 = rec->fx_coef[e][4];
				const int stages = rec->fx_stages[e];
				for (int q = 0; q < stages; q++) {
					float *hh = fxh[VAR == FT_E_SLOW ? e : 0][VAR == FT_E_SLOW ? q : 0];
					const float pre = y;
					y = y * b0 + hh[2] * b1 + hh[3] * b2 + hh[0] * a1 + hh[1] * a2;
					hh[1] = hh[0];
					hh[3] = hh[2];
					hh[2] = pre;
					hh[0] = y;
				}
			}
		}
		pk = fmaxf(pk, fabsf(y));
		return y;
	};
	// contraction: float offset of (bus of send s, pair 0, frame 0) in the bus layout
	int busoff[4];
#pragma unroll
	for (int q = 0; q < 4; q++) {
		busoff[q] = q < ci.n_send ? nth_set_bit(ci.mask, q) * C * F * 2 : 0;
	}
	for (int i0 = 0, ti = 0; i0 < F; i0 += kFtFrames, ti++) {
		float *buf = s_y + (ti & 1) * kFtYFloats;
		gas_cp_async_wait_all(); // this thread's copies of the tile have landed ...
		__syncthreads();         // ... and everybody else's; the other buffer (read by the previous contraction) is free
		if (i0 + kFtFrames < F) {
			ft_prefetch(src, src_stride, list, nv, i0 + kFtFrames, F, s_y + ((ti + 1) & 1) * kFtYFloats);
		}
		// ---- filter phase: in place on the tile ----
		if (active) {
			float *xr = buf + vl * kFtYStride + side;
			// whole 8-frame trips of the tile: no per-frame bounds test between the recurrence steps, and the next trip's frames are
			// read from shared memory while this trip's recurrence runs
			const int n_frames = min(kFtFrames, F - i0);
			const int trips = n_frames >> 3;
			float xv[8], xq[8];
#pragma unroll
			for (int k = 0; k < 8; k++) {
				xv[k] = trips > 0 ? xr[k * 2] : 0.f;
			}
#pragma unroll 2
			for (int tr = 0; tr < trips; tr++) {
				const int k0 = tr * 8;
				const bool more = tr + 1 < trips;
#pragma unroll
				for (int k = 0; k < 8; k++) {
					xq[k] = more ? xr[(k0 + 8 + k) * 2] : 0.f;
				}
#pragma unroll
				for (int k = 0; k < 8; k++) {
					xr[(k0 + k) * 2] = step(xv[k]);
				}
#pragma unroll
				for (int k = 0; k < 8; k++) {
					xv[k] = xq[k];
				}
			}
			for (int k = trips * 8; k < n_frames; k++) { // the block's last frames when F is not a multiple of 8
				xr[k * 2] = step(xr[k * 2]);
			}
		}
		__syncthreads();
		// ---- contraction ----
		switch (ci.n_send) { // (0: nothing to add, the block only advances state / peaks)
			case 1:
				ft_contract_split<C, 1>(buf, s_w, nv, busoff, i0, F, bus, s_tile);
				break;
			case 2:
				ft_contract_split<C, 2>(buf, s_w, nv, busoff, i0, F, bus, s_tile);
				break;
			case 3:
				ft_contract_split<C, 3>(buf, s_w, nv, busoff, i0, F, bus, s_tile);
				break;
			case 4:
				ft_contract_split<C, 4>(buf, s_w, nv, busoff, i0, F, bus, s_tile);
				break;
			default:
				break;
		}
	}
	__syncthreads(); // the next unit's copies and weights overwrite what the last contraction read
	if (active) {
		if (VAR == FT_A && filt) {
			st.b0 = cf[0];
			st.b1 = cf[1];
			st.b2 = cf[2];
			st.a1 = cf[3];
			st.a2 = cf[4];
			st.ha1 = h.ha1;
			st.ha2 = h.ha2;
			st.hb1 = h.hb1;
			st.hb2 = h.hb2;
			*ps = st;
		}
		if (ft_is_fast(VAR)) {
#pragma unroll
			for (int s = 0; s < kSlots; s++) {
				if (s < n_slot) {
#pragma unroll
					for (int k = 0; k < 4; k++) {
						fxs[((e_of[s] * 2 + side) * GAS_MAX_FILTER_STAGES + q_of[s]) * 4 + k] = sh[s][k];
					}
				}
			}
		}
		if (VAR == FT_E_SLOW) {
			for (int e = 0; e < n_fx; e++) {
				for (int q = 0; q < GAS_MAX_FILTER_STAGES; q++) {
					for (int k = 0; k < 4; k++) {
						fxs[((e * 2 + side) * GAS_MAX_FILTER_STAGES + q) * 4 + k] = fxh[VAR == FT_E_SLOW ? e : 0][VAR == FT_E_SLOW ? q : 0][k];
					}
				}
			}
		}
		if ((flags & GAS_VOICE_WANT_PEAK) && peaks) {
			reinterpret_cast<float *>(peaks + j)[side] = pk;
		}
	}
}


// Mode B unit (with or without the attenuation filter: up to 2C biquads per voice): one lane per (voice, pair, side) stream.
// The voices' source rows arrive in two small x tiles (this tile and the next one in flight); every stream applies the
// mix_channel ramp (reference audio_spatializer_3d.cpp:591-593) and its interpolated high-shelf (:594-595) and writes its own row
// of the y tile; the contraction then sums, for every (send, pair) row, stream `pair` of every voice under the AudioServer ramp.
constexpr int kFtVoicesB = 28; // x tiles: the second tile buffer, halved (ft_vpu_of)
static_assert(kFtVoicesB * kFtYStride * 2 <= kFtYFloats, "two x tiles share one tile buffer");

template <int C, bool FILT>
__device__ __noinline__ void ft_unit_b(const DevTables &t, const ClassInfo &ci, const VoiceRec *__restrict__ recs, const InstSends *__restrict__ sends,
		const int2 *__restrict__ list, int nv, const gas_frame *__restrict__ src, int src_stride, int F, float *__restrict__ bus, float *s_tile,
		float2 *__restrict__ peaks, float *s_y, float4 *s_w, float *s_t) {
	constexpr int kPer = 2 * C; // lanes per voice
	ft_prefetch(src, src_stride, list, nv, 0, F, s_y + kFtYFloats); // first x tile: requested before the set-up's dependent loads
	const int vl = (int)threadIdx.x / kPer, rem = (int)threadIdx.x % kPer, c = rem >> 1, side = rem & 1;
	const bool active = vl < nv;
	const int j = active ? list[vl].x : 0;
	const VoiceRec *rec = &recs[j];
	const int voice = active ? rec->voice : 0;
	const uint32_t flags = active ? rec->flags : 0u;
	const float m_prev = active ? rec->m_prev[c][side] : 0.f;
	const float m_new = active ? rec->m_new[c][side] : 0.f;
	constexpr bool filt = FILT; // CLS_FILT of the class: a template parameter keeps the test out of the recurrence
	gas_processor_state *ps = t.vs_proc + (size_t)voice * 8 + rem; // processor index pair * 2 + (left ? 0 : 1)
	gas_processor_state st{};
	if (active && filt) {
		st = *ps;
	}
	const bool clear = (flags >> (8 + c)) & 1u; // is_just_started => clear history (:583-586)
	Biquad h;
	h.ha1 = clear ? 0.f : st.ha1;
	h.ha2 = clear ? 0.f : st.ha2;
	h.hb1 = clear ? 0.f : st.hb1;
	h.hb2 = clear ? 0.f : st.hb2;
	float cf[5] = { st.b0, st.b1, st.b2, st.a1, st.a2 }, inc[5];
#pragma unroll
	for (int q = 0; q < 5; q++) { // update_coeffs(F): per-sample increment towards the target
		inc[q] = ((active ? rec->target[q] : 0.f) - cf[q]) / (float)F;
	}
	const int R = ci.n_send * C; // <= kFtRows (ft_class)
	ft_stage_weights<C>(sends, list, nv, R, s_w);
	int rowoff[kFtRows / 4];
#pragma unroll
	for (int q = 0; q < kFtRows / 4; q++) {
		const int r = (int)threadIdx.x / kFtFrames + q * 4;
		rowoff[q] = r < R ? (nth_set_bit(ci.mask, r / C) * C + r % C) * F * 2 : 0;
	}
	float pk = 0.f;
	float *ybuf = s_y, *xbuf = s_y + kFtYFloats; // y tile: nv * C stream rows; x tiles: 2 x nv voice rows
	auto step = [&](float x, float tt) -> float {
		const float omt = 1.0f - tt;
		const float vol = m_new * tt + omt * m_prev; // :592
		float y = vol * x;                            // :593
		if (filt) {
			y = biquad_step(h, y, cf[0], cf[1], cf[2], cf[3], cf[4]); // :594-595
#pragma unroll
			for (int q = 0; q < 5; q++) { // process_one_interp: coeffs += incr
				cf[q] += inc[q];
			}
		}
		pk = fmaxf(pk, fabsf(y));
		return y;
	};
	for (int i0 = 0, ti = 0; i0 < F; i0 += kFtFrames, ti++) {
		const float *xb = xbuf + (ti & 1) * (kFtVoicesB * kFtYStride);
		gas_cp_async_wait_all();
		if ((int)threadIdx.x < kFtFrames) {
			s_t[threadIdx.x] = (float)(i0 + (int)threadIdx.x) / (float)F; // t of the tile's frames (:591), once per frame instead of once per stream
		}
		__syncthreads();
		if (i0 + kFtFrames < F) {
			ft_prefetch(src, src_stride, list, nv, i0 + kFtFrames, F, xbuf + ((ti + 1) & 1) * (kFtVoicesB * kFtYStride));
		}
		// ---- filter phase ----
		if (active) {
			const float *xr = xb + vl * kFtYStride + side;
			float *yr = ybuf + (vl * C + c) * kFtYStride + side;
			const int n_frames = min(kFtFrames, F - i0);
			const int trips = n_frames >> 3;
			float xv[8], xq[8];
#pragma unroll
			for (int k = 0; k < 8; k++) {
				xv[k] = trips > 0 ? xr[k * 2] : 0.f;
			}
#pragma unroll 2
			for (int tr = 0; tr < trips; tr++) {
				const int k0 = tr * 8;
				const bool more = tr + 1 < trips;
				float tv[8];
#pragma unroll
				for (int k = 0; k < 8; k++) {
					xq[k] = more ? xr[(k0 + 8 + k) * 2] : 0.f;
					tv[k] = s_t[k0 + k];
				}
#pragma unroll
				for (int k = 0; k < 8; k++) {
					yr[(k0 + k) * 2] = step(xv[k], tv[k]);
				}
#pragma unroll
				for (int k = 0; k < 8; k++) {
					xv[k] = xq[k];
				}
			}
			for (int k = trips * 8; k < n_frames; k++) {
				yr[k * 2] = step(xr[k * 2], s_t[k]);
			}
		}
		__syncthreads();
		// ---- contraction ----
		ft_contract<C, true>(ybuf, s_w, nv, R, rowoff, i0, F, bus, s_tile);
	}
	__syncthreads();
	if (active && filt) {
		st.b0 = cf[0];
		st.b1 = cf[1];
		st.b2 = cf[2];
		st.a1 = cf[3];
		st.a2 = cf[4];
		st.ha1 = h.ha1;
		st.ha2 = h.ha2;
		st.hb1 = h.hb1;
		st.hb2 = h.hb2;
		*ps = st;
	}
	// block peak: max over the voice's pairs, per side (audio_spatializer.cpp:436-443), handed over through the (free) y tile
	if (peaks) {
		if (active) {
			ybuf[(vl * C + c) * kFtYStride + side] = pk;
		}
		__syncthreads();
		if (active && c == 0 && (flags & GAS_VOICE_WANT_PEAK)) {
			float m = 0.f;
#pragma unroll
			for (int cc = 0; cc < C; cc++) {
				m = fmaxf(m, ybuf[(vl * C + cc) * kFtYStride + side]);
			}
			reinterpret_cast<float *>(peaks + j)[side] = m;
		}
		__syncthreads();
	}
}

// biquads per side of the longest effect chain among voices list[0 .. nv) (CTA-uniform result; every thread calls it)
__device__ int ft_max_stages(const VoiceRec *__restrict__ recs, const int2 *__restrict__ list, int nv, int *s_scratch, bool *all_equal) {
	if (threadIdx.x == 0) {
		s_scratch[0] = 0;
		s_scratch[1] = 1 << 20;
	}
	__syncthreads();
	if ((int)threadIdx.x < nv) {
		const VoiceRec *rec = &recs[list[threadIdx.x].x];
		int total = 0;
		for (int e = 0; e < rec->n_fx && e < GAS_MAX_EFFECTS; e++) {
			total += rec->fx_stages[e];
		}
		atomicMax(&s_scratch[0], total);
		atomicMin(&s_scratch[1], total);
	}
	__syncthreads();
	const int m = s_scratch[0];
	*all_equal = s_scratch[1] == m;
	__syncthreads(); // the scratch words may be reset by the next unit
	return m;
}

// units of a class on the per-warp path: Mode B runs 4 voices per warp; the generic class (voices whose own class found no
// slot, or with more than two buses on a side) runs one voice per unit, because its voices do not share a send layout.
// Mode A / effect-chain classes take the filter-tile path (a class too wide for its weight tile would run 16 voices per warp).
template <int C>
__device__ __forceinline__ int units_of(const ClassInfo &ci, int ft_on) {
	if (ci.flags & CLS_GENERIC) {
		return ci.count;
	}
	if (ft_class<C>(ci, ft_on)) {
		return 0;
	}
	return ci.mode == MODE_B ? (ci.count + 3) / 4 : (ci.count + 15) / 16;
}

template <int C>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 2) k_mix_voice(DevTables t, GlobalCfg g, BlockPlan plan,
		const gas_frame *__restrict__ src, int src_stride, int F, float *__restrict__ bus, float2 *__restrict__ peaks,
		const float4 *__restrict__ rep, int bus_f4, int replicas, int tile_floats, int early_look, int ft_on) {
	GAS_DYN_SMEM(float, 16, s_tile);
	__shared__ ClassInfo s_cls[GAS_MAX_CLASSES];
	__shared__ int s_ncls;
	__shared__ int s_ft_scratch[2];
	GAS_GRID_DEP_LAUNCH();
	// Programmatic dependent launch: this kernel may become resident while the streaming kernel (its stream predecessor)
	// still runs.  What it reads first — the class table of the block — was written by the prologue, which completed
	// before the streaming kernel started, so it is read BEFORE the dependency wait: a block without voice-parallel work
	// (nothing filtered, no peaks: the common case of the unfiltered mix) ends here, one thread staying behind to keep the
	// stream order intact, and costs the step nothing but this look.
	if (!early_look) {
		GAS_GRID_DEP_WAIT(); // not behind the streaming kernel: the class table may not be final before the dependency is met
	}
	// Which block: launches of this kernel are counted on the device (BLK_Q), so that replayed graphs need no host-side
	// block index; the plan of block q lives in slot q % GAS_PLAN_DEPTH and was published before this kernel was launched.
	__shared__ int s_q;
	if (threadIdx.x == 0) {
		s_q = *(volatile const int32_t *)&t.blk[BLK_Q];
	}
	__syncthreads();
	const int slot_p = s_q & (GAS_PLAN_DEPTH - 1);
	const PlanHdr *hdr = &plan.hdr[slot_p];
	const int n_vcls = __ldcg(&hdr->n_vcls);
	if (n_vcls <= 0) {
		// nothing filtered, no peaks (the common case of the unfiltered mix): the block costs this kernel one look
		if (threadIdx.x == 0) {
			if (blockIdx.x == 0) {
				GAS_GRID_DEP_WAIT(); // keeps the stream order intact for whatever follows
			}
			if (atomicAdd(&t.blk[BLK_Q_TICKET], 1) == (int)gridDim.x - 1) {
				t.blk[BLK_Q_TICKET] = 0;
				__threadfence();
				*(volatile int32_t *)&t.blk[BLK_Q] = s_q + 1;
			}
		}
		return;
	}
	GAS_GRID_DEP_WAIT();
	{
		// compact table of the voice-parallel classes of this block (written by the planner's last CTA)
		const int4 *srcw = reinterpret_cast<const int4 *>(hdr->vcls);
		int4 *dstw = reinterpret_cast<int4 *>(s_cls);
		const int words = n_vcls * (int)(sizeof(ClassInfo) / 16);
		for (int i = threadIdx.x; i < words; i += blockDim.x) {
			dstw[i] = __ldcg(srcw + i);
		}
		if (threadIdx.x == 0) {
			s_ncls = n_vcls;
		}
	}
	__syncthreads();
	// CTA-wide accumulation tile: the voices of all chunks this CTA visits are summed in shared memory and reach the
	// bus buffers as one vector reduction per 16 bytes (instead of one scalar atomic per element per 32 voices)
	const gas_smem_addr tile = tile_floats > 0 ? gas_smem_u32(s_tile) : 0u;
	// filter-tile units (Mode A / effect chains): batches of `ft_vpu` voices, sized so that the block's units just cover the grid
	float *s_y = s_tile + tile_floats;
	float4 *s_w = reinterpret_cast<float4 *>(s_y + 2 * kFtYFloats);
	float *s_t = reinterpret_cast<float *>(s_w + kFtWFloats / 4);
	int ft_budget = 8, ft_units = 0;
	{
		int ft_streams = 0; // per side
		for (int c = 0; c < s_ncls; c++) {
			ft_streams += ft_class<C>(s_cls[c], ft_on) ? s_cls[c].count * (s_cls[c].mode == MODE_B ? C : 1) : 0;
		}
		ft_budget = (((ft_streams + (int)gridDim.x - 1) / (int)gridDim.x) + 7) & ~7;
		ft_budget = min(max(ft_budget, 8), kFtVoices);
		for (int c = 0; c < s_ncls; c++) {
			const int vpu = ft_vpu_of<C>(s_cls[c], ft_budget);
			ft_units += ft_class<C>(s_cls[c], ft_on) ? (s_cls[c].count + vpu - 1) / vpu : 0;
		}
	}
	bool cta_has_work = false;
	{
		int units = 0;
		for (int c = 0; c < s_ncls; c++) {
			units += units_of<C>(s_cls[c], ft_on);
		}
		// both kinds of units are dealt round-robin to the CTAs (the per-warp ones then to the warps of a CTA)
		cta_has_work = units > (int)blockIdx.x || ft_units > (int)blockIdx.x;
	}
	if (tile && cta_has_work) {
		for (int i = threadIdx.x; i < tile_floats / 4; i += blockDim.x) {
			reinterpret_cast<float4 *>(s_tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
		}
		__syncthreads();
	}
	// ---- filter-tile units: the whole CTA works on one unit at a time ----
	for (int unit = (int)blockIdx.x; unit < ft_units; unit += (int)gridDim.x) {
		int c = 0, k = unit;
		for (; c < s_ncls; c++) {
			const int vpu = ft_vpu_of<C>(s_cls[c], ft_budget);
			const int n = ft_class<C>(s_cls[c], ft_on) ? (s_cls[c].count + vpu - 1) / vpu : 0;
			if (k < n) {
				break;
			}
			k -= n;
		}
		const ClassInfo &ci = s_cls[c];
		const int ft_vpu = ft_vpu_of<C>(ci, ft_budget);
		const VoiceRec *recs = plan.rec + (size_t)slot_p * g.max_voices;
		const InstSends *sends = plan.sends + (size_t)slot_p * g.max_voices;
		const int2 *list = plan_list(plan, slot_p, ci.slot, g.max_voices) + k * ft_vpu;
		const int nv = min(ft_vpu, ci.count - k * ft_vpu);
		float *tile_p = tile_floats > 0 ? s_tile : nullptr;
		if (ci.mode == MODE_B) {
			if (ci.flags & CLS_FILT) {
				ft_unit_b<C, true>(t, ci, recs, sends, list, nv, src, src_stride, F, bus, tile_p, peaks, s_y, s_w, s_t);
			} else {
				ft_unit_b<C, false>(t, ci, recs, sends, list, nv, src, src_stride, F, bus, tile_p, peaks, s_y, s_w, s_t);
			}
		} else if (ci.mode == MODE_A) {
			if (ci.flags & CLS_FILT) {
				ft_unit<FT_A, C>(t, ci, recs, sends, list, nv, src, src_stride, F, bus, tile_p, peaks, s_y, s_w);
			} else {
				ft_unit<FT_COPY, C>(t, ci, recs, sends, list, nv, src, src_stride, F, bus, tile_p, peaks, s_y, s_w);
			}
		} else {
			bool same = false;
			const int longest = ft_max_stages(recs, list, nv, s_ft_scratch, &same);
			if (same && longest == 1) {
				ft_unit<FT_E_U1, C>(t, ci, recs, sends, list, nv, src, src_stride, F, bus, tile_p, peaks, s_y, s_w);
			} else if (same && longest == 2) {
				ft_unit<FT_E_U2, C>(t, ci, recs, sends, list, nv, src, src_stride, F, bus, tile_p, peaks, s_y, s_w);
			} else if (same && longest == 4) {
				ft_unit<FT_E_U4, C>(t, ci, recs, sends, list, nv, src, src_stride, F, bus, tile_p, peaks, s_y, s_w);
			} else if (longest <= 1) {
				ft_unit<FT_E_FAST1, C>(t, ci, recs, sends, list, nv, src, src_stride, F, bus, tile_p, peaks, s_y, s_w);
			} else if (longest == 2) {
				ft_unit<FT_E_FAST2, C>(t, ci, recs, sends, list, nv, src, src_stride, F, bus, tile_p, peaks, s_y, s_w);
			} else if (longest <= kFtFastStages) {
				ft_unit<FT_E_FAST, C>(t, ci, recs, sends, list, nv, src, src_stride, F, bus, tile_p, peaks, s_y, s_w);
			} else {
				ft_unit<FT_E_SLOW, C>(t, ci, recs, sends, list, nv, src, src_stride, F, bus, tile_p, peaks, s_y, s_w);
			}
		}
	}
	// ---- per-warp units (Mode B, generic classes) ----
	const int my_warp = threadIdx.x >> 5;
	// Units (4-voice chunks of the stream-parallel Mode B path, single voices of the generic classes) are numbered class by class and
	// dealt round-robin to the CTAs, then to the warps of a CTA: this warp owns units b + grid * (w + 8 r), r = 0, 1, ...
	int total_units = 0;
	for (int c = 0; c < s_ncls; c++) {
		total_units += units_of<C>(s_cls[c], ft_on);
	}
	for (int unit = (int)blockIdx.x + (int)gridDim.x * my_warp; unit < total_units; unit += (int)gridDim.x * kWarpsPerCta) {
		int c = 0, k = unit;
		for (; c < s_ncls; c++) { // class and chunk of the unit: Mode B 4 voices per warp, Mode A / effect chains 16
			const int chunks = units_of<C>(s_cls[c], ft_on);
			if (k < chunks) {
				break;
			}
			k -= chunks;
		}
		const ClassInfo &ci = s_cls[c];
		ChunkArgs a;
		a.rec = plan.rec + (size_t)slot_p * g.max_voices;
		a.sends = plan.sends + (size_t)slot_p * g.max_voices;
		a.list = plan_list(plan, slot_p, ci.slot, g.max_voices);
		a.count = ci.count;
		a.chunk = k;
		a.cls_flags = ci.flags;
		a.mask = ci.mask;
		a.n_send = ci.n_send;
		if (ci.flags & CLS_GENERIC) { // one voice, with its own send layout
			a.list += k;
			a.count = 1;
			a.chunk = 0;
			const InstSends *own = &a.sends[a.list[0].x];
			a.mask = own->mask;
			a.n_send = own->n;
		}
		switch (ci.mode) {
			case MODE_A:
				voice_chunk_s<MODE_A, C>(t, a, src, src_stride, F, bus, tile, peaks);
				break;
			case MODE_B:
				voice_chunk_b<C>(t, a, src, src_stride, F, bus, tile, peaks);
				break;
			default:
				voice_chunk_s<MODE_E, C>(t, a, src, src_stride, F, bus, tile, peaks);
				break;
		}
	}
	if (tile && cta_has_work) {
		__syncthreads();
		for (int i = threadIdx.x; i < tile_floats / 4; i += blockDim.x) {
			const float4 v = reinterpret_cast<const float4 *>(s_tile)[i];
			if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) {
				gas_red_add_v4(bus + (size_t)i * 4, v.x, v.y, v.z, v.w);
			}
		}
	}	__syncthreads();
	if (threadIdx.x == 0 && atomicAdd(&t.blk[BLK_Q_TICKET], 1) == (int)gridDim.x - 1) {
		t.blk[BLK_Q_TICKET] = 0;
		__threadfence();
		*(volatile int32_t *)&t.blk[BLK_Q] = s_q + 1;
	}
}

} // namespace

cudaError_t launch_mix_voice(gas_ctx *ctx, const gas_frame *d_src, int src_stride, int frames, gas_frame *d_bus,
		gas_frame *d_peaks, cudaStream_t st, bool after_stream) {
	const bool pdl = after_stream && (ctx->pdl & 4) != 0;
	const int early_look = after_stream ? 1 : 0;
	const int bus_f4 = gas_bus_f4(ctx, frames);
	const int threads = kWarpsPerCta * 32;
	int grid = ctx->num_sms * 2; // two CTAs per SM (128 registers): many chunks per CTA make the shared-memory tile pay
	const int fold_threads = bus_f4;
	if (grid * threads < fold_threads) {
		grid = (fold_threads + threads - 1) / threads;
	}
	int tile_floats = bus_f4 * 4;
	if ((size_t)tile_floats * sizeof(float) > (size_t)kMaxTileBytes) {
		tile_floats = 0; // too many buses x frames for shared memory: scalar atomics straight into the bus buffers
	}
	const size_t smem = (size_t)tile_floats * sizeof(float) + kFtSmemBytes; // bus tile | y tile | weight tile
	if (!ctx->k3_smem_attr_set) {
		cudaError_t ea = cudaFuncSetAttribute(k_mix_voice<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxTileBytes + kFtSmemBytes);
		ea = ea != cudaSuccess ? ea : cudaFuncSetAttribute(k_mix_voice<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxTileBytes + kFtSmemBytes);
		ea = ea != cudaSuccess ? ea : cudaFuncSetAttribute(k_mix_voice<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxTileBytes + kFtSmemBytes);
		ea = ea != cudaSuccess ? ea : cudaFuncSetAttribute(k_mix_voice<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxTileBytes + kFtSmemBytes);
		if (ea != cudaSuccess) {
			return ea;
		}
		// two CTAs of ~115 KB per SM need (nearly) the whole shared-memory carve-out; a hint, so its status does not matter
		cudaFuncSetAttribute(k_mix_voice<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
		cudaFuncSetAttribute(k_mix_voice<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
		cudaFuncSetAttribute(k_mix_voice<3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
		cudaFuncSetAttribute(k_mix_voice<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
		(void)cudaGetLastError();
		ctx->k3_smem_attr_set = true;
	}
	cudaError_t e = cudaSuccess;
	switch (ctx->g.channels) {
		case 1:
			e = gas_launch(k_mix_voice<1>, dim3(grid), dim3(threads), smem, st, pdl, ctx->t, ctx->g, ctx->plan, d_src, src_stride, frames, (float *)d_bus, (float2 *)d_peaks, (const float4 *)ctx->d_rep, bus_f4, ctx->replicas, tile_floats, early_look, ctx->k3_legacy ? 0 : 1);
			break;
		case 2:
			e = gas_launch(k_mix_voice<2>, dim3(grid), dim3(threads), smem, st, pdl, ctx->t, ctx->g, ctx->plan, d_src, src_stride, frames, (float *)d_bus, (float2 *)d_peaks, (const float4 *)ctx->d_rep, bus_f4, ctx->replicas, tile_floats, early_look, ctx->k3_legacy ? 0 : 1);
			break;
		case 3:
			e = gas_launch(k_mix_voice<3>, dim3(grid), dim3(threads), smem, st, pdl, ctx->t, ctx->g, ctx->plan, d_src, src_stride, frames, (float *)d_bus, (float2 *)d_peaks, (const float4 *)ctx->d_rep, bus_f4, ctx->replicas, tile_floats, early_look, ctx->k3_legacy ? 0 : 1);
			break;
		default:
			e = gas_launch(k_mix_voice<4>, dim3(grid), dim3(threads), smem, st, pdl, ctx->t, ctx->g, ctx->plan, d_src, src_stride, frames, (float *)d_bus, (float2 *)d_peaks, (const float4 *)ctx->d_rep, bus_f4, ctx->replicas, tile_floats, early_look, ctx->k3_legacy ? 0 : 1);
			break;
	}
	ctx->launches++;
	return e;
}
