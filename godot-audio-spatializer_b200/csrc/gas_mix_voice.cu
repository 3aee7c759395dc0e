// gas_mix_voice.cu — K3: the voice-parallel mix kernel (sm_100a).
//
// Voices whose block contains a per-sample recurrence — the attenuation high-shelf biquad of
// process_frames / mix_channel (reference audio_spatializer_3d.cpp:503-529, :568-597, upstream
// AudioFilterSW::Processor) or an AudioSpatializerEffect filter chain (reference
// audio_spatializer_effect.cpp:33-77) — and voices that must report a block peak
// (reference audio_spatializer.cpp:419-461) run here, serial in time with the filter state in registers for
// the whole block.  The parallel axis is the *stream*, not the voice:
//   Mode B   one lane per (voice, pair, side): the voice's 2C biquads run side by side, 4 voices per warp;
//   Mode A / effect chains   one lane per (voice, side), 16 voices per warp.
// After the per-stream work the AudioServer ramp (upstream _mix_step_for_channel) is applied per lane, the
// voices of the warp are summed with shuffles and the result is added to a CTA-wide accumulation tile in
// shared memory (the whole bus layout of the block, explicit red.shared); the CTA adds its tile to the bus
// buffers once, with one vector reduction per 16 bytes.  The launch also folds the streaming kernel's partial
// sums (replicas) into the bus buffers, so it runs every block.
//
// Sends are processed two at a time; a voice with more than two sends (bus transitions) is run in
// several passes from the same initial state — every pass recomputes bit-identical samples, only the
// last one stores state and peaks.
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

// units of a class: Mode B runs 4 voices per warp, Mode A / effect chains 16; the generic class (voices whose own class
// found no slot) runs one voice per unit, because its voices do not share a send layout
__device__ __forceinline__ int units_of(const ClassInfo &ci) {
	if (ci.flags & CLS_GENERIC) {
		return ci.count;
	}
	return ci.mode == MODE_B ? (ci.count + 3) / 4 : (ci.count + 15) / 16;
}

template <int C>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 2) k_mix_voice(DevTables t, GlobalCfg g, BlockPlan plan,
		const gas_frame *__restrict__ src, int src_stride, int F, float *__restrict__ bus, float2 *__restrict__ peaks,
		const float4 *__restrict__ rep, int bus_f4, int replicas, int tile_floats, int early_look) {
	GAS_DYN_SMEM(float, 16, s_tile);
	__shared__ ClassInfo s_cls[GAS_MAX_CLASSES];
	__shared__ int s_ncls;
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
	bool cta_has_work = false;
	{
		int units = 0;
		for (int c = 0; c < s_ncls; c++) {
			units += units_of(s_cls[c]);
		}
		cta_has_work = units > (int)blockIdx.x; // units are dealt round-robin to CTAs, then to the warps of a CTA
	}
	if (tile && cta_has_work) {
		for (int i = threadIdx.x; i < tile_floats / 4; i += blockDim.x) {
			reinterpret_cast<float4 *>(s_tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
		}
		__syncthreads();
	}
	const int my_warp = threadIdx.x >> 5;
	// Units (32-voice chunks, or 4-voice chunks for the stream-parallel Mode B path) are numbered class by class and
	// dealt round-robin to the CTAs, then to the warps of a CTA: this warp owns units b + grid * (w + 8 r), r = 0, 1, ...
	int total_units = 0;
	for (int c = 0; c < s_ncls; c++) {
		total_units += units_of(s_cls[c]);
	}
	for (int unit = (int)blockIdx.x + (int)gridDim.x * my_warp; unit < total_units; unit += (int)gridDim.x * kWarpsPerCta) {
		int c = 0, k = unit;
		for (; c < s_ncls; c++) { // class and chunk of the unit: Mode B 4 voices per warp, Mode A / effect chains 16
			const int chunks = units_of(s_cls[c]);
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
	const size_t smem = (size_t)tile_floats * sizeof(float);
	if (!ctx->k3_smem_attr_set) {
		cudaFuncSetAttribute(k_mix_voice<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxTileBytes);
		cudaFuncSetAttribute(k_mix_voice<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxTileBytes);
		cudaFuncSetAttribute(k_mix_voice<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxTileBytes);
		cudaFuncSetAttribute(k_mix_voice<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxTileBytes);
		ctx->k3_smem_attr_set = true;
	}
	cudaError_t e = cudaSuccess;
	switch (ctx->g.channels) {
		case 1:
			e = gas_launch(k_mix_voice<1>, dim3(grid), dim3(threads), smem, st, pdl, ctx->t, ctx->g, ctx->plan, d_src, src_stride, frames, (float *)d_bus, (float2 *)d_peaks, (const float4 *)ctx->d_rep, bus_f4, ctx->replicas, tile_floats, early_look);
			break;
		case 2:
			e = gas_launch(k_mix_voice<2>, dim3(grid), dim3(threads), smem, st, pdl, ctx->t, ctx->g, ctx->plan, d_src, src_stride, frames, (float *)d_bus, (float2 *)d_peaks, (const float4 *)ctx->d_rep, bus_f4, ctx->replicas, tile_floats, early_look);
			break;
		case 3:
			e = gas_launch(k_mix_voice<3>, dim3(grid), dim3(threads), smem, st, pdl, ctx->t, ctx->g, ctx->plan, d_src, src_stride, frames, (float *)d_bus, (float2 *)d_peaks, (const float4 *)ctx->d_rep, bus_f4, ctx->replicas, tile_floats, early_look);
			break;
		default:
			e = gas_launch(k_mix_voice<4>, dim3(grid), dim3(threads), smem, st, pdl, ctx->t, ctx->g, ctx->plan, d_src, src_stride, frames, (float *)d_bus, (float2 *)d_peaks, (const float4 *)ctx->d_rep, bus_f4, ctx->replicas, tile_floats, early_look);
			break;
	}
	ctx->launches++;
	return e;
}
