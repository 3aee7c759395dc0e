// gas_mix_stream.cu — the step kernel (sm_100a): streams the voices of block k into its bus buffers and, on the same
// SMs and in the same launch, computes the gains and the plan of block k + 1.
//
// Streaming part.  For every voice whose block needs no per-sample recurrence (no attenuation filter, no effect chain)
// the reference's per-voice work (mix_channel ramp, reference audio_spatializer_3d.cpp:600-604, the
// += into the instance mix buffer, audio_spatializer.cpp:433-434, and the AudioServer ramped bus
// accumulate, upstream _mix_step_for_channel) collapses to
//       bus[b][c][i] += (A + B t + C t^2) * x_v[i],   t = i / F
// with the polynomial coefficients prepared by the planner.  Summed over voices this is a skinny fp32
// contraction — HBM-bound: every source frame (8 bytes) is read exactly once.
//
// Structure: persistent, one CTA per SM, 16 warps, warp-specialised.
//   warp 8      producer: streams voice rows HBM -> shared memory with 1-D bulk async copies (TMA engine,
//               cp.async.bulk + mbarrier complete_tx, SASS UBLKCP) through a ring of stages; rows are
//               gathered by the class lists, so voices need not be contiguous in memory.  One copy per voice
//               row plus one per stage for the weight records: tools/streamtest.cu measured this pattern (4 KB
//               rows, 148 CTAs) at 4.6-4.7 TB/s for a 64 MiB pass including launch, against 4.1 TB/s for
//               per-warp cp.async rings; a bulk copy costs ~75 ns of issue whatever its size, so tiles
//               narrower than 512 frames (more, smaller copies) stream slower.
//   warps 0-7   consumers: each thread owns 2 frames of the tile and keeps one (L,R) accumulator per (row group, pair,
//               frame) in registers; per voice it evaluates the weight at its two frames, w = A + t (B + t C), and adds
//               w * x: the FMAs are packed FFMA2 (fma.rn.f32x2: one instruction per (L,R) pair), weights are 128-bit
//               shared-memory broadcasts.  (The first version accumulated sum A x, sum B x, sum C x separately and applied
//               the polynomial at the flush: the same FMA count, but 2-3 times the accumulator registers — 168 per thread,
//               which left no room on the SM for anything else.)
//   warps 9-15  control: gains (calculate_spatialization, gas_gain.cuh) and plan (gas_plan.cuh) of the NEXT block, two
//               lanes per emitter / voice: latency-bound dependent chains that need few issue slots and no bandwidth,
//               i.e. exactly what the streaming warps leave idle.  As separate kernels they cost the step 12 + 14 us
//               beside 21 us of streaming; here they hide behind it.
// A CTA owns a contiguous range of (class, frame tile, voice batch) units; ranges are cut in cost space
// (weight count + a fixed per-voice term).  When the class or tile changes the CTA adds its partial sums to the bus
// buffers with per-thread red.global.add.v4.f32 straight out of the accumulator registers.
//
// Chaining.  Block indices live on the device (BLK_S: launches of this kernel so far; BLK_P: plans produced), so a
// replayed CUDA graph needs no host-side state.  Launched with programmatic stream serialization, the next step's CTAs
// start on an SM as soon as this step's CTA there has left: its streaming warps only need the plan of their block
// (PlanHdr::seq, acquire) — published by the previous launch's control warps — while its control warps first wait for
// the previous launch to be complete (griddepcontrol.wait), because they overwrite what that launch may still be reading.
//
// Compiled with -fmad=false (the control warps reproduce the reference's rounding sequence); every FMA of the
// streaming part is explicit.
#include "gas_gain.cuh"
#include "gas_plan.cuh"

#include <stdlib.h>

namespace {

constexpr int kConsumerWarps = 8;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kStreamThreads = kConsumerThreads + 32; // consumers + producer
constexpr int kControlWarps = 7;
constexpr int kControlThreads = kControlWarps * 32;
constexpr int kThreads = kStreamThreads + kControlThreads; // 512
constexpr int kTileFrames = 512;       // frames per tile: 2 per consumer thread
constexpr int kMaxPairs = GAS_K2_MAX_ROWS * GAS_MAX_CHANNELS_PER_BUS; // 24 (L,R) weights per voice
constexpr int kMaxStages = 8;

struct StreamCfg {
	int frames;        // F
	int src_stride;    // frames between consecutive source rows
	int tile_frames;   // min(F, 512)
	int n_tiles;       // ceil(F / 512)
	int slots;         // tile_frames / 2
	int groups;        // voice groups working side by side inside a stage
	int vb;            // voices per stage
	int stages;
	int x_bytes;       // per stage
	int w_bytes;       // per stage
	int stage_bytes;
	int fixed_cost;    // per-voice term of the partition cost (the other term is the weight count)
	int vb_shift;      // log2(vb)
	double inv_grid;   // 1 / CTAs
	double inv_cost[kMaxPairs + 1]; // 1 / (weights + fixed_cost)
	int debug;         // GAS_K2_DEBUG bits (experiments only): 1 = skip the bus reductions, 2 = skip the FMAs, 4 = skip the copies, 8 = record a timeline
	unsigned long long *timeline; // [CTA][16] globaltimer stamps (debug & 8), see tools/k2bench.cpp for the slots
};

// Everything one launch needs.  Streaming part: block BLK_S (planned earlier).  Control part: the next block.
struct StepArgs {
	gasplan::PlanArgs p; // tables + the next block's voice list and outputs
	StreamCfg cf;
	const gas_frame *src;
	float *bus;
	int control_on;      // 0: the control warps leave at once (one-call-per-block entry points plan with k_plan)
	int n_emitters;      // gain side of the next block (0: parameters stay)
	const gas_emitter *emitters;
	int n_listeners;
	const gas_listener *listeners;
	const ListenerPre *listener_pre;
	const gas_area *areas;
	int n_areas;
};

// ---- unit iterator: (class, frame tile, voice batch), identical in every role ------------
struct UnitIter {
	int cid, tile, batch, nb; // cid: index into the CTA's compact table of streaming classes; nb: batches of the current class
	int remaining;
	int n_cls;
};

// CTA `cta` owns the units whose start lies in [total*cta/n, total*(cta+1)/n) of the cost line, where a
// unit of class c costs w_c = accumulators + fixed_cost.  Units are ordered class, tile, batch.
// Everything is 32-bit unsigned arithmetic (64-bit divisions are long software routines and this runs on the
// critical path of every CTA's start): the cost line of the largest context (65536 voices x 4 tiles x cost <= 90)
// times the grid size stays below 2^32; beyond that the wide variant takes over.
__device__ __noinline__ UnitIter unit_iter_init_wide(const ClassInfo *cls, int n_cls, int vb, int n_tiles, int fixed_cost, int C, int cta, int n_cta) {
	// by value in and out: a reference parameter of an out-of-line function would pin the caller's iterator (and the
	// whole StreamCfg) in local memory for the rest of the kernel
	UnitIter it;
	it.remaining = 0;
	it.cid = n_cls;
	it.n_cls = n_cls;
	it.tile = it.batch = it.nb = 0;
	long long total = 0;
	for (int c = 0; c < n_cls; c++) {
		const long long units = (long long)((cls[c].count + vb - 1) / vb) * n_tiles;
		total += units * (cls[c].n_rows * C + fixed_cost);
	}
	const long long lo = total * cta / n_cta, hi = total * (cta + 1) / n_cta;
	long long base = 0;
	for (int c = 0; c < n_cls; c++) {
		const int nb = (cls[c].count + vb - 1) / vb;
		const long long units = (long long)nb * n_tiles;
		const long long w = cls[c].n_rows * C + fixed_cost;
		long long u_lo = lo <= base ? 0 : (lo - base + w - 1) / w;
		long long u_hi = hi <= base ? 0 : (hi - base + w - 1) / w;
		u_lo = u_lo > units ? units : u_lo;
		u_hi = u_hi > units ? units : u_hi;
		if (u_hi > u_lo) {
			if (it.remaining == 0) {
				it.cid = c;
				it.nb = nb;
				it.tile = (int)(u_lo / nb);
				it.batch = (int)(u_lo % nb);
			}
			it.remaining += (int)(u_hi - u_lo);
		}
		base += units * w;
	}
	return it;
}

// floor(x / d) for x < 2^32 and 1 <= d < 2^31 through one double multiply: with inv = RN(1/d) the product
// (x + 0.5) * inv is off by less than 2^-20 from (x + 0.5) / d, which never comes closer than 1 / (2 d) to an integer.
__device__ __forceinline__ unsigned div_floor_u32(unsigned x, double inv) { return (unsigned)__double2uint_rz(((double)x + 0.5) * inv); }

__device__ __forceinline__ void unit_iter_init(UnitIter &it, const ClassInfo *cls, int n_cls, const StreamCfg &cf, int C, int cta, int n_cta) {
	const int vb_shift = 31 - __clz(cf.vb); // vb is 8, 16 or 32
	unsigned long long total64 = 0;
	for (int c = 0; c < n_cls; c++) {
		const unsigned units = (unsigned)((cls[c].count + cf.vb - 1) >> vb_shift) * (unsigned)cf.n_tiles;
		total64 += (unsigned long long)units * (unsigned)(cls[c].n_rows * C + cf.fixed_cost);
	}
	it.remaining = 0;
	it.cid = n_cls;
	it.n_cls = n_cls;
	it.tile = it.batch = it.nb = 0;
	if (total64 * (unsigned)(n_cta + 1) >= (1ULL << 32)) {
		it = unit_iter_init_wide(cls, n_cls, cf.vb, cf.n_tiles, cf.fixed_cost, C, cta, n_cta);
		return;
	}
	const unsigned total = (unsigned)total64;
	const double inv_n = 1.0 / (double)n_cta;
	const unsigned lo = div_floor_u32(total * (unsigned)cta, inv_n);
	const unsigned hi = div_floor_u32(total * (unsigned)(cta + 1), inv_n);
	unsigned base = 0;
	for (int c = 0; c < n_cls; c++) {
		const unsigned nb = (unsigned)(cls[c].count + cf.vb - 1) >> vb_shift;
		const unsigned units = nb * (unsigned)cf.n_tiles;
		const unsigned w = (unsigned)(cls[c].n_rows * C + cf.fixed_cost);
		const double inv_w = 1.0 / (double)w;
		unsigned u_lo = lo <= base ? 0u : div_floor_u32(lo - base + w - 1u, inv_w);
		unsigned u_hi = hi <= base ? 0u : div_floor_u32(hi - base + w - 1u, inv_w);
		u_lo = min(u_lo, units);
		u_hi = min(u_hi, units);
		if (u_hi > u_lo) {
			if (it.remaining == 0) {
				it.cid = c;
				it.nb = (int)nb;
				if (cf.n_tiles == 1) {
					it.tile = 0;
					it.batch = (int)u_lo;
				} else {
					it.tile = (int)(u_lo / nb);
					it.batch = (int)(u_lo - (u_lo / nb) * nb);
				}
			}
			it.remaining += (int)(u_hi - u_lo);
		}
		base += units * w;
	}
}

// The same partition computed by one warp, one class per lane (up to 32 streaming classes; more fall back to the
// loop above): class costs -> warp scan -> this CTA's [lo, hi) -> per-class unit ranges -> first class and unit count
// by ballot / warp reduction.  Reciprocals come from the host (StreamCfg).  Every lane returns the same iterator.
__device__ __forceinline__ void unit_iter_init_warp(UnitIter &it, const ClassInfo *cls, int n_cls, const StreamCfg &cf, int C, int cta, int n_cta, int lane) {
	if (n_cls > 32) {
		unit_iter_init(it, cls, n_cls, cf, C, cta, n_cta);
		return;
	}
	unsigned nb = 0, units = 0, w = 1;
	int np = 0;
	if (lane < n_cls) {
		nb = (unsigned)(cls[lane].count + cf.vb - 1) >> cf.vb_shift;
		units = nb * (unsigned)cf.n_tiles;
		np = cls[lane].n_rows * C;
		w = (unsigned)(np + cf.fixed_cost);
	}
	const unsigned cost = units * w; // <= 32768 units x ~90 per class: the sum over 32 classes fits 32 bits
	unsigned incl = cost;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const unsigned o = __shfl_up_sync(0xffffffffu, incl, d);
		incl += lane >= d ? o : 0u;
	}
	const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
	it.remaining = 0;
	it.cid = n_cls;
	it.n_cls = n_cls;
	it.tile = it.batch = it.nb = 0;
	if ((unsigned long long)total * (unsigned)(n_cta + 1) >= (1ULL << 32)) {
		it = unit_iter_init_wide(cls, n_cls, cf.vb, cf.n_tiles, cf.fixed_cost, C, cta, n_cta);
		return;
	}
	const unsigned base = incl - cost;
	const unsigned lo = div_floor_u32(total * (unsigned)cta, cf.inv_grid);
	const unsigned hi = div_floor_u32(total * (unsigned)(cta + 1), cf.inv_grid);
	const double inv_w = cf.inv_cost[np];
	unsigned u_lo = lo <= base ? 0u : div_floor_u32(lo - base + w - 1u, inv_w);
	unsigned u_hi = hi <= base ? 0u : div_floor_u32(hi - base + w - 1u, inv_w);
	u_lo = min(u_lo, units);
	u_hi = min(u_hi, units);
	const unsigned mine = u_hi > u_lo ? u_hi - u_lo : 0u;
	const unsigned have = __ballot_sync(0xffffffffu, mine > 0u);
	if (have == 0u) {
		return;
	}
	const int first = __ffs(have) - 1;
	it.remaining = (int)__reduce_add_sync(0xffffffffu, mine);
	it.cid = first;
	const unsigned nb0 = __shfl_sync(0xffffffffu, nb, first);
	const unsigned u0 = __shfl_sync(0xffffffffu, u_lo, first);
	it.nb = (int)nb0;
	if (cf.n_tiles == 1) {
		it.batch = (int)u0;
	} else {
		it.tile = (int)(u0 / nb0);
		it.batch = (int)(u0 - (u0 / nb0) * nb0);
	}
}

__device__ __forceinline__ void unit_iter_next(UnitIter &it, const ClassInfo *cls, const StreamCfg &cf) {
	it.remaining--;
	if (it.remaining <= 0) {
		return;
	}
	if (++it.batch < it.nb) {
		return;
	}
	it.batch = 0;
	if (++it.tile < cf.n_tiles) {
		return;
	}
	it.tile = 0;
	if (it.cid + 1 < it.n_cls) {
		it.cid++;
		it.nb = (cls[it.cid].count + cf.vb - 1) >> cf.vb_shift;
		return;
	}
	it.remaining = 0;
}

// One warp-wide load of source-row indices: 32 consecutive list positions = 32/vb consecutive units of
// one (class, tile).  `val` is the index of list position (first batch * vb + lane).
struct IdxBlock {
	int first_seq; // sequence number (within this CTA) of the first unit covered
	int n_units;
	int2 val; // {call-order index, source row}
};

__device__ __forceinline__ IdxBlock idx_block_load(UnitIter &pf, const ClassInfo *cls, const StreamCfg &cf, const int2 *__restrict__ list,
		int maxv, int lane, int seq) {
	IdxBlock b;
	b.first_seq = seq;
	b.n_units = 0;
	b.val = make_int2(0, 0);
	if (pf.remaining <= 0) {
		return b;
	}
	const int upb = 32 >> cf.vb_shift; // vb is 8, 16 or 32
	const int n = min(upb, min(pf.nb - pf.batch, pf.remaining));
	const int pos = pf.batch * cf.vb + lane;
	if (lane < n * cf.vb && pos < cls[pf.cid].count) {
		b.val = __ldcg(list + (size_t)cls[pf.cid].slot * maxv + pos);
	}
	b.n_units = n;
	// advance by n units: they all lie in the current (class, tile) row of batches
	pf.remaining -= n - 1;
	pf.batch += n - 1;
	unit_iter_next(pf, cls, cf);
	return b;
}

struct ConsumerCtx {
	unsigned char *smem;    // stage ring
	uint64_t *full;
	uint64_t *empty;
	const ClassInfo *cls;
	int tid, lane, slot, group;
	bool worker;
	int stage;
	uint32_t phase;
	unsigned long long *tl;
	bool tl_first;
};

// The FMAs of one stage for one thread: `n` voices spaced `step` apart (rows `row_bytes` apart starting at xp, weight
// records `w_stride` bytes apart starting at wp).  E = row groups x pairs: element e of a record is {A_L, A_R, B_L, B_R} at
// e * 16, its t^2 part {C_L, C_R} at E * 16 + e * 8.  The loads of several voices are issued before their FMAs: two consumer
// warps per scheduler cannot hide the shared-memory latency of one voice at a time.  STATIC: step 1 and 4 KB rows.
template <int E, bool QUAD, bool STATIC>
__device__ __forceinline__ void voice_loop(float2 (&acc)[E][2], const unsigned char *xp, const unsigned char *wp, int n, int step, int row_bytes,
		int w_stride, const float2 T0, const float2 T1) {
	constexpr int U = QUAD ? (E <= 2 ? 4 : (E <= 4 ? 2 : 1)) : (E <= 4 ? 4 : (E <= 8 ? 2 : 1));
	const int xs = STATIC ? kTileFrames * 8 : step * row_bytes; // bytes between consecutive voices of this thread
	const int ws = step * w_stride;
	int left = STATIC ? n : (n + step - 1) / step; // voices this thread still has to do
	for (; left >= U; left -= U, xp += U * xs, wp += U * ws) {
		float4 x[U];
		float4 w[U][E];
		float2 qd[U][QUAD ? E : 1];
#pragma unroll
		for (int u = 0; u < U; u++) {
			x[u] = *reinterpret_cast<const float4 *>(xp + u * xs);
			const unsigned char *wv = wp + u * ws;
#pragma unroll
			for (int e = 0; e < E; e++) {
				w[u][e] = *reinterpret_cast<const float4 *>(wv + e * 16);
				if (QUAD) {
					qd[u][e] = *reinterpret_cast<const float2 *>(wv + E * 16 + e * 8);
				}
			}
		}
#pragma unroll
		for (int u = 0; u < U; u++) {
			const float2 x0 = make_float2(x[u].x, x[u].y), x1 = make_float2(x[u].z, x[u].w);
#pragma unroll
			for (int e = 0; e < E; e++) {
				const float2 a = make_float2(w[u][e].x, w[u][e].y);
				float2 b0 = make_float2(w[u][e].z, w[u][e].w), b1 = b0;
				if (QUAD) {
					b0 = gas_ffma2(qd[u][e], T0, b0);
					b1 = gas_ffma2(qd[u][e], T1, b1);
				}
				const float2 w0 = gas_ffma2(b0, T0, a), w1 = gas_ffma2(b1, T1, a);
				acc[e][0] = gas_ffma2(w0, x0, acc[e][0]);
				acc[e][1] = gas_ffma2(w1, x1, acc[e][1]);
			}
		}
	}
	for (; left > 0; left--, xp += xs, wp += ws) { // remainder
		const float4 x = *reinterpret_cast<const float4 *>(xp);
		const float2 x0 = make_float2(x.x, x.y), x1 = make_float2(x.z, x.w);
#pragma unroll
		for (int e = 0; e < E; e++) {
			const float4 w = *reinterpret_cast<const float4 *>(wp + e * 16);
			const float2 a = make_float2(w.x, w.y);
			float2 b0 = make_float2(w.z, w.w), b1 = b0;
			if (QUAD) {
				const float2 qd = *reinterpret_cast<const float2 *>(wp + E * 16 + e * 8);
				b0 = gas_ffma2(qd, T0, b0);
				b1 = gas_ffma2(qd, T1, b1);
			}
			const float2 w0 = gas_ffma2(b0, T0, a), w1 = gas_ffma2(b1, T1, a);
			acc[e][0] = gas_ffma2(w0, x0, acc[e][0]);
			acc[e][1] = gas_ffma2(w1, x1, acc[e][1]);
		}
	}
}

// One run = every consecutive unit of this CTA that shares (class, tile).  NG row groups x C channel pairs accumulators
// per frame, in registers for the whole run; at its end the partial sums are added to the bus buffers.
template <int NG, int C, bool QUAD>
__device__ __forceinline__ void consumer_run(UnitIter &it, ConsumerCtx &cc, const StreamCfg &cf, float *__restrict__ bus) {
	constexpr int E = NG * C;
	float2 acc[E][2];
#pragma unroll
	for (int e = 0; e < E; e++) {
		acc[e][0] = make_float2(0.f, 0.f);
		acc[e][1] = make_float2(0.f, 0.f);
	}
	const int cid = it.cid, tile = it.tile;
	const ClassInfo &ci = cc.cls[cid];
	const int tile_w = min(cf.tile_frames, cf.frames - tile * cf.tile_frames);
	const int row_bytes = tile_w * 8;
	const int w_stride = cls_row_floats(NG, QUAD ? 1 : 0, C) * 4;
	const bool mine = cc.worker && cc.slot * 2 < tile_w;
	const int F = cf.frames;
	const int frame0 = tile * cf.tile_frames + cc.slot * 2;
	const float t0 = (float)frame0 / (float)F;
	const float t1 = (float)(frame0 + 1) / (float)F;
	const float2 T0 = make_float2(t0, t0), T1 = make_float2(t1, t1);
	do {
		const int v0 = it.batch * cf.vb;
		const int nv = min(cf.vb, ci.count - v0);
		const unsigned char *sx = cc.smem + (size_t)cc.stage * cf.stage_bytes;
		const unsigned char *sw = sx + cf.x_bytes;
		gas_mbar_wait(&cc.full[cc.stage], cc.phase);
		if (cc.tl && !cc.tl_first) { // (a register flag: reading the stamp back from global memory every stage cost 0.3 us per stage)
			cc.tl[2] = gas_globaltimer();
			cc.tl_first = true;
		}
		if (mine && !(cf.debug & 2)) {
			// full 512-frame tiles (one voice group, 4 KB rows) take the loop whose row stride is a compile-time constant
			if (cf.groups == 1 && row_bytes == kTileFrames * 8) {
				voice_loop<E, QUAD, true>(acc, sx + cc.slot * 16, sw, nv, 1, kTileFrames * 8, w_stride, T0, T1);
			} else {
				voice_loop<E, QUAD, false>(acc, sx + (size_t)cc.group * row_bytes + cc.slot * 16, sw + cc.group * w_stride, nv - cc.group, cf.groups,
						row_bytes, w_stride, T0, T1);
			}
		}
		__syncwarp();
		if (cc.lane == 0) {
			gas_mbar_arrive(&cc.empty[cc.stage]);
		}
		if (++cc.stage == cf.stages) {
			cc.stage = 0;
			cc.phase ^= 1u;
		}
		unit_iter_next(it, cc.cls, cf);
	} while (it.remaining > 0 && it.cid == cid && it.tile == tile);
	if (cc.tl) {
		cc.tl[3] = gas_globaltimer();
	}
	if ((cf.debug & 1) || !mine) {
		return;
	}
	// ---- flush: bus[b][c][i] += the sums of this thread's two frames ------------------------------------------------
	if (cc.tl) {
		cc.tl[9] = gas_globaltimer();
	}
	if (ci.flags & (CLS_SHARED | CLS_SCALED)) {
		// one row group fanned out to every bus of the mask: as it is (shared), or times the send's scale (scaled)
		uint32_t m = ci.mask;
		float sc = 1.f, nxt = (ci.flags & CLS_SCALED) ? ci.scale[0] : 1.f;
		const float nxt2 = (ci.flags & CLS_SCALED) ? ci.scale[1] : 1.f;
		while (m) {
			const int b = __ffs(m) - 1;
			m &= m - 1;
#pragma unroll
			for (int c = 0; c < C; c++) {
				gas_red_add_v4(bus + ((size_t)(b * C + c) * F + frame0) * 2, acc[c][0].x * sc, acc[c][0].y * sc, acc[c][1].x * sc, acc[c][1].y * sc);
			}
			sc = nxt;
			nxt = nxt2;
		}
	} else {
		uint32_t rest = ci.mask;
#pragma unroll
		for (int k = 0; k < NG; k++) {
			const int b = __ffs(rest) - 1; // k-th bus of the mask (sends ascend by bus)
			rest &= rest - 1;
#pragma unroll
			for (int c = 0; c < C; c++) {
				gas_red_add_v4(bus + ((size_t)(b * C + c) * F + frame0) * 2, acc[k * C + c][0].x, acc[k * C + c][0].y, acc[k * C + c][1].x, acc[k * C + c][1].y);
			}
		}
	}
}

__global__ void __launch_bounds__(kThreads, 1) k_step(const __grid_constant__ StepArgs A) {
	GAS_DYN_SMEM(unsigned char, 128, smem);
	__shared__ ClassInfo s_cls[GAS_MAX_CLASSES];
	__shared__ UnitIter s_it;
	__shared__ __align__(8) uint64_t s_full[kMaxStages];
	__shared__ __align__(8) uint64_t s_empty[kMaxStages];
	__shared__ gasplan::PlanSmem s_plan;

	const int tid = threadIdx.x;
	const int warp = tid >> 5, lane = tid & 31;
	const StreamCfg &cf = A.cf;
	const GlobalCfg &g = A.p.g;
	const BlockPlan &plan = A.p.plan;
	int32_t *blk = A.p.t.blk;
	const int C = g.channels;
	const int maxv = g.max_voices;

	if (warp > kConsumerWarps) {
		// ===== control warps: gains and plan of the next block =====
		if (!A.control_on) {
			return;
		}
		const int ctl_tid = tid - kStreamThreads;
		gasplan::PlanGroup G;
		G.tl = (cf.debug & 8) && ctl_tid == 0 ? cf.timeline + blockIdx.x * 32 : nullptr;
		gasplan::group_stamp(G, 16);
		// the previous launch is complete after this (its streaming warps read the plan slot and the bus buffers that
		// are rewritten here; its control warps wrote the tables that are read here)
		GAS_GRID_DEP_WAIT();
		int wait_total = 0, own_q = -1;
		G.tid = ctl_tid;
		G.nthreads = kControlThreads;
		G.cta = blockIdx.x;
		G.n_cta = gridDim.x;
		G.bar_id = 3;
		// this lane pair's first voice of the next block: fetched now, needed after the gain tasks
		gas_voice v_pre{};
		v_pre.voice = -1;
		if (gasplan::plan_first_voice(G) < A.p.n_voices) {
			v_pre = A.p.voices[gasplan::plan_first_voice(G)];
		}
		if (A.n_emitters > 0) {
			// Gains of the next block, two lanes per emitter.  No grid-wide barrier separates them from the plan: every instance
			// carries a flag (block index + 1, release) that the planner of a voice waits for, and every control warp reports
			// when its gain tasks are done, which releases the voices of instances that got no emitter this block.
			const int b_next = gasplan::ld_volatile(&blk[BLK_P]);
			const int pairs = kControlThreads / 2;
			const int gbase = lane & 30;
			const unsigned gm = 3u << gbase;
			for (int e = blockIdx.x * pairs + (ctl_tid >> 1); e < A.n_emitters; e += gridDim.x * pairs) {
				own_q = gasgain::gain_emitter<2>(A.p.t, g, e, ctl_tid & 1, gbase, gm, A.emitters, A.n_listeners, A.listeners, A.listener_pre, A.areas, A.n_areas,
						nullptr, A.p.t.inst_seq, b_next + 1, G.tl);
			}
			__syncwarp();
			if (lane == 0) {
				gas_red_release_gpu_add_s32(&blk[BLK_GAIN_DONE], 1);
			}
			wait_total = (int)gridDim.x * kControlWarps;
			gasplan::group_stamp(G, 17);
		}
		gasplan::group_stamp(G, 18);
		gasplan::plan_block(G, s_plan, A.p, wait_total, own_q, &v_pre);
		return;
	}

	// timeline (debug & 8): the first lane of the producer warp stamps the start-up, thread 0 the consumer side
	unsigned long long *tlp = (cf.debug & 8) && tid == kConsumerThreads ? cf.timeline + blockIdx.x * 32 : nullptr;
	unsigned long long *tl = (cf.debug & 8) && tid == 0 ? cf.timeline + blockIdx.x * 32 : nullptr;

	// Start-up runs on the producer warp alone, with nothing but its own dependent loads in the way of the first
	// copy: block index -> plan header (class table) -> partition -> first source-row indices -> copies.  The consumer
	// warps wait on a named barrier for the table and their iterator; they have nothing to do before data lands anyway.
	UnitIter it;
	int slot_p = 0, k = 0;
	if (warp == kConsumerWarps) {
		if (tlp) {
			tlp[0] = gas_globaltimer();
			tlp[6] = gas_smid();
		}
		// Which block: every CTA reads the launch counter, then takes a ticket; the CTA with the last ticket advances the
		// counter, and only then lets the dependent launch go (a dependent launch starts once EVERY CTA of this one has
		// triggered, so all of its CTAs read the advanced counter and none of this launch's reads a value later than its own).
		k = gasplan::ld_volatile(&blk[BLK_S]);
		// (every lane keeps lane 0's reading: the warp loads before lane 0 takes the ticket, but nothing orders the OTHER lanes'
		// loads before the counter's advance except convergent execution; the shuffle makes it explicit)
		k = __shfl_sync(0xffffffffu, k, 0);
		if (lane == 0) {
			const int ticket = atomicAdd(&blk[BLK_S_TICKET], 1);
			if (ticket == (int)gridDim.x - 1) {
				blk[BLK_S_TICKET] = 0;
				*(volatile int32_t *)&blk[BLK_S] = k + 1;
				__threadfence();
			}
			for (int s = 0; s < cf.stages; s++) {
				gas_mbar_init(&s_full[s], 1);
				gas_mbar_init(&s_empty[s], kConsumerWarps);
			}
			gas_mbar_init_fence();
			GAS_GRID_DEP_LAUNCH();
		}
		slot_p = k & (GAS_PLAN_DEPTH - 1);
		const PlanHdr *hdr = &plan.hdr[slot_p];
		// the plan of this block: published by the planner's last CTA (normally long before this launch starts)
		while (gasplan::ld_acquire(&hdr->seq) != k + 1) {
			__nanosleep(100);
		}
		// Compact table of the streaming classes of this block, in slot order (identical in every CTA): one round of loads
		// of the slot table and of this block's counts.
		int n_cls = 0;
		{
			constexpr int R = GAS_MAX_CLASSES / 32;
			unsigned long long key[R], auxw[R];
			int cnt[R];
			const int32_t *counts = plan.cls_count + slot_p * GAS_MAX_CLASSES;
#pragma unroll
			for (int r = 0; r < R; r++) {
				key[r] = __ldcg(plan.cls_key + r * 32 + lane);
				auxw[r] = __ldcg(plan.cls_aux + r * 32 + lane);
				cnt[r] = __ldcg(counts + r * 32 + lane);
			}
#pragma unroll
			for (int r = 0; r < R; r++) {
				const bool on = key[r] != 0ULL && (int)(key[r] & 3u) == PATH_STREAM && cnt[r] > 0;
				const unsigned m = __ballot_sync(0xffffffffu, on);
				if (on) {
					ClassInfo ci = cls_decode(key[r], cnt[r]);
					ci.slot = r * 32 + lane;
					if (ci.flags & CLS_SCALED) {
						ci.scale[0] = __uint_as_float((unsigned)(auxw[r] & 0xffffffffu));
						ci.scale[1] = __uint_as_float((unsigned)(auxw[r] >> 32));
					}
					s_cls[n_cls + __popc(m & ((1u << lane) - 1u))] = ci;
				}
				n_cls += __popc(m);
			}
		}
		__syncwarp();
		if (tlp) {
			tlp[11] = gas_globaltimer();
		}
		unit_iter_init_warp(it, s_cls, n_cls, cf, C, blockIdx.x, gridDim.x, lane);
		if (lane == 0) {
			s_it = it;
		}
		__syncwarp();
		GAS_BAR_ARRIVE_IMM(2, kStreamThreads);
		if (tlp) {
			tlp[1] = gas_globaltimer();
			tlp[5] = (unsigned long long)it.remaining;
		}
	} else {
		GAS_BAR_SYNC_IMM(2, kStreamThreads);
		it = s_it;
	}
	if (it.remaining <= 0) {
		return;
	}
	unsigned char *ring = smem;

	if (warp == kConsumerWarps) {
		// ===== producer =====
		int stage = 0;
		uint32_t phase = 0;
		const uint64_t pol_stream = gas_l2_policy_evict_first();
		const int2 *list = plan.list + (size_t)slot_p * GAS_MAX_CLASSES * maxv;
		const float *rows = plan_rows(plan, k, 0, maxv);
		// Source-row indices are fetched one warp-wide load (32 list positions = 32/vb units) at a time,
		// two blocks ahead of the copies that need them, so the gather indirection never stalls the ring.
		UnitIter pf = it;
		IdxBlock cur = idx_block_load(pf, s_cls, cf, list, maxv, lane, 0);
		IdxBlock nxt = idx_block_load(pf, s_cls, cf, list, maxv, lane, cur.n_units);
		int seq = 0;
		if (tlp) {
			tlp[12] = gas_globaltimer() + (cur.val.y == -12345 ? 1ULL : 0ULL); // first indices have arrived
		}
		while (it.remaining > 0) {
			if (seq >= cur.first_seq + cur.n_units) {
				cur = nxt;
				nxt = idx_block_load(pf, s_cls, cf, list, maxv, lane, cur.first_seq + cur.n_units);
			}
			const ClassInfo &ci = s_cls[it.cid];
			const int v0 = it.batch * cf.vb;
			const int nv = min(cf.vb, ci.count - v0);
			const int tile_w = min(cf.tile_frames, cf.frames - it.tile * cf.tile_frames); // frames in this tile
			const uint32_t row_bytes = (uint32_t)tile_w * 8u;
			const int nf = cls_row_floats(ci.n_group, (int)ci.quad, C); // floats of weights per voice
			const uint32_t w_bytes = (uint32_t)(nv * nf * 4);
			unsigned char *sx = ring + (size_t)stage * cf.stage_bytes;
			unsigned char *sw = sx + cf.x_bytes;
			gas_mbar_wait(&s_empty[stage], phase ^ 1u);
			const bool no_copy = (cf.debug & 4) != 0; // experiment: arm the stage without copying anything into it
			if (lane == 0) {
				gas_mbar_arrive_expect_tx(&s_full[stage], no_copy ? 0u : row_bytes * (uint32_t)nv + w_bytes);
				if (!no_copy) {
					gas_bulk_g2s(sw, rows + (size_t)ci.slot * maxv * GAS_K2_ROW_FLOATS + (size_t)v0 * nf, w_bytes, &s_full[stage]);
				}
			}
			__syncwarp();
			{
				const int v = lane - (seq - cur.first_seq) * cf.vb; // this lane's voice inside the stage
				if (v >= 0 && v < nv && !no_copy) {
					gas_bulk_g2s_hint(sx + (size_t)v * row_bytes, A.src + (size_t)cur.val.y * cf.src_stride + (size_t)it.tile * cf.tile_frames, row_bytes,
							&s_full[stage], pol_stream);
				}
			}
			if (tlp && seq < 2) {
				tlp[13 + seq] = gas_globaltimer(); // copies of the first / second stage issued
			}
			seq++;
			if (++stage == cf.stages) {
				stage = 0;
				phase ^= 1u;
			}
			unit_iter_next(it, s_cls, cf);
		}
	} else {
		// ===== consumers =====
		ConsumerCtx cc;
		cc.smem = ring;
		cc.tid = tid;
		cc.tl = tl;
		cc.tl_first = false;
		cc.full = s_full;
		cc.empty = s_empty;
		cc.cls = s_cls;
		cc.lane = lane;
		cc.slot = tid % cf.slots;
		cc.group = tid / cf.slots;
		cc.worker = cc.group < cf.groups;
		cc.stage = 0;
		cc.phase = 0;
		float *bus = A.bus;
		while (it.remaining > 0) {
			// dispatch on (row groups, quadratic, channel pairs): groups in 1..3 (quadratic: 1..2), pairs in 1..4
			const ClassInfo &ci = s_cls[it.cid];
			const int code = (ci.n_group * 2 + (int)ci.quad) * 8 + C;
#define GAS_RUN(G_, Q_, C_) \
	case ((G_) * 2 + (Q_)) * 8 + (C_): consumer_run<G_, C_, (Q_) != 0>(it, cc, cf, bus); break;
			switch (code) {
				GAS_RUN(1, 0, 1) GAS_RUN(1, 0, 2) GAS_RUN(1, 0, 3) GAS_RUN(1, 0, 4)
				GAS_RUN(1, 1, 1) GAS_RUN(1, 1, 2) GAS_RUN(1, 1, 3) GAS_RUN(1, 1, 4)
				GAS_RUN(2, 0, 1) GAS_RUN(2, 0, 2) GAS_RUN(2, 0, 3) GAS_RUN(2, 0, 4)
				GAS_RUN(2, 1, 1) GAS_RUN(2, 1, 2) GAS_RUN(2, 1, 3) GAS_RUN(2, 1, 4)
				GAS_RUN(3, 0, 1) GAS_RUN(3, 0, 2) GAS_RUN(3, 0, 3) GAS_RUN(3, 0, 4)
				default: // cannot happen; drain so the producer never stalls
					consumer_run<1, 1, false>(it, cc, cf, bus);
					break;
			}
#undef GAS_RUN
		}
		if (tl) {
			tl[4] = gas_globaltimer();
		}
	}
}

} // namespace

static int env_int(const char *name, int dflt, int lo, int hi) {
	const char *e = getenv(name);
	if (!e || !*e) {
		return dflt;
	}
	const int v = atoi(e);
	return v < lo ? lo : (v > hi ? hi : v);
}

static StreamCfg make_cfg(int frames, int src_stride, int smem_limit, int n_cta) {
	StreamCfg cf{};
	cf.frames = frames;
	cf.src_stride = src_stride;
	// Frame tile: 512 (4 KB copies) streams fastest; narrower tiles (GAS_K2_TILE) trade copy size for less
	// reduction traffic into the bus buffers.
	int tile = env_int("GAS_K2_TILE", kTileFrames, 64, kTileFrames);
	tile = tile >= 512 ? 512 : (tile >= 256 ? 256 : (tile >= 128 ? 128 : 64));
	cf.tile_frames = frames < tile ? frames : tile;
	cf.n_tiles = (frames + cf.tile_frames - 1) / cf.tile_frames;
	cf.slots = cf.tile_frames / 2;
	cf.groups = kConsumerThreads / cf.slots;
	if (cf.groups < 1) {
		cf.groups = 1;
	}
	int vb = 32768 / (cf.tile_frames * 8); // 32 KB stages (64 KB stages, i.e. 16 voices: measured +1.3 us per block, coarser partition)
	vb = vb >= 32 ? 32 : (vb >= 16 ? 16 : 8); // a divisor of the warp size (index prefetch blocks)
	cf.vb = vb;
	cf.x_bytes = vb * cf.tile_frames * 8;
	cf.w_bytes = (vb * kMaxPairs * 8 + 127) & ~127;
	cf.stage_bytes = cf.x_bytes + cf.w_bytes;
	int stages = smem_limit / cf.stage_bytes;
	cf.stages = stages > kMaxStages ? kMaxStages : stages;
	// tuning / experiment knobs (environment, read per launch: they never change results, only schedules)
	cf.stages = env_int("GAS_K2_STAGES", cf.stages, 2, cf.stages);
	cf.fixed_cost = env_int("GAS_K2_FIXED_COST", 40, 0, 1024);
	cf.debug = env_int("GAS_K2_DEBUG", 0, 0, 15);
	cf.vb_shift = vb >= 32 ? 5 : (vb >= 16 ? 4 : 3);
	cf.inv_grid = 1.0 / (double)n_cta;
	for (int np = 0; np <= kMaxPairs; np++) {
		cf.inv_cost[np] = 1.0 / (double)(np + cf.fixed_cost > 0 ? np + cf.fixed_cost : 1);
	}
	return cf;
}

// One launch of the step kernel on `st`: streams the planned block (src rows -> d_bus) and, when `next` is given, computes
// the gains (next->n_emitters > 0) and the plan of the next block on its control warps.
cudaError_t launch_step(gas_ctx *ctx, const gas_frame *d_src, int src_stride, int frames, gas_frame *d_bus, const StepNext *next, cudaStream_t st,
		bool pdl) {
	static const int kSmemLimit = 206 * 1024;
	StepArgs A{};
	A.cf = make_cfg(frames, src_stride, kSmemLimit, ctx->num_sms);
	if (A.cf.stages < 2) {
		return cudaErrorInvalidConfiguration;
	}
	const size_t smem = (size_t)A.cf.stages * A.cf.stage_bytes;
	if (!ctx->k2_smem_attr_set) {
		cudaError_t e = cudaFuncSetAttribute(k_step, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
		if (e != cudaSuccess) {
			return e;
		}
		ctx->k2_smem_attr_set = true;
	}
	if (A.cf.debug & 8) {
		if (!ctx->d_timeline) {
			cudaMalloc((void **)&ctx->d_timeline, 256 * 32 * sizeof(unsigned long long));
		}
		cudaMemsetAsync(ctx->d_timeline, 0, 256 * 32 * sizeof(unsigned long long), st);
		A.cf.timeline = ctx->d_timeline;
	}
	A.p.t = ctx->t;
	A.p.g = ctx->g;
	A.p.plan = ctx->plan;
	A.src = d_src;
	A.bus = (float *)d_bus;
	if (next) {
		A.control_on = 1;
		A.p.inst_hwm = ctx->inst_hwm;
		A.p.n_voices = next->n_voices;
		A.p.voices = next->d_voices;
		A.p.src_rows = next->src_rows;
		A.p.bus = (float4 *)next->d_bus;
		A.p.bus_f4 = gas_bus_f4(ctx, next->frames);
		A.p.peaks = (float2 *)next->d_peaks;
		A.p.scaled_classes = ctx->scaled_classes ? 1 : 0;
		A.n_emitters = next->n_emitters;
		A.emitters = next->d_emitters;
		A.n_listeners = ctx->n_listeners_res;
		A.listeners = ctx->d_listeners;
		A.listener_pre = ctx->d_listener_pre;
		A.areas = ctx->d_areas;
		A.n_areas = ctx->n_areas_res;
	}
	cudaError_t e = gas_launch(k_step, dim3(ctx->num_sms), dim3(kThreads), smem, st, pdl, A);
	ctx->launches++;
	return e;
}
