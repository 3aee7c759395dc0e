// gas_mix_stream.cu — K2: the streaming mix kernel (sm_100a).
//
// For every voice whose block needs no per-sample recurrence (no attenuation filter, no effect chain)
// the reference's per-voice work (mix_channel ramp, reference audio_spatializer_3d.cpp:600-604, the
// += into the instance mix buffer, audio_spatializer.cpp:433-434, and the AudioServer ramped bus
// accumulate, upstream _mix_step_for_channel) collapses to
//       bus[b][c][i] += (A + B t + C t^2) * x_v[i],   t = i / F
// with the polynomial rows prepared by the prologue.  Summed over voices this is a skinny fp32
// contraction  Out[rows, F] = W[rows, V] . X[V, F]  followed by the evaluation in t — HBM-bound: every
// source frame (8 bytes) is read exactly once and feeds 2*rows FMAs.
//
// Structure: persistent, one CTA per SM, warp-specialised.
//   warp 8      producer: streams voice rows HBM -> shared memory with 1-D bulk async copies (TMA engine,
//               cp.async.bulk + mbarrier complete_tx, SASS UBLKCP) through a ring of stages; rows are
//               gathered by the class lists, so voices need not be contiguous in memory.  One copy per voice
//               row plus one per stage for the weights: tools/streamtest.cu measured this pattern (4 KB
//               rows, 148 CTAs) at 4.6-4.7 TB/s for a 64 MiB pass including launch, against 4.1 TB/s for
//               per-warp cp.async rings; a bulk copy costs ~75 ns of issue whatever its size, so tiles
//               narrower than 512 frames (more, smaller copies) stream slower.
//   warps 0-7   consumers: each thread owns 2 frames of the tile and keeps rows x 2 (L,R) accumulators in
//               registers; weights are broadcast from shared memory; the FMAs are packed FFMA2
//               (fma.rn.f32x2: one instruction per (L,R) pair).
// A CTA owns a contiguous range of (class, frame tile, voice batch) units; ranges are cut in cost space
// (accumulator count + a fixed per-voice term) so that the FMA-heavy classes do not pile up on a few SMs.
// When the class or tile changes the CTA evaluates the polynomial in t and adds its partial sums to the bus
// buffers with per-thread red.global.add.v4.f32, straight out of the accumulator registers (the row layout of a
// class is resolved by a switch over compile-time layouts: a first version parked the accumulators in a
// thread-local array and indexed it at run time, and with 216 KB of the SM's SRAM carved out as shared memory
// those local loads went to L2 one dependent round trip after another, ~3 us per flush).  Two other flush
// shapes were measured and dropped: shared-memory staging + bulk async reductions (cp.reduce.async.bulk, +4 us)
// and plain stores into per-CTA slabs folded later (no gain).  No per-voice state is written here.
#include "gas_internal.h"

#include <stdlib.h>

#ifndef GAS_USE_FFMA2
#define GAS_USE_FFMA2 1
#endif

namespace {

constexpr int kConsumerWarps = 8;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kThreads = kConsumerThreads + 32;
constexpr int kTileFrames = 512;       // frames per tile: 2 per consumer thread
constexpr int kMaxPairs = GAS_K2_MAX_ROWS * GAS_MAX_CHANNELS_PER_BUS; // 24 (L,R) weight pairs per voice
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
	int fixed_cost;    // per-voice term of the partition cost (the other term is the accumulator count)
	int vb_shift;      // log2(vb)
	double inv_grid;   // 1 / CTAs
	double inv_cost[kMaxPairs + 1]; // 1 / (accumulators + fixed_cost)
	int debug;         // GAS_K2_DEBUG bits (experiments only): 1 = skip the bus reductions, 2 = skip the FMAs, 4 = skip the copies, 8 = record a timeline
	unsigned long long *timeline; // [CTA][16] globaltimer stamps (debug & 8), see tools/k2bench.cpp for the slots
};

// ---- PTX helpers -------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long gtime() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
	uint32_t ok = 0;
	do {
		asm volatile(
				"{\n"
				".reg .pred p;\n"
				"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
				"selp.u32 %0, 1, 0, p;\n"
				"}\n"
				: "=r"(ok)
				: "r"(smem_u32(bar)), "r"(parity)
				: "memory");
	} while (!ok);
}
// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
			"l"(src), "r"(bytes), "r"(smem_u32(bar))
			: "memory");
}
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
	asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// packed (L,R) FMA: acc += w * x on both halves with one instruction (SASS: FFMA2)
__device__ __forceinline__ void fma2(float2 &acc, const float2 w, const float2 x) {
#if GAS_USE_FFMA2
	asm("{\n"
		".reg .b64 a, ww, xx;\n"
		"mov.b64 a, {%0, %1};\n"
		"mov.b64 ww, {%2, %3};\n"
		"mov.b64 xx, {%4, %5};\n"
		"fma.rn.f32x2 a, ww, xx, a;\n"
		"mov.b64 {%0, %1}, a;\n"
		"}\n"
		: "+f"(acc.x), "+f"(acc.y)
		: "f"(w.x), "f"(w.y), "f"(x.x), "f"(x.y));
#else
	acc.x = fmaf(w.x, x.x, acc.x);
	acc.y = fmaf(w.y, x.y, acc.y);
#endif
}

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
		b.val = __ldg(list + (size_t)cls[pf.cid].slot * maxv + pos);
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
	int C, tid, lane, slot, group;
	bool worker;
	int stage;
	uint32_t phase;
	unsigned long long *tl;
	bool tl_first;
};

// The FMAs of one stage for one thread: `n` voices spaced `step` apart (rows `row_bytes` apart starting at xp, weight
// records NP*8 bytes apart starting at wp).  The loads of several voices are issued before their FMAs (4 voices for
// the small classes, 2 for the large ones: the register budget): two consumer warps per scheduler cannot hide the
// shared-memory latency of one voice at a time.  STATIC: step 1 and 4 KB rows, every offset an immediate.
template <int NP, bool STATIC>
__device__ __forceinline__ void voice_loop(float2 (&acc)[NP][2], const unsigned char *xp, const unsigned char *wp, int n, int step, int row_bytes) {
	constexpr int U = NP <= 10 ? 4 : 2;
	const int xs = STATIC ? kTileFrames * 8 : step * row_bytes; // bytes between consecutive voices of this thread
	const int ws = STATIC ? NP * 8 : step * (NP * 8);
	int left = STATIC ? n : (n + step - 1) / step; // voices this thread still has to do
	for (; left >= U; left -= U, xp += U * xs, wp += U * ws) {
		float4 x[U];
		float4 w[U][(NP + 1) / 2];
#pragma unroll
		for (int u = 0; u < U; u++) {
			x[u] = *reinterpret_cast<const float4 *>(xp + u * xs);
			const unsigned char *wv = wp + u * ws;
#pragma unroll
			for (int p = 0; p < NP; p += 2) {
				if (NP % 2 == 0) {
					w[u][p / 2] = *reinterpret_cast<const float4 *>(wv + p * 8);
				} else {
					const float2 a = *reinterpret_cast<const float2 *>(wv + p * 8);
					const float2 b = p + 1 < NP ? *reinterpret_cast<const float2 *>(wv + p * 8 + 8) : make_float2(0.f, 0.f);
					w[u][p / 2] = make_float4(a.x, a.y, b.x, b.y);
				}
			}
		}
#pragma unroll
		for (int u = 0; u < U; u++) {
			const float2 x0 = make_float2(x[u].x, x[u].y), x1 = make_float2(x[u].z, x[u].w);
#pragma unroll
			for (int p = 0; p < NP; p += 2) {
				const float2 wa = make_float2(w[u][p / 2].x, w[u][p / 2].y), wb = make_float2(w[u][p / 2].z, w[u][p / 2].w);
				fma2(acc[p][0], wa, x0);
				fma2(acc[p][1], wa, x1);
				if (p + 1 < NP) {
					fma2(acc[p + 1][0], wb, x0);
					fma2(acc[p + 1][1], wb, x1);
				}
			}
		}
	}
	for (; left > 0; left--, xp += xs, wp += ws) { // remainder
		const float4 x = *reinterpret_cast<const float4 *>(xp);
		const float2 x0 = make_float2(x.x, x.y), x1 = make_float2(x.z, x.w);
#pragma unroll
		for (int p = 0; p < NP; p++) {
			const float2 wa = *reinterpret_cast<const float2 *>(wp + p * 8);
			fma2(acc[p][0], wa, x0);
			fma2(acc[p][1], wa, x1);
		}
	}
}

// One row group (rows R0, R0+1 and, when `quad`, R0+2 of a class with R rows) of one run: evaluate the
// polynomial at this thread's two frames and add the result to bus `b_own`, or to every bus of `fan` when the
// group is shared by all sends.  Every accumulator index is a compile-time constant.
template <int R, int C, int R0>
__device__ __forceinline__ void flush_group(const float2 (&acc)[R * C][2], bool quad, uint32_t fan, float sc1, float sc2, int b_own, float t0, float t1,
		float *__restrict__ bus, int F, int frame0) {
#pragma unroll
	for (int c = 0; c < C; c++) {
		const float2 a0 = acc[R0 * C + c][0], a1 = acc[R0 * C + c][1];
		const float2 b0 = acc[(R0 + 1) * C + c][0], b1 = acc[(R0 + 1) * C + c][1];
		float2 c0 = make_float2(0.f, 0.f), c1 = make_float2(0.f, 0.f);
		if (R0 + 2 < R) {
			if (quad) {
				c0 = acc[(R0 + 2 < R ? R0 + 2 : 0) * C + c][0];
				c1 = acc[(R0 + 2 < R ? R0 + 2 : 0) * C + c][1];
			}
		}
		float4 v;
		v.x = fmaf(t0, fmaf(t0, c0.x, b0.x), a0.x);
		v.y = fmaf(t0, fmaf(t0, c0.y, b0.y), a0.y);
		v.z = fmaf(t1, fmaf(t1, c1.x, b1.x), a1.x);
		v.w = fmaf(t1, fmaf(t1, c1.y, b1.y), a1.y);
		if (fan) { // one row group fanned out to every bus of the mask: as it is (shared), or times the send's scale (scaled)
			uint32_t m = fan;
			float sc = 1.f, nxt = sc1;
			while (m) {
				const int b = __ffs(m) - 1;
				m &= m - 1;
				red_add_v4(bus + ((size_t)(b * C + c) * F + frame0) * 2, v.x * sc, v.y * sc, v.z * sc, v.w * sc);
				sc = nxt;
				nxt = sc2;
			}
		} else {
			red_add_v4(bus + ((size_t)(b_own * C + c) * F + frame0) * 2, v.x, v.y, v.z, v.w);
		}
	}
}

// One run = every consecutive unit of this CTA that shares (class, tile).  R rows x C channel pairs = NP (L,R)
// weight pairs per voice.  The accumulators live in registers for the whole run; at its end the polynomial
// is evaluated in t and the partial sums are added to the bus buffers.
template <int R, int C>
__device__ __forceinline__ void consumer_run(UnitIter &it, ConsumerCtx &cc, const StreamCfg &cf, float *__restrict__ bus) {
	constexpr int NP = R * C;
	float2 acc[NP][2];
#pragma unroll
	for (int p = 0; p < NP; p++) {
		acc[p][0] = make_float2(0.f, 0.f);
		acc[p][1] = make_float2(0.f, 0.f);
	}
	const int cid = it.cid, tile = it.tile;
	const ClassInfo &ci = cc.cls[cid];
	const int tile_w = min(cf.tile_frames, cf.frames - tile * cf.tile_frames);
	const int row_bytes = tile_w * 8;
	const bool mine = cc.worker && cc.slot * 2 < tile_w;
	do {
		const int v0 = it.batch * cf.vb;
		const int nv = min(cf.vb, ci.count - v0);
		const unsigned char *sx = cc.smem + (size_t)cc.stage * cf.stage_bytes;
		const unsigned char *sw = sx + cf.x_bytes;
		mbar_wait(&cc.full[cc.stage], cc.phase);
		if (cc.tl && !cc.tl_first) { // (a register flag: reading the stamp back from global memory every stage cost 0.3 us per stage)
			cc.tl[2] = gtime();
			cc.tl_first = true;
		}
		if (mine && !(cf.debug & 2)) {
			// full 512-frame tiles (one voice group, 4 KB rows) take the loop whose strides are compile-time constants:
			// the generic one spends as many instructions on addresses as on loads
			if (cf.groups == 1 && row_bytes == kTileFrames * 8) {
				voice_loop<NP, true>(acc, sx + cc.slot * 16, sw, nv, 1, kTileFrames * 8);
			} else {
				voice_loop<NP, false>(acc, sx + (size_t)cc.group * row_bytes + cc.slot * 16, sw + cc.group * (NP * 8), nv - cc.group, cf.groups, row_bytes);
			}
		}
		__syncwarp();
		if (cc.lane == 0) {
			mbar_arrive(&cc.empty[cc.stage]);
		}
		if (++cc.stage == cf.stages) {
			cc.stage = 0;
			cc.phase ^= 1u;
		}
		unit_iter_next(it, cc.cls, cf);
	} while (it.remaining > 0 && it.cid == cid && it.tile == tile);
	if (cc.tl) {
		cc.tl[3] = gtime();
	}
	if ((cf.debug & 1) || !mine) {
		return;
	}
	// ---- flush: bus[b][c][i] += A + B t (+ C t^2), rows ordered [group][poly][pair] --------------------
	if (cc.tl) {
		cc.tl[9] = gtime();
	}
	const int F = cf.frames;
	const int frame0 = tile * cf.tile_frames + cc.slot * 2;
	const float t0 = (float)frame0 / (float)F;
	const float t1 = (float)(frame0 + 1) / (float)F;
	const uint32_t fan = (ci.flags & (CLS_SHARED | CLS_SCALED)) ? ci.mask : 0u;
	const float sc1 = (ci.flags & CLS_SCALED) ? ci.scale[0] : 1.f, sc2 = (ci.flags & CLS_SCALED) ? ci.scale[1] : 1.f;
	const int RG = ci.n_group; // row groups; group k owns 2 rows (A, B) plus a t^2 row when its quad bit is set
	uint32_t rest = ci.mask;
	int r0 = 0; // first row of group k
	for (int k = 0; k < RG; k++) {
		const bool quad = (ci.quad >> k) & 1u;
		const int b_own = __ffs(rest) - 1; // k-th bus of the mask (sends ascend by bus)
		rest &= rest - 1;
		// a group starts at row 0, 2, 3 or 4 (at most GAS_K2_MAX_ROWS = 6 rows, 2 or 3 per group)
		switch (r0) {
			case 0: flush_group<R, C, 0>(acc, quad, fan, sc1, sc2, b_own, t0, t1, bus, F, frame0); break;
			case 2:
				if (R >= 4) {
					flush_group<R, C, (R >= 4 ? 2 : 0)>(acc, quad, fan, 1.f, 1.f, b_own, t0, t1, bus, F, frame0);
				}
				break;
			case 3:
				if (R >= 5) {
					flush_group<R, C, (R >= 5 ? 3 : 0)>(acc, quad, fan, 1.f, 1.f, b_own, t0, t1, bus, F, frame0);
				}
				break;
			default:
				if (R >= 6) {
					flush_group<R, C, (R >= 6 ? 4 : 0)>(acc, quad, fan, 1.f, 1.f, b_own, t0, t1, bus, F, frame0);
				}
				break;
		}
		r0 += quad ? 3 : 2;
	}
}

__global__ void __launch_bounds__(kThreads, 1) k_mix_stream(BlockPlan plan, GlobalCfg g, StreamCfg cf,
		const gas_frame *__restrict__ src, float *__restrict__ bus, int rep_stride, int replicas, const int32_t *__restrict__ blk) {
	// flush target: replica (CTA % replicas) of the bus layout
	bus += (size_t)(blockIdx.x % replicas) * rep_stride;
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ ClassInfo s_cls[GAS_MAX_CLASSES];
	__shared__ UnitIter s_it;
	__shared__ __align__(8) uint64_t s_full[kMaxStages];
	__shared__ __align__(8) uint64_t s_empty[kMaxStages];

	const int tid = threadIdx.x;
	const int warp = tid >> 5, lane = tid & 31;
	const int C = g.channels;
	const int maxv = g.max_voices;
	// timeline (debug & 8): the first lane of the producer warp stamps the start-up, thread 0 the consumer side
	unsigned long long *tlp = (cf.debug & 8) && tid == kConsumerThreads ? cf.timeline + blockIdx.x * 16 : nullptr;
	unsigned long long *tl = (cf.debug & 8) && tid == 0 ? cf.timeline + blockIdx.x * 16 : nullptr;

	// Start-up runs on the producer warp alone, with nothing but its own dependent loads in the way of the first
	// copy: class table (one round of loads) -> partition -> first source-row indices -> copies.  The consumer warps
	// wait on a named barrier for the table and their iterator; they have nothing to do before data lands anyway.
	UnitIter it;
	if (warp == kConsumerWarps) {
		if (tlp) {
			tlp[0] = gtime();
			unsigned smid;
			asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
			tlp[6] = smid;
		}
		// programmatic dependent launch: the plan the prologue wrote is visible after this wait (a no-op when
		// launched without the attribute)
		GAS_GRID_DEP_WAIT();
		// Compact table of the streaming classes of this block, in slot order (identical in every CTA).  The
		// counts of both parities are fetched together with the block counter so that no load waits for another.
		constexpr int R = GAS_MAX_CLASSES / 32;
		unsigned long long key[R], auxw[R];
		int cnt[2][R], idle[R];
		const int n = *(volatile const int32_t *)blk;
#pragma unroll
		for (int r = 0; r < R; r++) {
			key[r] = plan.cls_key[r * 32 + lane];
			auxw[r] = plan.cls_aux[r * 32 + lane];
			cnt[0][r] = plan.cls_count[r * 32 + lane];
			cnt[1][r] = plan.cls_count[GAS_MAX_CLASSES + r * 32 + lane];
			idle[r] = blockIdx.x == 0 ? plan.cls_idle[r * 32 + lane] : 0;
		}
		if (lane == 0) { // while the loads fly
			for (int s = 0; s < cf.stages; s++) {
				mbar_init(&s_full[s], 1);
				mbar_init(&s_empty[s], kConsumerWarps);
			}
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
		const int par = (n + 1) & 1; // the prologue already advanced the counter
		int base = 0;
#pragma unroll
		for (int r = 0; r < R; r++) {
			const int count = par ? cnt[1][r] : cnt[0][r];
			const bool on = key[r] != 0ULL && (int)(key[r] & 3u) == PATH_STREAM && count > 0;
			const unsigned m = __ballot_sync(0xffffffffu, on);
			if (on) {
				ClassInfo ci = cls_decode(key[r], count);
				ci.slot = r * 32 + lane;
				if (ci.flags & CLS_SCALED) {
					ci.scale[0] = __uint_as_float((unsigned)(auxw[r] & 0xffffffffu));
					ci.scale[1] = __uint_as_float((unsigned)(auxw[r] >> 32));
				}
				s_cls[base + __popc(m & ((1u << lane) - 1u))] = ci;
			}
			base += __popc(m);
			// Slot recycling (one CTA, once per block, between two prologues): a slot whose class stayed empty for
			// GAS_CLS_IDLE_BLOCKS blocks is handed back.  Nothing reads a slot with a zero count, so clearing it here
			// cannot race with the other CTAs of this launch or with the voice-parallel kernel.
			if (blockIdx.x == 0 && r * 32 + lane < GAS_CLS_DYNAMIC && key[r] != 0ULL) {
				const int age = count > 0 ? 0 : idle[r] + 1;
				if (age >= GAS_CLS_IDLE_BLOCKS) {
					plan.cls_aux[r * 32 + lane] = CLS_AUX_NONE;
					plan.cls_key[r * 32 + lane] = 0ULL;
					plan.cls_idle[r * 32 + lane] = 0;
				} else if (age != idle[r]) {
					plan.cls_idle[r * 32 + lane] = age;
				}
			}
		}
		__syncwarp();
		if (tlp) {
			tlp[11] = gtime();
		}
		unit_iter_init_warp(it, s_cls, base, cf, C, blockIdx.x, gridDim.x, lane);
		if (lane == 0) {
			s_it = it;
		}
		__syncwarp();
		asm volatile("bar.arrive 2, %0;" ::"n"(kThreads) : "memory");
		if (tlp) {
			tlp[1] = gtime();
			tlp[5] = (unsigned long long)it.remaining;
		}
	} else {
		asm volatile("bar.sync 2, %0;" ::"n"(kThreads) : "memory");
		it = s_it;
		GAS_GRID_DEP_LAUNCH();
	}
	if (it.remaining <= 0) {
		return;
	}
	unsigned char *ring = smem;

	if (warp == kConsumerWarps) {
		// ===== producer =====
		int stage = 0;
		uint32_t phase = 0;
		// Source-row indices are fetched one warp-wide load (32 list positions = 32/vb units) at a time,
		// two blocks ahead of the copies that need them, so the gather indirection never stalls the ring.
		UnitIter pf = it;
		IdxBlock cur = idx_block_load(pf, s_cls, cf, plan.list, maxv, lane, 0);
		IdxBlock nxt = idx_block_load(pf, s_cls, cf, plan.list, maxv, lane, cur.n_units);
		int seq = 0;
		if (tlp) {
			tlp[12] = gtime() + (cur.val.y == -12345 ? 1ULL : 0ULL); // first indices have arrived
		}
		while (it.remaining > 0) {
			if (seq >= cur.first_seq + cur.n_units) {
				cur = nxt;
				nxt = idx_block_load(pf, s_cls, cf, plan.list, maxv, lane, cur.first_seq + cur.n_units);
			}
			const ClassInfo &ci = s_cls[it.cid];
			const int v0 = it.batch * cf.vb;
			const int nv = min(cf.vb, ci.count - v0);
			const int tile_w = min(cf.tile_frames, cf.frames - it.tile * cf.tile_frames); // frames in this tile
			const uint32_t row_bytes = (uint32_t)tile_w * 8u;
			const int nf = ci.n_rows * C * 2; // floats of weights per voice
			const uint32_t w_bytes = ((uint32_t)(nv * nf * 4) + 15u) & ~15u;
			unsigned char *sx = ring + (size_t)stage * cf.stage_bytes;
			unsigned char *sw = sx + cf.x_bytes;
			mbar_wait(&s_empty[stage], phase ^ 1u);
			const bool no_copy = (cf.debug & 4) != 0; // experiment: arm the stage without copying anything into it
			if (lane == 0) {
				mbar_arrive_expect_tx(&s_full[stage], no_copy ? 0u : row_bytes * (uint32_t)nv + w_bytes);
				if (!no_copy) {
					bulk_g2s(sw, plan.k2_rows + (size_t)ci.slot * maxv * GAS_K2_ROW_FLOATS + (size_t)v0 * nf, w_bytes, &s_full[stage]);
				}
			}
			__syncwarp();
			{
				const int v = lane - (seq - cur.first_seq) * cf.vb; // this lane's voice inside the stage
				if (v >= 0 && v < nv && !no_copy) {
					bulk_g2s(sx + (size_t)v * row_bytes, src + (size_t)cur.val.y * cf.src_stride + (size_t)it.tile * cf.tile_frames, row_bytes,
							&s_full[stage]);
				}
			}
			if (tlp && seq < 2) {
				tlp[13 + seq] = gtime(); // copies of the first / second stage issued
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
		cc.C = C;
		cc.lane = lane;
		cc.slot = tid % cf.slots;
		cc.group = tid / cf.slots;
		cc.worker = cc.group < cf.groups;
		cc.stage = 0;
		cc.phase = 0;
		while (it.remaining > 0) {
			// dispatch on (rows, channel pairs): rows in 2..6, pairs in 1..4
			const int code = s_cls[it.cid].n_rows * 8 + C;
#define GAS_RUN(R_, C_) \
	case (R_) * 8 + (C_): consumer_run<R_, C_>(it, cc, cf, bus); break;
			switch (code) {
				GAS_RUN(2, 1) GAS_RUN(2, 2) GAS_RUN(2, 3) GAS_RUN(2, 4)
				GAS_RUN(3, 1) GAS_RUN(3, 2) GAS_RUN(3, 3) GAS_RUN(3, 4)
				GAS_RUN(4, 1) GAS_RUN(4, 2) GAS_RUN(4, 3) GAS_RUN(4, 4)
				GAS_RUN(5, 1) GAS_RUN(5, 2) GAS_RUN(5, 3) GAS_RUN(5, 4)
				GAS_RUN(6, 1) GAS_RUN(6, 2) GAS_RUN(6, 3) GAS_RUN(6, 4)
				default: // cannot happen; drain so the producer never stalls
					consumer_run<2, 1>(it, cc, cf, bus);
					break;
			}
#undef GAS_RUN
		}
		if (tl) {
			tl[4] = gtime();
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

cudaError_t launch_mix_stream(gas_ctx *ctx, const gas_frame *d_src, int src_stride, int frames, gas_frame *d_bus, cudaStream_t st) {
	static const int kSmemLimit = 216 * 1024;
	StreamCfg cf = make_cfg(frames, src_stride, kSmemLimit, ctx->num_sms);
	if (cf.stages < 2) {
		return cudaErrorInvalidConfiguration;
	}
	const size_t smem = (size_t)cf.stages * cf.stage_bytes;
	if (!ctx->k2_smem_attr_set) {
		cudaError_t e = cudaFuncSetAttribute(k_mix_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
		if (e != cudaSuccess) {
			return e;
		}
		ctx->k2_smem_attr_set = true;
	}
	// Partial sums: atomic adds into replica (CTA % replicas) of the bus layout; the voice-parallel kernel's
	// launch folds the replicas into d_bus.
	float *target = ctx->replicas > 1 ? (float *)ctx->d_rep : (float *)d_bus;
	const int rep_stride = ctx->replicas > 1 ? gas_bus_f4(ctx, frames) * 4 : 0; // floats
	if (cf.debug & 8) {
		if (!ctx->d_timeline) {
			cudaMalloc((void **)&ctx->d_timeline, 256 * 16 * sizeof(unsigned long long));
		}
		cudaMemsetAsync(ctx->d_timeline, 0, 256 * 16 * sizeof(unsigned long long), st);
		cf.timeline = ctx->d_timeline;
	}
	cudaError_t e = gas_launch_ev(k_mix_stream, dim3(ctx->num_sms), dim3(kThreads), smem, st, (ctx->pdl & 2) != 0,
			ctx->gain_after_stream ? ctx->ev_stream_started : (cudaEvent_t) nullptr, ctx->plan, ctx->g, cf, d_src, target, rep_stride, ctx->replicas,
			(const int32_t *)ctx->t.blk);
	ctx->stream_started_pending = ctx->gain_after_stream;
	ctx->launches++;
	return e;
}
