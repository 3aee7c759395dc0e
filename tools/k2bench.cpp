// k2bench — native micro-harness around the C ABI: mix-side kernels alone (no gain kernel, no Python),
// per-kernel CUDA-event timing from gas_profile_read.  Build: make -C tools.  Usage on the GPU box:
//   tools/k2bench [voices] [frames] [reverb_fraction] [blocks] [sets]
#include "../include/gas.h"
extern "C" void *gas_debug_timeline(gas_ctx *ctx);

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x)                                                                        \
	do {                                                                             \
		int _s = (x);                                                                \
		if (_s != 0) {                                                               \
			fprintf(stderr, "%s failed: %d %s\n", #x, _s, gas_last_error(ctx));      \
			return 1;                                                                \
		}                                                                            \
	} while (0)

static uint64_t sm64(uint64_t &s) {
	uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}
static float u01(uint64_t &s) { return (float)((sm64(s) >> 40) * (1.0 / 16777216.0)); }

int main(int argc, char **argv) {
	const int V = argc > 1 ? atoi(argv[1]) : 16384;
	const int F = argc > 2 ? atoi(argv[2]) : 512;
	const float rev = argc > 3 ? (float)atof(argv[3]) : 0.25f;
	const int blocks = argc > 4 ? atoi(argv[4]) : 64;
	const int sets = argc > 5 ? atoi(argv[5]) : 8;
	gas_ctx *ctx = nullptr;
	gas_config cfg;
	gas_config_defaults(&cfg);
	cfg.max_instances = V;
	cfg.max_voices = V;
	cfg.max_frames = F;
	cfg.num_buses = 2;
	cfg.speaker_mode = GAS_SPEAKER_SURROUND_71;
	cfg.mix_rate = 48000.f;
	if (gas_create(&cfg, &ctx) != 0) {
		fprintf(stderr, "gas_create: %s\n", gas_last_error(nullptr));
		return 1;
	}
	gas_spatializer sp;
	gas_spatializer_defaults(&sp);
	sp.mix_channel_mode = 1;
	CK(gas_spatializer_set(ctx, 0, &sp));
	std::vector<int32_t> ids(V), zeros(V, 0);
	for (int i = 0; i < V; i++) {
		ids[i] = i;
	}
	CK(gas_instance_init(ctx, V, ids.data(), zeros.data()));
	uint64_t seed = 12345;
	std::vector<gas_params> pa(V), pb(V);
	for (int i = 0; i < V; i++) {
		for (int which = 0; which < 2; which++) {
			gas_params &p = which ? pb[i] : pa[i];
			memset(&p, 0, sizeof(p));
			p.pitch_scale = 1.f;
			p.linear_attenuation = 0.f; // filter off
			p.attenuation_filter_cutoff_hz = 5000.f;
			p.update_parameters = 1;
			for (int c = 0; c < 4; c++) {
				p.mix_volumes[c][0] = 0.05f + 0.5f * u01(seed);
				p.mix_volumes[c][1] = 0.05f + 0.5f * u01(seed);
			}
			const bool r = (i % 1000) < (int)(rev * 1000.f);
			p.n_bus = r ? 2 : 1;
			p.bus[0] = 0;
			p.bus[1] = 1;
			for (int c = 0; c < 4; c++) {
				for (int x = 0; x < 2; x++) {
					p.bus_volumes[0][c][x] = p.mix_volumes[c][x];
					p.bus_volumes[1][c][x] = p.mix_volumes[c][x] * (0.3f + 0.4f * u01(seed));
				}
			}
		}
	}
	CK(gas_params_set(ctx, V, ids.data(), pa.data()));
	CK(gas_instance_start(ctx, V, ids.data()));
	CK(gas_voice_init(ctx, V, ids.data()));
	std::vector<gas_voice> voices(V);
	for (int i = 0; i < V; i++) {
		voices[i].voice = i;
		voices[i].instance = i;
		voices[i].src_row = i;
		voices[i].flags = 0;
	}
	gas_voice *d_voices = nullptr;
	cudaMalloc(&d_voices, V * sizeof(gas_voice));
	cudaMemcpy(d_voices, voices.data(), V * sizeof(gas_voice), cudaMemcpyHostToDevice);
	std::vector<gas_frame *> d_src(sets);
	{
		std::vector<gas_frame> h((size_t)V * F);
		for (size_t k = 0; k < h.size(); k++) {
			h[k].l = u01(seed) - 0.5f;
			h[k].r = u01(seed) - 0.5f;
		}
		for (int s = 0; s < sets; s++) {
			cudaMalloc(&d_src[s], h.size() * sizeof(gas_frame));
			cudaMemcpy(d_src[s], h.data(), h.size() * sizeof(gas_frame), cudaMemcpyHostToDevice);
		}
	}
	gas_frame *d_bus = nullptr;
	cudaMalloc(&d_bus, (size_t)2 * 4 * F * sizeof(gas_frame));
	// warm-up (first block fades the buses in), then timed blocks with parameters flipping every block
	for (int b = 0; b < 4; b++) {
		CK(gas_params_set(ctx, V, ids.data(), (b & 1) ? pa.data() : pb.data()));
		CK(gas_mix_block_device(ctx, V, d_voices, d_src[b % sets], V, F, F, d_bus, nullptr));
	}
	CK(gas_sync(ctx));
	CK(gas_profile_enable(ctx, 1));
	for (int b = 0; b < blocks; b++) {
		CK(gas_params_set(ctx, V, ids.data(), (b & 1) ? pa.data() : pb.data()));
		CK(gas_sync(ctx));
		CK(gas_mix_block_device(ctx, V, d_voices, d_src[b % sets], V, F, F, d_bus, nullptr));
	}
	double ms[GAS_KERNEL_KINDS];
	uint64_t n[GAS_KERNEL_KINDS];
	CK(gas_profile_read(ctx, ms, n));
	CK(gas_profile_enable(ctx, 0));
	// graph mode: R blocks per graph (parameters fixed => linear ramps after the first block), replayed
	double graph_us = 0.0;
	{
		const int R = 8, reps = 20;
		CK(gas_sync(ctx));
		CK(gas_capture_begin(ctx));
		for (int r = 0; r < R; r++) {
			CK(gas_mix_block_device(ctx, V, d_voices, d_src[r % sets], V, F, F, d_bus, nullptr));
		}
		int32_t graph = -1;
		CK(gas_capture_end(ctx, &graph));
		for (int k = 0; k < 3; k++) {
			CK(gas_graph_launch(ctx, graph));
		}
		CK(gas_sync(ctx));
		cudaEvent_t e0, e1;
		cudaEventCreate(&e0);
		cudaEventCreate(&e1);
		cudaStream_t st = (cudaStream_t)gas_mix_stream(ctx);
		cudaEventRecord(e0, st);
		for (int k = 0; k < reps; k++) {
			CK(gas_graph_launch(ctx, graph));
		}
		cudaEventRecord(e1, st);
		CK(gas_sync(ctx));
		float t = 0.f;
		cudaEventElapsedTime(&t, e0, e1);
		graph_us = 1e3 * t / (R * reps);
	}
	if (getenv("GAS_K2_DEBUG") && (atoi(getenv("GAS_K2_DEBUG")) & 8)) { // one eager block with the timeline on
		CK(gas_mix_block_device(ctx, V, d_voices, d_src[0], V, F, F, d_bus, nullptr));
		CK(gas_sync(ctx));
		unsigned long long h[148 * 32];
		cudaMemcpy(h, gas_debug_timeline(ctx), sizeof(h), cudaMemcpyDeviceToHost);
		unsigned long long t0 = ~0ULL;
		for (int c = 0; c < 148; c++) {
			if (h[c * 32] && h[c * 32] < t0) t0 = h[c * 32];
		}
		const char *names[16] = { "start", "table+partition", "first data", "last data", "flushed", "", "", "before table loads", "table loaded", "flush begins", "partition again (dbg 4)", "table in smem", "first indices", "stage 0 issued", "stage 1 issued", "flush pass 1 (dbg 4)" };
		for (int k = 0; k < 16; k++) {
			if (k == 5 || k == 6) continue;
			double mn = 1e30, mx = 0, av = 0;
			int n = 0;
			for (int c = 0; c < 148; c++) {
				if (!h[c * 32 + k]) continue;
				const double d = (double)(h[c * 32 + k] - t0) * 1e-3;
				mn = d < mn ? d : mn;
				mx = d > mx ? d : mx;
				av += d;
				n++;
			}
			printf("  timeline %-18s min %7.2f avg %7.2f max %7.2f us (n=%d)\n", names[k], mn, n ? av / n : 0, mx, n);
		}
		int umin = 1 << 30, umax = 0;
		for (int c = 0; c < 148; c++) {
			umin = (int)h[c * 32 + 5] < umin ? (int)h[c * 32 + 5] : umin;
			umax = (int)h[c * 32 + 5] > umax ? (int)h[c * 32 + 5] : umax;
		}
		printf("  units per CTA: %d..%d\n", umin, umax);
		if (getenv("GAS_K2_DUMP")) { // per CTA: SM id, units, first data, last data, flushed (us after the first CTA started)
			for (int c = 0; c < 148; c++) {
				printf("  cta %3d sm %3d units %2d  %6.2f %6.2f %6.2f %6.2f\n", c, (int)h[c * 32 + 6], (int)h[c * 32 + 5], (double)(h[c * 32 + 1] - t0) * 1e-3,
						(double)(h[c * 32 + 2] - t0) * 1e-3, (double)(h[c * 32 + 3] - t0) * 1e-3, (double)(h[c * 32 + 4] - t0) * 1e-3);
			}
		}
	}
	std::vector<gas_frame> hb((size_t)2 * 4 * F);
	cudaMemcpy(hb.data(), d_bus, hb.size() * sizeof(gas_frame), cudaMemcpyDeviceToHost);
	double cs = 0;
	for (auto &f : hb) {
		cs += (double)f.l + (double)f.r;
	}
	const double bytes = 8.0 * V * F;
	printf("V=%d F=%d rev=%.2f  eager-events: prologue %.2f K2 %.2f K3 %.2f us | graph: %.2f us/block (%.0f GB/s)  checksum %.6f\n", V, F, rev,
			1e3 * ms[0] / n[0], 1e3 * ms[1] / n[1], 1e3 * ms[2] / n[2], graph_us, bytes / (graph_us * 1e-6) / 1e9, cs);
	gas_destroy(ctx);
	return 0;
}
