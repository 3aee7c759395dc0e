// gas_bus.cu — the bus graph after the mix: what upstream AudioServer::_mix_step does with the bus channel
// buffers once every playback has been mixed into them (reference README.md:98-100, steps 3-5; the demo layout
// examples/godot-gd-spatializer/default_bus_layout.tres:9-17 sends its Reverb bus to Master): buses are visited from the last
// to the first; each applies its volume (0 when muted, or when another bus is soloed and this one is not on a soloed send chain)
// and adds its buffers to its send bus, down to Master, whose buffers are what the audio driver gets.  Bus effects
// (AudioEffectReverb ...) are not part of this library: a bus that carries effects is processed by the caller between two calls.
//
// Element-wise: one thread owns frame pair (c, i) of every bus and walks the buses in upstream's order, so the float operations
// (buf *= volume, then target += buf) happen in the same sequence per sample.  Compiled with -fmad=false.  Restated from Godot
// 4.x as recalled: the engine is not in the reference tree.
#include "gas_internal.h"

namespace {

struct BusGraphArgs {
	int n_buses, channels, frames;
	float volume[GAS_MAX_BUSES]; // linear, 0 for silenced buses
	int send[GAS_MAX_BUSES];     // target bus (< own index); unused for bus 0
};

__global__ void __launch_bounds__(256) k_bus_graph(BusGraphArgs a, float4 *__restrict__ bus) {
	const int per_bus = a.channels * a.frames / 2; // 16-byte elements (two frames) per bus
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= per_bus) {
		return;
	}
	for (int b = a.n_buses - 1; b >= 0; b--) {
		const float vol = a.volume[b];
		float4 v = bus[(size_t)b * per_bus + i];
		v.x *= vol;
		v.y *= vol;
		v.z *= vol;
		v.w *= vol;
		bus[(size_t)b * per_bus + i] = v;
		if (b > 0) {
			float4 t = bus[(size_t)a.send[b] * per_bus + i];
			t.x += v.x;
			t.y += v.y;
			t.z += v.z;
			t.w += v.w;
			bus[(size_t)a.send[b] * per_bus + i] = t;
		}
	}
}

} // namespace

cudaError_t launch_bus_graph(gas_ctx *ctx, gas_frame *d_bus, int frames, cudaStream_t st) {
	BusGraphArgs a{};
	a.n_buses = ctx->g.num_buses;
	a.channels = ctx->g.channels;
	a.frames = frames;
	for (int b = 0; b < GAS_MAX_BUSES; b++) {
		a.volume[b] = ctx->bus_volume_lin[b];
		a.send[b] = ctx->bus_send[b];
	}
	const int per_bus = a.channels * frames / 2;
	k_bus_graph<<<(per_bus + 255) / 256, 256, 0, st>>>(a, (float4 *)d_bus);
	ctx->launches++;
	return cudaGetLastError();
}
