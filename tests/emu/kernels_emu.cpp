// kernels_emu.cpp — TEST INFRASTRUCTURE: the SOURCE of csrc/gas_resample.cu and csrc/gas_bus.cu executed on the CPU (cuda_emu.h),
// behind a small C interface for tests/test_zz_kernel_emulation.py.  The product library is not involved.
#include "cuda_emu.h"

#include "../../godot-audio-spatializer_b200/csrc/gas_resample.cu"
#define kThreads kThreadsBus // (both files are included into one translation unit)
#include "../../godot-audio-spatializer_b200/csrc/gas_bus.cu"
#undef kThreads

#include <cstring>

extern "C" {

// One block of the resampler for n_voices voices that all play `pcm` (n_frames, loop) from start[i] at pitch[i], continuing from
// pos[i] (updated).  rows: [n_voices][frames], mixed: [n_voices].
int emu_resample(const gas_frame *pcm, int n_frames, int loop, float sample_rate, float mix_rate, int n_voices, const int *start, const float *pitch,
		unsigned long long *pos, int frames, gas_frame *rows, int *mixed) {
	DevTables t{};
	GlobalCfg g{};
	g.max_voices = n_voices;
	g.max_instances = n_voices;
	g.mix_rate = mix_rate;
	std::vector<int32_t> vs_src(n_voices, 0);
	std::vector<long long> vs_start(start, start + n_voices);
	std::vector<gas_params> params(n_voices);
	std::vector<gas_voice> voices(n_voices);
	for (int i = 0; i < n_voices; i++) {
		std::memset(&params[i], 0, sizeof(gas_params));
		params[i].pitch_scale = pitch[i];
		voices[i].voice = i;
		voices[i].instance = i;
		voices[i].src_row = i;
		voices[i].flags = 0;
	}
	t.vs_src = vs_src.data();
	t.vs_start = vs_start.data();
	t.vs_pos = pos;
	t.inst_params = params.data();
	SourceDesc sd{};
	sd.pcm = pcm;
	sd.n_frames = n_frames;
	sd.loop = loop;
	sd.sample_rate = sample_rate;
	const gas_voice *vp = voices.data();
	emu::launch((unsigned)n_voices, 128, [&] { k_resample(t, g, &sd, 1, n_voices, vp, frames, rows, frames, n_voices, mixed); });
	return 0;
}

int emu_bus_graph(int n_buses, int channels, int frames, const float *volume_lin, const int *send, gas_frame *bus) {
	BusGraphArgs a{};
	a.n_buses = n_buses;
	a.channels = channels;
	a.frames = frames;
	for (int b = 0; b < n_buses; b++) {
		a.volume[b] = volume_lin[b];
		a.send[b] = send[b];
	}
	const int per_bus = channels * frames / 2;
	emu::launch((unsigned)((per_bus + 255) / 256), 256, [&] { k_bus_graph(a, reinterpret_cast<float4 *>(bus)); });
	return 0;
}
}
