// gas_prologue.cu — the stand-alone planner kernel: one launch plans one mix block (see gas_plan.cuh for what a plan
// is).  Used by the one-call-per-block entry points (gas_mix_block*, the stream form, the first block of a pipelined run);
// in a pipelined run (gas_step_device) the same planner code runs on the control warps of the step kernel instead, beside
// the streaming of the previous block.
//
// Compiled with -fmad=false (coefficient preparation is double arithmetic narrowed to float).
#include "gas_plan.cuh"

namespace {

constexpr int kPlanThreads = 256; // 128 voices per CTA and pass

__global__ void __launch_bounds__(kPlanThreads) k_plan(gasplan::PlanArgs a) {
	__shared__ gasplan::PlanSmem S;
	gasplan::PlanGroup G;
	G.tid = threadIdx.x;
	G.nthreads = kPlanThreads;
	G.cta = blockIdx.x;
	G.n_cta = gridDim.x;
	G.bar_id = 1;
	G.tl = nullptr;
	GAS_GRID_DEP_WAIT(); // programmatic dependent launch: the previous block's kernels are complete after this
	gasplan::plan_block(G, S, a, 0);
}

} // namespace

cudaError_t launch_plan(gas_ctx *ctx, int n_voices, const gas_voice *d_voices, int src_rows, int frames, gas_frame *d_bus, gas_frame *d_peaks,
		cudaStream_t st) {
	gasplan::PlanArgs a{};
	a.t = ctx->t;
	a.g = ctx->g;
	a.plan = ctx->plan;
	a.inst_hwm = ctx->inst_hwm;
	a.n_voices = n_voices;
	a.voices = d_voices;
	a.src_rows = src_rows;
	a.bus = (float4 *)d_bus;
	a.bus_f4 = gas_bus_f4(ctx, frames);
	a.peaks = (float2 *)d_peaks;
	a.scaled_classes = ctx->scaled_classes ? 1 : 0;
	int work = ctx->inst_hwm > n_voices ? ctx->inst_hwm : n_voices;
	work = work > 1 ? work : 1;
	int blocks = (work + kPlanThreads / 2 - 1) / (kPlanThreads / 2);
	const int cap = ctx->num_sms * 8; // one wave: larger blocks loop
	blocks = blocks > cap ? cap : blocks;
	cudaError_t e = gas_launch(k_plan, dim3(blocks), dim3(kPlanThreads), 0, st, (ctx->pdl & 1) != 0, a);
	ctx->launches++;
	return e;
}
