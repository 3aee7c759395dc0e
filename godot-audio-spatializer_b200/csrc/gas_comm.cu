// gas_comm.cu — multi-GPU sum of the per-rank partial bus buffers over peer memory (NVLink / NVSwitch).
//
// Voices are sharded over the ranks of one box; after a mix block every rank holds a partial [bus][pair][frame]
// buffer.  Each rank owns an exchange allocation {two parity buffers, an arrival counter} that every other rank
// has mapped through CUDA IPC.  Per block n:
//   k_comm_push   (after the mix kernels) zeroes the own buffer of the NEXT block, adds the local partial sums into
//                 buffer n&1 of EVERY rank (red.global.add.v4.f32 on peer pointers: the transfer is the reduction,
//                 no staging copy, no separate collective launch), fences at system scope and bumps every rank's
//                 arrival counter once;
//   k_comm_finish spins (one thread) until the own counter shows that all ranks have pushed block n, then copies
//                 the complete sum over the partial one.
// The two halves can be issued apart (gas_reduce_bus_begin_device / _end_device), and k_comm_exchange does
// finish(n - 1) + push(n) in one launch on the context's exchange stream: a caller that keeps one block in flight
// runs the whole exchange beside the mix of the next block, off the critical path.
// One arrival round per block is all the synchronisation there is: a rank can only push block n+1 after it has seen
// every rank's push of block n, and every rank zeroed its buffer for n+1 before that push.  Both kernels are plain
// stream work and can be captured into the step graph (targets derive from a device-side sequence number).
#include "gas_internal.h"

namespace {

struct CommArgs {
	float4 *xbuf[8];                 // exchange buffers of all ranks (own one included), 2 parities each
	unsigned long long *arrived[8];  // arrival counters of all ranks
	int n_ranks, rank;
	int bus_f4;                      // 16-byte elements of one bus buffer
	int xstride_f4;                  // elements between the two parity buffers
};

__global__ void __launch_bounds__(256) k_comm_push(CommArgs a, const float4 *__restrict__ partial, unsigned long long *seq, int *ticket) {
	const unsigned long long n = *(volatile unsigned long long *)seq; // blocks pushed so far (seq[0])
	const int par = (int)(n & 1ULL);
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < a.bus_f4) {
		a.xbuf[a.rank][(size_t)(par ^ 1) * a.xstride_f4 + i] = make_float4(0.f, 0.f, 0.f, 0.f); // next block's buffer
		const float4 v = partial[i];
		for (int r = 0; r < a.n_ranks; r++) {
			float *dst = reinterpret_cast<float *>(a.xbuf[r] + (size_t)par * a.xstride_f4 + i);
			gas_red_add_v4(dst, v.x, v.y, v.z, v.w);
		}
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence_system(); // cumulative: orders the whole CTA's peer adds (they happened before the barrier) before the ticket
		const int t = atomicAdd(ticket, 1);
		if (t == (int)gridDim.x - 1) { // every CTA's adds are fenced: tell all ranks this rank is in (fire and forget)
			*ticket = 0;
			seq[0] = n + 1ULL;
			__threadfence_system();
			for (int r = 0; r < a.n_ranks; r++) {
				gas_red_release_sys_add_u64(a.arrived[r], 1ULL);
			}
		}
	}
}

__global__ void __launch_bounds__(256) k_comm_finish(CommArgs a, float4 *__restrict__ bus, unsigned long long *seq, int *ticket) {
	__shared__ unsigned long long s_n;
	if (threadIdx.x == 0) {
		const unsigned long long n = *(volatile unsigned long long *)(seq + 1); // blocks finished so far (seq[1])
		const unsigned long long want = (n + 1ULL) * (unsigned long long)a.n_ranks;
		const unsigned long long *flag = a.arrived[a.rank];
		unsigned long long seen;
		do {
			seen = gas_ld_acquire_sys_u64(flag);
		} while (seen < want);
		s_n = n;
	}
	__syncthreads();
	const int par = (int)(s_n & 1ULL);
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < a.bus_f4) {
		bus[i] = __ldcg(a.xbuf[a.rank] + (size_t)par * a.xstride_f4 + i);
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		const int t = atomicAdd(ticket + 1, 1);
		if (t == (int)gridDim.x - 1) {
			ticket[1] = 0;
			seq[1] = s_n + 1ULL;
		}
	}
}

// finish(n - 1) + push(n) in one launch, for callers that keep one block in flight: the complete sum of the
// previous block goes to `prev_sum`, the partial sums of the new block are pushed to every rank.
__global__ void __launch_bounds__(256) k_comm_exchange(CommArgs a, const float4 *__restrict__ partial, float4 *__restrict__ prev_sum,
		unsigned long long *seq, int *ticket) {
	__shared__ unsigned long long s_push, s_fin;
	if (threadIdx.x == 0) {
		const unsigned long long n_push = *(volatile unsigned long long *)seq;
		const unsigned long long n_fin = *(volatile unsigned long long *)(seq + 1);
		if (n_fin < n_push) { // a block is outstanding: wait until every rank has pushed it
			const unsigned long long want = (n_fin + 1ULL) * (unsigned long long)a.n_ranks;
			const unsigned long long *flag = a.arrived[a.rank];
			unsigned long long seen;
			do {
				seen = gas_ld_acquire_sys_u64(flag);
			} while (seen < want);
		}
		s_push = n_push;
		s_fin = n_fin;
	}
	__syncthreads();
	const unsigned long long n_push = s_push, n_fin = s_fin;
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < a.bus_f4) {
		if (n_fin < n_push && prev_sum) {
			prev_sum[i] = __ldcg(a.xbuf[a.rank] + (size_t)(n_fin & 1ULL) * a.xstride_f4 + i);
		}
		const int par = (int)(n_push & 1ULL);
		a.xbuf[a.rank][(size_t)(par ^ 1) * a.xstride_f4 + i] = make_float4(0.f, 0.f, 0.f, 0.f); // next block's buffer (= the one just read)
		const float4 v = partial[i];
		for (int r = 0; r < a.n_ranks; r++) {
			float *dst = reinterpret_cast<float *>(a.xbuf[r] + (size_t)par * a.xstride_f4 + i);
			gas_red_add_v4(dst, v.x, v.y, v.z, v.w);
		}
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence_system();
		const int t = atomicAdd(ticket, 1);
		if (t == (int)gridDim.x - 1) {
			*ticket = 0;
			seq[0] = n_push + 1ULL;
			if (n_fin < n_push) {
				seq[1] = n_fin + 1ULL;
			}
			__threadfence_system();
			for (int r = 0; r < a.n_ranks; r++) {
				gas_red_release_sys_add_u64(a.arrived[r], 1ULL);
			}
		}
	}
}

} // namespace

static CommArgs comm_args(gas_ctx *ctx, int frames) {
	CommArgs a{};
	a.n_ranks = ctx->comm_ranks;
	a.rank = ctx->comm_rank;
	a.bus_f4 = gas_bus_f4(ctx, frames);
	a.xstride_f4 = ctx->comm_stride_f4;
	for (int r = 0; r < a.n_ranks; r++) {
		a.xbuf[r] = reinterpret_cast<float4 *>(ctx->peer_exchange[r]);
		a.arrived[r] = reinterpret_cast<unsigned long long *>(reinterpret_cast<float4 *>(ctx->peer_exchange[r]) + 2 * (size_t)ctx->comm_stride_f4);
	}
	return a;
}

cudaError_t launch_comm_push(gas_ctx *ctx, const gas_frame *d_bus, int frames, cudaStream_t st) {
	const CommArgs a = comm_args(ctx, frames);
	k_comm_push<<<(a.bus_f4 + 255) / 256, 256, 0, st>>>(a, (const float4 *)d_bus, ctx->d_comm_seq, ctx->d_comm_ticket);
	ctx->launches++;
	return cudaGetLastError();
}

cudaError_t launch_comm_finish(gas_ctx *ctx, gas_frame *d_bus, int frames, cudaStream_t st) {
	const CommArgs a = comm_args(ctx, frames);
	k_comm_finish<<<(a.bus_f4 + 255) / 256, 256, 0, st>>>(a, (float4 *)d_bus, ctx->d_comm_seq, ctx->d_comm_ticket);
	ctx->launches++;
	return cudaGetLastError();
}

cudaError_t launch_comm_exchange(gas_ctx *ctx, const gas_frame *d_partial, gas_frame *d_prev_sum, int frames, cudaStream_t st) {
	const CommArgs a = comm_args(ctx, frames);
	k_comm_exchange<<<(a.bus_f4 + 255) / 256, 256, 0, st>>>(a, (const float4 *)d_partial, (float4 *)d_prev_sum, ctx->d_comm_seq, ctx->d_comm_ticket);
	ctx->launches++;
	return cudaGetLastError();
}
