// peer.cu -- halo planes over NVLink peer memory (z-slab decomposition).
//
// One process per GPU; each rank maps its neighbours' receive buffers (CUDA IPC) into its own
// address space.  Per step and neighbour:
//   k_peer_put   copies one boundary plane from local HBM straight into the neighbour's staging
//                buffer with plain vectorised stores over NVLink (no NCCL, no host), then -- after
//                a system-scope fence, by the last block through a ticket -- publishes the sender's
//                step counter in the neighbour's flag word;
//   k_peer_wait  spins (bounded) on the local flag word until the expected step has arrived and
//                copies the staged plane into the halo plane of the local field.
// Both are ordinary stream-ordered launches: a whole slab step (puts, waits, the FEM kernel) can
// be captured in a CUDA graph, which NCCL point-to-point on this stack could not.
#include "dn_common.cuh"

namespace dn {

__global__ void __launch_bounds__(256) k_peer_put(float4* __restrict__ dst, const float4* __restrict__ src,
                                                  long long n4, int* remote_flag, int* local_counter,
                                                  unsigned int* ticket) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) dst[i] = src[i];
  __threadfence_system();                      // this thread's peer stores are visible system-wide
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    if (t == gridDim.x - 1) {                  // every block has fenced its stores
      *ticket = 0u;
      const int step = *local_counter + 1;
      *local_counter = step;
      __threadfence_system();
      *reinterpret_cast<volatile int*>(remote_flag) = step;
    }
  }
}

// status: 0 = ok; on time-out the kernel writes 1 there and still copies (the caller checks it
// when it next synchronises) -- a lost neighbour must not hang the GPU.
// `staged` is written by the remote GPU while this kernel may already be resident: it is read with
// L2-coherent loads (__ldcg), never through the non-coherent path a const __restrict__ pointer allows.
__global__ void __launch_bounds__(256) k_peer_wait(float4* __restrict__ halo, const float4* staged,
                                                   long long n4, const int* flag, int* expect,
                                                   long long max_spins, int* status) {
  __shared__ int want;
  if (threadIdx.x == 0) {
    want = *expect + 1;
    long long spins = 0;
    while (*reinterpret_cast<const volatile int*>(flag) < want) {
      if (++spins > max_spins) { *status = 1; break; }
      __nanosleep(64);
    }
    __threadfence_system();
  }
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) halo[i] = __ldcg(staged + i);
  // the LAST block to finish bumps the expectation (same ticket idea, on the status word's neighbour)
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(reinterpret_cast<unsigned int*>(status) + 1, 1u);
    if (t == gridDim.x - 1) {
      reinterpret_cast<unsigned int*>(status)[1] = 0u;
      *expect = want;
    }
  }
}

// Sum of one float per rank over all ranks, through peer memory, in RANK ORDER (bit-identical on
// every rank and from run to run).  `slots`: this rank's receive area, double[world] values then
// int[world] flags (one pair of areas per parity, chosen by the caller); `peer_slots[r]`: the same
// area of rank r mapped here (entry `rank` unused).  One CTA of 32 threads; world <= 32.
__global__ void __launch_bounds__(32) k_peer_allreduce(const float* __restrict__ partial, float* __restrict__ out,
                                                       double* slots, double* const* peer_slots, int rank,
                                                       int world, int* counter, long long max_spins,
                                                       int* status) {
  const int t = threadIdx.x;
  const int step = *counter + 1;
  const double mine = (double)*partial;
  int* flags = reinterpret_cast<int*>(slots + world);
  if (t < world && t != rank) {
    double* ps = peer_slots[t];
    ps[rank] = mine;
    __threadfence_system();
    *reinterpret_cast<volatile int*>(reinterpret_cast<int*>(ps + world) + rank) = step;
    long long spins = 0;
    while (*reinterpret_cast<volatile int*>(flags + t) < step) {
      if (++spins > max_spins) { *status = 1; break; }
      __nanosleep(64);
    }
    __threadfence_system();
  }
  __syncwarp();
  if (t == 0) {
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += (r == rank) ? mine : *reinterpret_cast<volatile double*>(slots + r);
    *out = (float)s;
    *counter = step;
  }
}

// Sum of the loss slots a linked z-slab launch filled (fem3d_tma.cuh / finish_loss_w0): waits (bounded) until
// every rank's flag has reached this rank's launch counter, then adds the doubles in rank order.
__global__ void __launch_bounds__(32) k_peer_loss_sum(const double* slots, int world, const int* step,
                                                      long long max_spins, int* status, float* out) {
  const int t = threadIdx.x;
  const int want = *step;                       // launches completed on this rank with this parity
  const int* flags = reinterpret_cast<const int*>(slots + world);
  if (t < world) {
    long long spins = 0;
    while (*reinterpret_cast<const volatile int*>(flags + t) < want) {
      if (++spins > max_spins) { *status = 1; break; }
      __nanosleep(64);
    }
    __threadfence_system();
  }
  __syncwarp();
  if (t == 0) {
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += *reinterpret_cast<const volatile double*>(slots + r);
    *out = (float)s;
  }
}

cudaError_t launch_peer_loss_sum(const double* slots, int world, const int* step, long long max_spins, int* status,
                                 float* out, cudaStream_t s) {
  k_peer_loss_sum<<<1, 32, 0, s>>>(slots, world, step, max_spins, status, out);
  return cudaGetLastError();
}

cudaError_t launch_peer_allreduce(const float* partial, float* out, double* slots, double* const* peer_slots,
                                  int rank, int world, int* counter, long long max_spins, int* status,
                                  cudaStream_t s) {
  k_peer_allreduce<<<1, 32, 0, s>>>(partial, out, slots, peer_slots, rank, world, counter, max_spins, status);
  return cudaGetLastError();
}

cudaError_t launch_peer_put(float* dst_peer, const float* src, size_t n, int* remote_flag, int* local_counter,
                            unsigned int* ticket, cudaStream_t s) {
  const long long n4 = (long long)(n / 4);
  int grid = (int)((n4 + 255) / 256);
  if (grid > 64) grid = 64;
  if (grid < 1) grid = 1;
  k_peer_put<<<grid, 256, 0, s>>>(reinterpret_cast<float4*>(dst_peer), reinterpret_cast<const float4*>(src), n4,
                                  remote_flag, local_counter, ticket);
  return cudaGetLastError();
}

cudaError_t launch_peer_wait(float* halo, const float* staged, size_t n, const int* flag, int* expect,
                             long long max_spins, int* status, cudaStream_t s) {
  const long long n4 = (long long)(n / 4);
  int grid = (int)((n4 + 255) / 256);
  if (grid > 32) grid = 32;
  if (grid < 1) grid = 1;
  k_peer_wait<<<grid, 256, 0, s>>>(reinterpret_cast<float4*>(halo), reinterpret_cast<const float4*>(staged), n4,
                                   flag, expect, max_spins, status);
  return cudaGetLastError();
}

}  // namespace dn
