// Sum all-reduce of the per-level statistics over NVLink peer memory - the one exchange step of the sharded path
// (K*(dim+2) 64-bit integers per split level, at most 459 KB at config 3).
//
// The payload is far below the size where link bandwidth matters, so the cost of a collective here is launch and
// synchronisation latency.  Instead of a library ring/tree this is a direct all-to-all READ through peer-mapped
// memory (NVSwitch gives every GPU full bandwidth to every peer):
//   publish   every rank copies its words into its own exchange buffer and then writes the epoch number into its
//             slot of EVERY rank's flag array (system-scope release);
//   reduce    every rank waits until all slots of its own (local) flag array carry the epoch, then reads the
//             words of all ranks, adds them in rank order and writes the sums over its input.
// Two small kernels on the caller's stream, no host synchronisation, no third-party library.  Buffers alternate
// between two halves by epoch parity: a rank can only publish epoch e + 2 after it has reduced e + 1, which needed
// every peer's publish of e + 1, which those peers issued after they finished reading e - so nobody is still
// reading the half that is overwritten.  Integer sums: the result does not depend on the order, every rank gets
// identical bits.  Waits are bounded (a lost peer traps with a message instead of hanging the GPU).
//
// The exchange blocks are reached either directly (one process driving several devices with peer access enabled,
// qb200_create_multi) or through CUDA IPC handles (one process per GPU, qb200_comm_export / qb200_comm_attach).
#include "qb200_launch.hpp"

#include <cstdio>

namespace qb {

namespace {

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// src -> my exchange buffer; the last block to finish announces the epoch to every rank.
__global__ void __launch_bounds__(256)
    comm_publish_kernel(const unsigned long long *__restrict__ src, const size_t count, unsigned long long *__restrict__ my_buf,
                        const CommPeers peers, const int rank, const int world, const unsigned long long epoch,
                        unsigned int *__restrict__ ticket) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) my_buf[i] = src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    if (t == gridDim.x - 1) {
      *ticket = 0;  // ready for the next call (stream order: nothing else touches it before this kernel ends)
      __threadfence_system();
      for (int p = 0; p < world; p++) st_release_sys(peers.flags[p] + rank, epoch);
    }
  }
}

// wait for every rank's announcement of `epoch`, then dst[i] = sum over ranks of their buffer[i]
__global__ void __launch_bounds__(256)
    comm_reduce_kernel(unsigned long long *__restrict__ dst, const size_t count, const CommPeers peers, const size_t buf_offset_words,
                       const int rank, const int world, const unsigned long long epoch) {
  if (threadIdx.x < world) {
    const unsigned long long *flag = peers.flags[rank] + threadIdx.x;
    unsigned int spins = 0;
    while (ld_acquire_sys(flag) < epoch) {
      __nanosleep(64);
      if (++spins > (1u << 25)) {  // ~ 2+ seconds: a peer never arrived
        printf("libqb200: all-reduce wait timed out (rank %d waiting for rank %d, epoch %llu)\n", rank, (int)threadIdx.x, epoch);
        __trap();
      }
    }
  }
  __syncthreads();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    unsigned long long acc = 0;
    for (int p = 0; p < world; p++) acc += __ldcv(peers.bufs[p] + buf_offset_words + i);  // not cached: peers rewrite it
    dst[i] = acc;
  }
}

}  // namespace

// One all-reduce round of `count` <= capacity words (the caller splits larger payloads).  `epoch` must be the same
// on every rank and increase by one per call; the buffer half is epoch & 1.
cudaError_t launch_comm_allreduce(unsigned long long *dev_words, size_t count, const CommPeers &peers, size_t cap_words, int rank,
                                  int world, unsigned long long epoch, unsigned int *ticket, int sm_count, cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  if (count > cap_words || world > kCommMaxWorld) return cudaErrorInvalidValue;
  const size_t half = (epoch & 1ull) * cap_words;
  unsigned int blocks = (unsigned int)((count + 255) / 256);
  const unsigned int cap = (unsigned int)sm_count * 2;
  if (blocks > cap) blocks = cap;
  comm_publish_kernel<<<blocks, 256, 0, stream>>>(dev_words, count, peers.bufs[rank] + half, peers, rank, world, epoch, ticket);
  count_launch();
  comm_reduce_kernel<<<blocks, 256, 0, stream>>>(dev_words, count, peers, half, rank, world, epoch);
  count_launch();
  return cudaGetLastError();
}

}  // namespace qb
