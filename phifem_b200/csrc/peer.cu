// The one exchange step of the sharded path, over NVLink peer memory instead of a collective library.
//
// `_tag_facets` changes the meaning of mesh-boundary facets when the mesh has NO exterior cell at all (reference
// src/phifem/mesh_scripts.py:469-474), so a sharded classification needs one global number per step: the sum over ranks
// of the exterior-cell counts.  An NCCL all-reduce of 8 bytes costs a kernel of its own that cannot start while the
// persistent interior-facet kernel holds every SM -- 50-70 us per step that nothing hides, a quarter of the step once
// config E is strong-scaled over 8 GPUs.  Here every rank owns an array of slots in its own HBM, mapped into every peer
// with CUDA IPC; after its cell kernel a rank STORES (epoch, count) into its slot on every peer (one 8-byte store per
// peer through NVLink), and before its mesh-boundary-facet kernel it reads the slots in its own memory -- by then long
// written, the interior-facet kernel ran in between.  Two one-warp kernels, no SM contention, no host involvement,
// capturable in a CUDA graph (the epoch lives in device memory).
//
// Slots are double-buffered by epoch parity: a rank cannot get two steps ahead of a peer (its next collect needs that
// peer's next publish, which the peer issues after its own collect), so a slot is never overwritten before it was read.
// A collect that does not see its peers within ~1 s sets the error flag instead of hanging the device.
#include <string.h>

#include "common.cuh"

struct phifem_peer_flags {
  int world, rank;
  unsigned long long* local;      // [2][world] slots in this rank's memory: (epoch << 32) | value
  unsigned long long* peers[64];  // the same array of every rank (peers[rank] == local)
  unsigned long long** table;     // device copy of `peers` (what the publish kernel reads)
  unsigned int* epoch;            // device: number of publishes so far
  int* error;                     // device: set when a collect timed out
  bool opened[64];
};

namespace phifem {
namespace {

__global__ void k_peer_publish(unsigned long long* const* __restrict__ peers, int world, int rank,
                               unsigned int* __restrict__ epoch, const int64_t* __restrict__ value) {
  __shared__ unsigned int e_s;
  if (threadIdx.x == 0) {
    e_s = *epoch + 1u;
    *epoch = e_s;
  }
  __syncthreads();
  const int q = threadIdx.x;
  if (q >= world) return;
  const int64_t v = *value;
  const unsigned long long low = v < 0 ? 0ull : (v > 0xffffffffll ? 0xffffffffull : (unsigned long long)v);
  volatile unsigned long long* slot = peers[q] + (size_t)(e_s & 1u) * world + rank;
  *slot = ((unsigned long long)e_s << 32) | low;
  __threadfence_system();
}

__global__ void k_peer_collect(const unsigned long long* __restrict__ local, int world,
                               const unsigned int* __restrict__ epoch, int64_t* __restrict__ out, int* __restrict__ error) {
  __shared__ unsigned long long sum_s;
  if (threadIdx.x == 0) sum_s = 0ull;
  __syncthreads();
  const int q = threadIdx.x;
  if (q < world) {
    const unsigned int e = *epoch;
    const volatile unsigned long long* slot = local + (size_t)(e & 1u) * world + q;
    unsigned long long w = *slot;
    int spins = 0;
    while ((unsigned int)(w >> 32) != e) {
      if (++spins > (1 << 21)) {  // ~1 s: a peer never published this epoch
        *error = 1;
        break;
      }
      __nanosleep(400);
      w = *slot;
    }
    atomicAdd(&sum_s, w & 0xffffffffull);
  }
  __syncthreads();
  if (threadIdx.x == 0) *out = (int64_t)sum_s;
}

}  // namespace
}  // namespace phifem

using namespace phifem;

extern "C" int phifem_peer_flags_create(int32_t world, int32_t rank, phifem_peer_flags** out, void* handle64) {
  PHIFEM_CHECK_ARG(out && handle64, "null pointer");
  PHIFEM_CHECK_ARG(world >= 1 && world <= 64 && rank >= 0 && rank < world, "world / rank");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  phifem_peer_flags* p = new phifem_peer_flags();
  p->world = world;
  p->rank = rank;
  p->table = nullptr;
  for (int q = 0; q < 64; ++q) {
    p->peers[q] = nullptr;
    p->opened[q] = false;
  }
  bool ok = cudaMalloc(&p->local, sizeof(unsigned long long) * 2 * world + 64) == cudaSuccess &&
            cudaMalloc(&p->epoch, sizeof(unsigned int)) == cudaSuccess && cudaMalloc(&p->error, sizeof(int)) == cudaSuccess;
  if (ok) {
    cudaMemset(p->local, 0, sizeof(unsigned long long) * 2 * world + 64);
    cudaMemset(p->epoch, 0, sizeof(unsigned int));
    cudaMemset(p->error, 0, sizeof(int));
    p->peers[rank] = p->local;
    cudaIpcMemHandle_t h;
    ok = cudaIpcGetMemHandle(&h, p->local) == cudaSuccess;
    if (ok) memcpy(handle64, &h, 64);
  }
  if (!ok || cudaDeviceSynchronize() != cudaSuccess) {
    set_error("phifem_peer_flags_create: %s", cudaGetErrorString(cudaGetLastError()));
    delete p;
    return PHIFEM_ERR_CUDA;
  }
  *out = p;
  return PHIFEM_OK;
}

extern "C" int phifem_peer_flags_connect(phifem_peer_flags* p, const void* handles) {
  PHIFEM_CHECK_ARG(p && handles, "null pointer");
  for (int q = 0; q < p->world; ++q) {
    if (q == p->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + 64 * q, 64);
    void* ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      set_error("phifem_peer_flags_connect: rank %d cannot map the slots of rank %d: %s", p->rank, q,
                cudaGetErrorString(cudaGetLastError()));
      return PHIFEM_ERR_CUDA;
    }
    p->peers[q] = (unsigned long long*)ptr;
    p->opened[q] = true;
  }
  if ((!p->table && cudaMalloc(&p->table, sizeof(void*) * 64) != cudaSuccess) ||
      cudaMemcpy(p->table, p->peers, sizeof(void*) * 64, cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("phifem_peer_flags_connect: %s", cudaGetErrorString(cudaGetLastError()));
    return PHIFEM_ERR_CUDA;
  }
  return PHIFEM_OK;
}

extern "C" int phifem_peer_flags_publish(phifem_peer_flags* p, const int64_t* value, void* stream) {
  PHIFEM_CHECK_ARG(p && value, "null pointer");
  PHIFEM_CHECK_ARG(p->table != nullptr, "connect first");
  k_peer_publish<<<1, 64, 0, (cudaStream_t)stream>>>(p->table, p->world, p->rank, p->epoch, value);
  PHIFEM_CHECK_LAUNCH();
  return PHIFEM_OK;
}

extern "C" int phifem_peer_flags_collect(phifem_peer_flags* p, int64_t* value_out, void* stream) {
  PHIFEM_CHECK_ARG(p && value_out, "null pointer");
  k_peer_collect<<<1, 64, 0, (cudaStream_t)stream>>>(p->local, p->world, p->epoch, value_out, p->error);
  PHIFEM_CHECK_LAUNCH();
  return PHIFEM_OK;
}

extern "C" int phifem_peer_flags_error(phifem_peer_flags* p) {  // synchronises; 1 = a collect timed out
  if (!p) return 0;
  int e = 0;
  cudaMemcpy(&e, p->error, sizeof(int), cudaMemcpyDeviceToHost);
  return e;
}

extern "C" void phifem_peer_flags_destroy(phifem_peer_flags* p) {
  if (!p) return;
  cudaDeviceSynchronize();
  if (p->table) cudaFree(p->table);
  for (int q = 0; q < p->world; ++q)
    if (p->opened[q]) cudaIpcCloseMemHandle(p->peers[q]);
  cudaFree(p->local);
  cudaFree(p->epoch);
  cudaFree(p->error);
  delete p;
}
