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

// ---- halo exchange of CSR / load-vector contributions over NVLink peer memory ------------------------------------------
// The "halo-row contributions are exchanged ... over NVLink" step of a sharded assembly whose ranks assemble their own
// CELLS (per-entity kernels, mode "exchange" of phifem_b200/dist.py) instead of their own ROWS: contributions to rows of
// another rank accumulate in a send segment and must be added into that rank's CSR values.  One kernel does the whole
// exchange -- no pack, no collective library, no host:
//   push     every CTA copies its share of the send values straight into the owners' receive buffers (remote stores
//            through the CUDA IPC mapping: NVLink), fences, and takes a ticket; the CTA with the last ticket stores the
//            epoch into this rank's flag slot on every peer;
//   receive  every CTA waits until the flags of all peers carry the epoch (they are stores into THIS rank's memory), then
//            adds its share of the received values into the destination array through the precomputed slot list
//            (fp64 reductions: two peers may contribute to the same entry).
// Receive buffers and flags alternate with the epoch's parity, as the exterior-cell slots above do: a rank cannot be two
// exchanges ahead of a peer.  The epoch lives in device memory (a one-thread kernel advances it), so the pair of
// launches replays inside a CUDA graph.
struct phifem_halo {
  int world, rank;
  int64_t capacity;                // doubles per receive buffer (the same on every rank)
  int64_t n_recv, n_send;          // values this rank receives / sends per exchange
  char* base;                      // local allocation: [2][64] flags (8 bytes each), then [2][capacity] doubles
  char* peers[64];                 // the same allocation of every rank (peers[rank] == base)
  bool opened[64];
  char** table;                    // device copy of `peers`
  int64_t* remote_offset;          // device [world]: where this rank's values start in peer q's receive buffer
  int64_t* send_ptr;               // device [world + 1]: value range of peer q in the send list
  int64_t* send_src;               // device [world]: first element of peer q's contiguous segment (send_index == NULL)
  unsigned int* epoch;             // device
  unsigned int* ticket;            // device
  int* error;                      // device
};

namespace phifem {
namespace {
constexpr size_t kHaloFlagBytes = 2 * 64 * sizeof(unsigned long long);

__global__ void k_halo_begin(unsigned int* epoch, unsigned int* ticket) {
  *epoch += 1u;
  *ticket = 0u;
}

__global__ void __launch_bounds__(256) k_halo_exchange(char* const* __restrict__ peers, int world, int rank,
                                                       int64_t capacity, const unsigned int* __restrict__ epoch,
                                                       unsigned int* __restrict__ ticket,
                                                       const int64_t* __restrict__ remote_offset,
                                                       const int64_t* __restrict__ send_ptr,
                                                       const int64_t* __restrict__ send_src,
                                                       const double* __restrict__ src,
                                                       const int64_t* __restrict__ send_index,
                                                       const int64_t* __restrict__ recv_index, int64_t n_recv,
                                                       double* __restrict__ dst, int* __restrict__ error) {
  const unsigned int e = *epoch;
  const int par = (int)(e & 1u);
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  // push
  for (int q = 0; q < world; ++q) {
    if (q == rank) continue;
    const int64_t lo = send_ptr[q], n = send_ptr[q + 1] - lo;
    double* out = reinterpret_cast<double*>(peers[q] + kHaloFlagBytes) + (size_t)par * capacity + remote_offset[q];
    const int64_t s0 = send_src[q];
    for (int64_t j = tid; j < n; j += nth) out[j] = send_index ? src[send_index[lo + j]] : src[s0 + j];
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last) {  // every CTA of this rank has pushed and fenced: tell the peers
    __threadfence();
    if (threadIdx.x < world && (int)threadIdx.x != rank) {
      volatile unsigned long long* flag =
          reinterpret_cast<unsigned long long*>(peers[threadIdx.x]) + (size_t)par * 64 + rank;
      *flag = (unsigned long long)e;
    }
    __threadfence_system();
  }
  // receive
  if (threadIdx.x < world && (int)threadIdx.x != rank) {
    const volatile unsigned long long* flag =
        reinterpret_cast<const unsigned long long*>(peers[rank]) + (size_t)par * 64 + threadIdx.x;
    int spins = 0;
    while ((unsigned int)*flag != e) {
      if (++spins > (1 << 21)) {  // ~1 s: a peer never pushed this epoch
        *error = 1;
        break;
      }
      __nanosleep(400);
    }
  }
  __syncthreads();
  __threadfence_system();
  const double* in = reinterpret_cast<const double*>(peers[rank] + kHaloFlagBytes) + (size_t)par * capacity;
  for (int64_t i = tid; i < n_recv; i += nth) atomicAdd(dst + recv_index[i], __ldcv(in + i));  // .cv: not from a stale L1 line
}
}  // namespace
}  // namespace phifem

extern "C" int phifem_halo_create(int32_t world, int32_t rank, int64_t capacity, phifem_halo** out, void* handle64) {
  PHIFEM_CHECK_ARG(out && handle64, "null pointer");
  PHIFEM_CHECK_ARG(world >= 1 && world <= 64 && rank >= 0 && rank < world && capacity >= 0, "world / rank / capacity");
  phifem_halo* h = new phifem_halo();
  memset(h, 0, sizeof(*h));
  h->world = world;
  h->rank = rank;
  h->capacity = capacity;
  const size_t bytes = kHaloFlagBytes + 2 * (size_t)(capacity > 0 ? capacity : 1) * sizeof(double);
  bool ok = cudaMalloc(&h->base, bytes) == cudaSuccess && cudaMalloc(&h->epoch, sizeof(unsigned int)) == cudaSuccess &&
            cudaMalloc(&h->ticket, sizeof(unsigned int)) == cudaSuccess && cudaMalloc(&h->error, sizeof(int)) == cudaSuccess &&
            cudaMalloc(&h->table, sizeof(char*) * 64) == cudaSuccess &&
            cudaMalloc(&h->remote_offset, sizeof(int64_t) * 64) == cudaSuccess &&
            cudaMalloc(&h->send_ptr, sizeof(int64_t) * 65) == cudaSuccess &&
            cudaMalloc(&h->send_src, sizeof(int64_t) * 64) == cudaSuccess;
  if (ok) {
    cudaMemset(h->base, 0, bytes);
    cudaMemset(h->epoch, 0, sizeof(unsigned int));
    cudaMemset(h->ticket, 0, sizeof(unsigned int));
    cudaMemset(h->error, 0, sizeof(int));
    h->peers[rank] = h->base;
    cudaIpcMemHandle_t ipc;
    ok = cudaIpcGetMemHandle(&ipc, h->base) == cudaSuccess;
    if (ok) memcpy(handle64, &ipc, 64);
  }
  if (!ok || cudaDeviceSynchronize() != cudaSuccess) {
    set_error("phifem_halo_create: %s", cudaGetErrorString(cudaGetLastError()));
    delete h;
    return PHIFEM_ERR_CUDA;
  }
  *out = h;
  return PHIFEM_OK;
}

extern "C" int phifem_halo_connect(phifem_halo* h, const void* handles, const int64_t* remote_offset,
                                   const int64_t* send_ptr, const int64_t* send_src, int64_t n_recv) {
  PHIFEM_CHECK_ARG(h && handles && remote_offset && send_ptr && send_src, "null pointer");
  PHIFEM_CHECK_ARG(n_recv >= 0 && n_recv <= h->capacity, "n_recv exceeds the capacity agreed at creation");
  for (int q = 0; q < h->world; ++q) {
    if (q == h->rank) continue;
    PHIFEM_CHECK_ARG(remote_offset[q] >= 0 && remote_offset[q] + (send_ptr[q + 1] - send_ptr[q]) <= h->capacity,
                     "a send segment does not fit the peer's receive buffer");
    cudaIpcMemHandle_t ipc;
    memcpy(&ipc, (const char*)handles + 64 * q, 64);
    void* ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, ipc, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      set_error("phifem_halo_connect: rank %d cannot map the receive buffer of rank %d: %s", h->rank, q,
                cudaGetErrorString(cudaGetLastError()));
      return PHIFEM_ERR_CUDA;
    }
    h->peers[q] = (char*)ptr;
    h->opened[q] = true;
  }
  h->n_recv = n_recv;
  h->n_send = send_ptr[h->world];
  if (cudaMemcpy(h->table, h->peers, sizeof(char*) * 64, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(h->remote_offset, remote_offset, sizeof(int64_t) * h->world, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(h->send_ptr, send_ptr, sizeof(int64_t) * (h->world + 1), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(h->send_src, send_src, sizeof(int64_t) * h->world, cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("phifem_halo_connect: %s", cudaGetErrorString(cudaGetLastError()));
    return PHIFEM_ERR_CUDA;
  }
  return PHIFEM_OK;
}

extern "C" int phifem_halo_exchange(phifem_halo* h, const double* src, const int64_t* send_index,
                                    const int64_t* recv_index, double* dst, void* stream) {
  PHIFEM_CHECK_ARG(h && h->table, "create and connect first");
  PHIFEM_CHECK_ARG((h->n_send == 0 || src) && (h->n_recv == 0 || (recv_index && dst)), "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t work = h->n_send > h->n_recv ? h->n_send : h->n_recv;
  int64_t grid = (work + 256 * 4 - 1) / (256 * 4);
  if (grid < 1) grid = 1;
  if (grid > 2 * kNumSMs) grid = 2 * kNumSMs;  // all CTAs resident: the receive phase spins on the peers' flags
  k_halo_begin<<<1, 1, 0, st>>>(h->epoch, h->ticket);
  k_halo_exchange<<<(unsigned)grid, 256, 0, st>>>(h->table, h->world, h->rank, h->capacity, h->epoch, h->ticket,
                                                  h->remote_offset, h->send_ptr, h->send_src, src, send_index, recv_index,
                                                  h->n_recv, dst, h->error);
  PHIFEM_CHECK_LAUNCH();
  return PHIFEM_OK;
}

extern "C" int phifem_halo_error(phifem_halo* h) {  // synchronises; 1 = a receive timed out
  if (!h) return 0;
  int e = 0;
  cudaMemcpy(&e, h->error, sizeof(int), cudaMemcpyDeviceToHost);
  return e;
}

extern "C" void phifem_halo_destroy(phifem_halo* h) {
  if (!h) return;
  cudaDeviceSynchronize();
  for (int q = 0; q < h->world; ++q)
    if (h->opened[q]) cudaIpcCloseMemHandle(h->peers[q]);
  cudaFree(h->base);
  cudaFree(h->table);
  cudaFree(h->remote_offset);
  cudaFree(h->send_ptr);
  cudaFree(h->send_src);
  cudaFree(h->epoch);
  cudaFree(h->ticket);
  cudaFree(h->error);
  delete h;
}
