// pgo_peer.cuh -- halo exchange, slice all-gather and scalar all-reduce of the partitioned PCG iteration over NVLink
// PEER MEMORY instead of NCCL calls.  Included by pgo_b200.cu (before pgo_amg.cuh).
//
// Why: a PCG iteration of the row-partitioned multilevel solver is ~0.2 ms of partitioned work at 8 GPUs, and it needs
// five halo exchanges, one slice gather and one 2-scalar all-reduce.  As NCCL calls those are seven host-enqueued
// collectives of 25-35 us each (latency, not bandwidth: the payloads are tens of KB), and NCCL keeps the iteration out
// of a CUDA graph.  Here every rank owns a WINDOW of device memory (flags + double-buffered staging), exported with
// cudaIpcGetMemHandle and mapped by every other rank of the box; an exchange is
//     push:  pack my boundary values and STORE them straight into the neighbours' staging over NVLink, system-scope
//            fence, then a release store of the exchange's sequence number into the neighbour's flag slot
//     wait:  spin (acquire loads, local memory) until every source's flag carries this sequence number, then copy the
//            staging into the halo tail of the vector
// -- two small kernels, no host involvement, no NCCL: the whole multi-GPU iteration is captured into a CUDA graph.
// The scalar all-reduce is "every rank stores its partial sums into slot [rank] of every peer, then each rank adds the
// slots in rank order": bit-identical results on all ranks (the solver relies on that, see pgo_amg.cuh).
//
// Channels.  A channel is one recurring exchange with a fixed, symmetric set of partners (the halo of one distributed
// level, the residual gather of the first replicated level, the scalar all-reduce).  Every channel has its own
// sequence counter, flag row and double-buffered staging, which is what makes two buffers enough: partner k cannot
// push exchange t+2 before it has seen my push t+1, which I issue (stream order) after my wait t.
#pragma once

namespace pgo {

constexpr int kPeerMaxChannels = 6;
constexpr int kPeerMaxWorld = 16;
constexpr int kPeerThreads = 256;

// what a rank publishes about its window (all-gathered once at setup through NCCL)
struct PeerLayout {
  cudaIpcMemHandle_t handle;
  unsigned long long bytes;
  unsigned long long flag_off[kPeerMaxChannels];     // byte offset of unsigned int flags[kPeerMaxWorld] (indexed by sender rank)
  unsigned long long stage_off[kPeerMaxChannels];    // byte offset of double staging[2][stage_cap]
  unsigned long long stage_cap[kPeerMaxChannels];    // doubles per parity
  int recv_item_off[kPeerMaxChannels][kPeerMaxWorld]; // item (6 doubles) offset where sender r's values land; -1: r is no source
  int ok;
};

struct PeerTarget {
  double* stage[2];        // REMOTE staging of the partner, parity 0 / 1
  unsigned int* flag;      // REMOTE flag slot [my rank] of the partner
  int dst_item;            // where my items start in its staging
  int src0, src1;          // my items [src0, src1) go to this partner (positions in the index list / the contiguous range)
};
struct PeerPushArgs {
  int n_targets;
  unsigned int* seq;       // local: number of completed pushes on this channel
  unsigned int* ticket;    // local [1 + targets]: finished partners / finished slices of each partner of the running push
  PeerTarget t[kPeerMaxWorld - 1];
};
struct PeerWaitArgs {
  int n_sources;
  int src_rank[kPeerMaxWorld - 1];
  const unsigned int* flags;     // local flag row of the channel
  const unsigned int* seq;
  const double* stage[2];        // local staging
  int n_copies;
  int copy_src[2], copy_dst[2], copy_n[2];   // in doubles: v[copy_dst + e] = stage[copy_src + e]
  int* timeout_flag;             // set when a partner never arrived (the solve then fails instead of hanging the GPU)
};

__device__ __forceinline__ void st_release_sys_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long peer_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
constexpr unsigned long long kPeerTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

// Pack + remote store + flag.  idx: my local node of list position k (nullptr: node k itself); items are 6 doubles.
// gridDim = (slices, targets): the CTAs of column q serve partner q alone -- its values, ONE system-scope fence per
// CTA, and the flag as soon as the partner's last slice is through (the partners' fences run side by side instead of
// one after the other; measured 12 -> ~6 us per push at 8 GPUs).  ticket[q + 1] counts the finished slices of partner
// q, ticket[0] the finished partners.
__global__ void __launch_bounds__(kPeerThreads) peer_push_kernel(const PeerPushArgs A, const int* __restrict__ idx,
                                                                 const double* __restrict__ v, const int* skip) {
  if (skip && *skip) return;
  const unsigned int s = *A.seq + 1;       // read by every CTA before the last one to finish bumps it
  const int q = blockIdx.y;
  const PeerTarget& T = A.t[q];
  double* dst = T.stage[s & 1] + 6 * (size_t)T.dst_item;
  const int n = 6 * (T.src1 - T.src0);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const int k = T.src0 + e / 6, c = e % 6;
    const int node = idx ? __ldg(idx + k) : k;
    dst[e] = v[6 * (size_t)node + c];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    bool last_slice = true;
    if (gridDim.x > 1) last_slice = atomicAdd(A.ticket + 1 + q, 1u) == gridDim.x - 1;
    if (last_slice) {
      if (gridDim.x > 1) { __threadfence_system(); A.ticket[1 + q] = 0; }
      st_relaxed_sys_u32(T.flag, s);
      if (atomicAdd(A.ticket, 1u) == gridDim.y - 1) { A.ticket[0] = 0; *A.seq = s; }
    }
  }
}

__device__ __forceinline__ bool peer_spin(const unsigned int* flag, unsigned int s) {
  const unsigned long long t0 = peer_now_ns();
  int polls = 0;
  while ((int)(ld_acquire_sys_u32(flag) - s) < 0) {
    if (((++polls) & 1023) == 0 && peer_now_ns() - t0 > kPeerTimeoutNs) return false;
  }
  return true;
}

// Wait for every source's push of this exchange, then unpack the staging into the vector.
__global__ void __launch_bounds__(kPeerThreads) peer_wait_kernel(const PeerWaitArgs A, double* __restrict__ v, const int* skip) {
  if (skip && *skip) return;
  const unsigned int s = *A.seq;           // my own push of this exchange is complete (stream order)
  if (threadIdx.x < A.n_sources) {
    if (!peer_spin(A.flags + A.src_rank[threadIdx.x], s) && A.timeout_flag) *A.timeout_flag = 1;
  }
  __syncthreads();
  const double* stage = A.stage[s & 1];
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
  for (int c = 0; c < A.n_copies; ++c)
    for (int e = gtid; e < A.copy_n[c]; e += gsize) v[A.copy_dst[c] + (size_t)e] = __ldcg(stage + A.copy_src[c] + (size_t)e);
}

// (The scalar all-reduce: every rank pushes its partial sums as ONE item into slot [my rank] of every peer;
// amg_pcg_scalar_peer_kernel in pgo_amg.cuh waits for all slots, adds them in rank order and takes the CG step.)

}  // namespace pgo

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct PeerChannelSpec {                  // what the caller wants of a channel, in its own terms
  size_t stage_items = 0;                 // items (6 doubles) per parity of MY staging
  std::vector<int> recv_item_off;         // [world]: where sender r's items land in my staging, -1 = r sends me nothing
};

struct PeerCtx {
  int world = 1, rank = 0, n_channels = 0;
  void* window = nullptr;
  size_t bytes = 0;
  std::vector<void*> base;                // [world] mapped windows (base[rank] == window)
  std::vector<pgo::PeerLayout> layout;    // [world]
  unsigned long long seq_off = 0, ticket_off = 0, timeout_off = 0;
  int* barrier_buf = nullptr;
  long long pushes = 0, push_bytes = 0;   // statistics of this rank
  unsigned int* seq(int ch) const { return reinterpret_cast<unsigned int*>(static_cast<char*>(window) + seq_off) + ch; }
  unsigned int* ticket(int ch) const { return reinterpret_cast<unsigned int*>(static_cast<char*>(window) + ticket_off) + (size_t)ch * pgo::kPeerMaxWorld; }
  int* timeout_flag() const { return reinterpret_cast<int*>(static_cast<char*>(window) + timeout_off); }
  unsigned int* flags(int r, int ch) const { return reinterpret_cast<unsigned int*>(static_cast<char*>(base[r]) + layout[r].flag_off[ch]); }
  double* stage(int r, int ch, int parity) const {
    return reinterpret_cast<double*>(static_cast<char*>(base[r]) + layout[r].stage_off[ch]) + (size_t)parity * layout[r].stage_cap[ch];
  }
};

static void peer_destroy(pgo_graph* g, PeerCtx* P);

// Collective over the graph's communicator.  *out stays nullptr (and PGO_OK is returned) when peer windows are not
// available on this box -- every rank takes the same decision -- and the callers keep using NCCL.
// local_ok == false (this rank's plans do not qualify): the rank still takes part in the collectives and vetoes.
static int peer_create(pgo_graph* g, const std::vector<PeerChannelSpec>& ch, PeerCtx** out, bool local_ok = true) {
  using namespace pgo;
  *out = nullptr;
  if (g->world <= 1 || g->world > kPeerMaxWorld || (int)ch.size() > kPeerMaxChannels) return PGO_OK;
  if (const char* e = getenv("PGO_PEER")) if (atoi(e) == 0) return PGO_OK;
  PeerCtx* P = new PeerCtx();
  P->world = g->world; P->rank = g->rank; P->n_channels = (int)ch.size();
  PeerLayout mine;
  std::memset(&mine, 0, sizeof mine);
  auto align = [](unsigned long long x) { return (x + 255ull) & ~255ull; };
  unsigned long long off = 0;
  P->seq_off = off; off = align(off + kPeerMaxChannels * sizeof(unsigned int));
  P->ticket_off = off; off = align(off + kPeerMaxChannels * kPeerMaxWorld * sizeof(unsigned int));
  P->timeout_off = off; off = align(off + sizeof(int));
  for (int c = 0; c < P->n_channels; ++c) {
    mine.flag_off[c] = off; off = align(off + kPeerMaxWorld * sizeof(unsigned int));
    mine.stage_cap[c] = 6 * (unsigned long long)std::max<size_t>(ch[c].stage_items, 1);
    mine.stage_off[c] = off; off = align(off + 2 * mine.stage_cap[c] * sizeof(double));
    for (int r = 0; r < kPeerMaxWorld; ++r) mine.recv_item_off[c][r] = r < (int)ch[c].recv_item_off.size() ? ch[c].recv_item_off[r] : -1;
  }
  mine.bytes = off;
  mine.ok = 1;
  P->bytes = (size_t)off;
  bool ok = cudaMalloc(&P->window, P->bytes) == cudaSuccess;
  if (ok) ok = cudaMalloc(reinterpret_cast<void**>(&P->barrier_buf), sizeof(int)) == cudaSuccess;
  if (ok) ok = cudaMemsetAsync(P->barrier_buf, 0, sizeof(int), g->stream) == cudaSuccess;
  if (ok) ok = cudaMemsetAsync(P->window, 0, P->bytes, g->stream) == cudaSuccess;
  if (ok) ok = cudaIpcGetMemHandle(&mine.handle, P->window) == cudaSuccess;
  if (!ok) { cudaGetLastError(); mine.ok = 0; }
  if (!local_ok) mine.ok = 0;
  // all-gather the layouts (device staging through the pool)
  PeerLayout* dev = nullptr;
  PGO_TRY(dev_alloc(g, &dev, (size_t)g->world + 1));
  CUDA_TRY(cudaMemcpyAsync(dev + g->world, &mine, sizeof mine, cudaMemcpyHostToDevice, g->stream));
  NCCL_TRY(ncclAllGather(dev + g->world, dev, sizeof(PeerLayout), ncclChar, g->comm, g->stream));
  P->layout.resize(g->world);
  CUDA_TRY(cudaMemcpyAsync(P->layout.data(), dev, sizeof(PeerLayout) * g->world, cudaMemcpyDeviceToHost, g->stream));
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  bool all_ok = true;
  for (int r = 0; r < g->world; ++r) all_ok = all_ok && P->layout[r].ok == 1;
  P->base.assign(g->world, nullptr);
  int opened = 1;
  if (all_ok) {
    P->base[g->rank] = P->window;
    for (int r = 0; r < g->world && opened; ++r) {
      if (r == g->rank) continue;
      if (cudaIpcOpenMemHandle(&P->base[r], P->layout[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        P->base[r] = nullptr;
        opened = 0;
      }
    }
  } else {
    opened = 0;
  }
  // consensus: one rank that could not map a window sends everybody back to NCCL (this all-reduce also orders every
  // rank's window memset before anybody's first push)
  int* flag_d = nullptr;
  PGO_TRY(dev_alloc(g, &flag_d, 1));
  CUDA_TRY(cudaMemcpyAsync(flag_d, &opened, sizeof(int), cudaMemcpyHostToDevice, g->stream));
  NCCL_TRY(ncclAllReduce(flag_d, flag_d, 1, ncclInt, ncclMin, g->comm, g->stream));
  CUDA_TRY(cudaMemcpyAsync(&opened, flag_d, sizeof(int), cudaMemcpyDeviceToHost, g->stream));
  CUDA_TRY(cudaStreamSynchronize(g->stream));
  if (!opened) {
    if (getenv("PGO_PROFILE_HOST")) fprintf(stderr, "[pgo peer] rank %d: peer windows unavailable, staying with NCCL\n", g->rank);
    peer_destroy(g, P);
    return PGO_OK;
  }
  *out = P;
  return PGO_OK;
}

static void peer_destroy(pgo_graph* g, PeerCtx* P) {
  if (!P) return;
  for (int r = 0; r < (int)P->base.size(); ++r)
    if (r != P->rank && P->base[r]) cudaIpcCloseMemHandle(P->base[r]);
  // nobody may still be writing into my window: a last collective before it goes (skipped after a failure)
  if (g->comm && !g->failed && P->barrier_buf) {
    if (ncclAllReduce(P->barrier_buf, P->barrier_buf, 1, ncclInt, ncclSum, g->comm, g->stream) == ncclSuccess) cudaStreamSynchronize(g->stream);
  }
  if (P->barrier_buf) cudaFree(P->barrier_buf);
  if (P->window) cudaFree(P->window);
  cudaGetLastError();
  delete P;
}

// push args of a channel towards partners `nbr` (ranks), partner q receiving my items [src0[q], src1[q])
static pgo::PeerPushArgs peer_push_args(const PeerCtx* P, int ch, const std::vector<int>& nbr, const std::vector<int>& src0,
                                        const std::vector<int>& src1) {
  pgo::PeerPushArgs A;
  std::memset(&A, 0, sizeof A);
  A.seq = P->seq(ch); A.ticket = P->ticket(ch);
  for (size_t q = 0; q < nbr.size(); ++q) {
    const int r = nbr[q];
    if (src1[q] <= src0[q]) continue;
    pgo::PeerTarget& T = A.t[A.n_targets++];
    T.stage[0] = P->stage(r, ch, 0); T.stage[1] = P->stage(r, ch, 1);
    T.flag = P->flags(r, ch) + P->rank;
    T.dst_item = P->layout[r].recv_item_off[ch][P->rank];
    T.src0 = src0[q]; T.src1 = src1[q];
  }
  return A;
}
static pgo::PeerWaitArgs peer_wait_args(const PeerCtx* P, int ch, const std::vector<int>& sources) {
  pgo::PeerWaitArgs A;
  std::memset(&A, 0, sizeof A);
  A.n_sources = (int)sources.size();
  for (size_t q = 0; q < sources.size(); ++q) A.src_rank[q] = sources[q];
  A.flags = P->flags(P->rank, ch);
  A.seq = P->seq(ch);
  A.stage[0] = P->stage(P->rank, ch, 0); A.stage[1] = P->stage(P->rank, ch, 1);
  A.timeout_flag = P->timeout_flag();
  return A;
}
