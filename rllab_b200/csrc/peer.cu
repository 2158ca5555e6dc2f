// Exchange windows (CUDA IPC), the bound communicator and the stand-alone mixed all-reduce; see peer.cuh.
#include "peer.cuh"

namespace b200rl {

static PeerArgs g_peer = {};          // bound communicator of this process (one process per GPU); world == 0: none
static bool g_peer_fuse = false;

bool peer_fused() { return g_peer.world > 1 && g_peer_fuse; }
bool peer_bound() { return g_peer.world > 1; }
// arguments of the next collective: every call consumes one sequence number (all ranks issue the same call sequence)
PeerArgs peer_next() {
  PeerArgs p = g_peer;
  p.seq = ++g_peer.seq;
  return p;
}

// t[i] <- reduce over ranks (sum for i < n_sum, max above), in place, one CTA
__global__ void __launch_bounds__(1024) peer_allreduce_kernel(PeerArgs p, double* __restrict__ t, long long n,
                                                              long long n_sum) {
  __shared__ int ok_s;
  const int par = (int)(p.seq & 1ull), tid = threadIdx.x;
  if (tid == 0) ok_s = 1;
  for (long long i = tid; i < n; i += blockDim.x) {
    const double v = t[i];
    for (int r = 0; r < p.world; ++r) peer_slot(p, r, par, p.rank)[i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (tid < p.world) {
    peer_signal(p, tid);
    if (!peer_wait(p, tid)) ok_s = 0;
  }
  __syncthreads();
  const bool ok = ok_s != 0;
  for (long long i = tid; i < n; i += blockDim.x) {
    double acc = peer_slot(p, p.rank, par, 0)[i];
    for (int r = 1; r < p.world; ++r) {
      const double v = peer_slot(p, p.rank, par, r)[i];
      acc = i < n_sum ? acc + v : fmax(acc, v);
    }
    t[i] = ok ? acc : __longlong_as_double(0x7FF8000000000000ll);    // a peer never arrived: poison
  }
}

}  // namespace b200rl

using namespace b200rl;

extern "C" {

long long b200rl_peer_window_bytes(int world, long long n_cap) {
  if (world < 1 || world > PEER_MAX_RANKS || n_cap < 1) return -1;
  return (long long)PEER_SLOT_OFFSET + 2ll * world * n_cap * (long long)sizeof(double);
}

int b200rl_peer_window_create(int world, long long n_cap, void** window_out, unsigned char* handle_out) {
  B200RL_REQUIRE(world >= 1 && world <= PEER_MAX_RANKS && n_cap >= 1 && window_out && handle_out,
                 "peer_window_create: bad arguments");
  void* w = nullptr;
  const size_t bytes = (size_t)b200rl_peer_window_bytes(world, n_cap);
  B200RL_CUDA_CHECK(cudaMalloc(&w, bytes));
  B200RL_CUDA_CHECK(cudaMemset(w, 0, bytes));
  B200RL_CUDA_CHECK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  static_assert(sizeof(h) == B200RL_IPC_HANDLE_BYTES, "IPC handle size");
  B200RL_CUDA_CHECK(cudaIpcGetMemHandle(&h, w));
  memcpy(handle_out, &h, sizeof(h));
  *window_out = w;
  return 0;
}

int b200rl_peer_window_open(const unsigned char* handle, void** window_out) {
  B200RL_REQUIRE(handle && window_out, "peer_window_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  B200RL_CUDA_CHECK(cudaIpcOpenMemHandle(window_out, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int b200rl_peer_window_close(void* window) {
  if (window) B200RL_CUDA_CHECK(cudaIpcCloseMemHandle(window));
  return 0;
}

int b200rl_peer_window_destroy(void* window) {
  if (window) B200RL_CUDA_CHECK(cudaFree(window));
  return 0;
}

int b200rl_peer_bind(void* const* windows, int rank, int world, long long n_cap) {
  if (windows == nullptr || world <= 1) {
    g_peer = PeerArgs{};
    g_peer_fuse = false;
    return 0;
  }
  B200RL_REQUIRE(world <= PEER_MAX_RANKS && rank >= 0 && rank < world && n_cap >= 1, "peer_bind: bad arguments");
  PeerArgs p = {};
  for (int r = 0; r < world; ++r) {
    B200RL_REQUIRE(windows[r] != nullptr, "peer_bind: window %d is NULL", r);
    p.win[r] = reinterpret_cast<unsigned char*>(windows[r]);
  }
  p.rank = rank; p.world = world; p.n_cap = n_cap; p.seq = 0;
  g_peer = p;
  return 0;
}

int b200rl_peer_fuse_updates(int enable) {
  B200RL_REQUIRE(!enable || peer_bound(), "peer_fuse_updates: no communicator bound");
  g_peer_fuse = enable != 0;
  return 0;
}

int b200rl_peer_timeouts(unsigned int* count_out_host) {
  B200RL_REQUIRE(count_out_host != nullptr, "peer_timeouts: bad arguments");
  *count_out_host = 0;
  if (!peer_bound()) return 0;
  B200RL_CUDA_CHECK(cudaMemcpy(count_out_host, g_peer.win[g_peer.rank] + PEER_FLAGS_BYTES + 8, sizeof(unsigned int),
                               cudaMemcpyDeviceToHost));       // synchronising: call it outside the hot path
  return 0;
}

int b200rl_peer_allreduce_mixed(double* t, long long n, long long n_sum, void* stream) {
  B200RL_REQUIRE(peer_bound(), "peer_allreduce_mixed: no communicator bound");
  B200RL_REQUIRE(t && n >= 1 && n <= g_peer.n_cap && n_sum >= 0 && n_sum <= n, "peer_allreduce_mixed: bad arguments");
  const PeerArgs p = peer_next();
  peer_allreduce_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(p, t, n, n_sum);
  B200RL_LAUNCH_CHECK("peer_allreduce_kernel");
  return 0;
}
}
