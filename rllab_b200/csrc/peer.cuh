// Peer-memory exchange over NVLink / NVSwitch for the KB-sized float64 reduction vectors of an iteration (flat gradient,
// Fisher-vector products, loss / KL triples, advantage statistics, baseline normal equations).
//
// Every rank owns one exchange WINDOW in its own HBM (cudaMalloc + CUDA IPC, mapped by all peers):
//     flags [2][PEER_MAX_RANKS] u64 | done counter | slots [2][world][n_cap] float64
// One collective = one kernel on every rank:  push this rank's vector into slot[parity][rank] of EVERY window (remote
// stores through NVLink are fire-and-forget), release-store the sequence number into flag[parity][rank] of every window,
// spin (acquire) on the own window's flags until every peer's sequence number has arrived, then reduce the `world` slots
// of the own window in RANK ORDER (sums for index < n_sum, maxima above) -- every rank computes bit-identical results.
// The parity (sequence number & 1) double-buffers slots and flags: a rank can run at most one collective ahead of the
// slowest peer (it needs that peer's flag of the current collective to finish it), so two buffers suffice.
//
// The same exchange is FUSED into the finalize of the policy-update passes (common.cu, finalize_update_kernel): the blocks
// that fold the per-block partials of a gradient / Fisher-vector pass push their 32 outputs straight into the peers'
// windows and reduce over ranks in the same launch -- compute step and collective are one kernel, no NCCL call, no
// intermediate vector in between.
//
// No counterpart in the reference (its update is single-process).
#pragma once
#include "common.cuh"

namespace b200rl {

constexpr size_t PEER_FLAGS_BYTES = 2 * PEER_MAX_RANKS * 8;          // flags [2][16]
constexpr size_t PEER_SLOT_OFFSET = 512;                             // flags, done counter, padding
// PeerArgs: common.cuh

__device__ __forceinline__ unsigned long long* peer_flag(unsigned char* win, int par, int r) {
  return reinterpret_cast<unsigned long long*>(win) + par * PEER_MAX_RANKS + r;
}
__device__ __forceinline__ unsigned int* peer_done_counter(unsigned char* win) {
  return reinterpret_cast<unsigned int*>(win + PEER_FLAGS_BYTES);
}
// number of collectives of this rank that gave up waiting for a peer (their results were poisoned with NaN); the host
// reads it back through b200rl_peer_timeouts()
__device__ __forceinline__ unsigned int* peer_timeout_counter(unsigned char* win) {
  return reinterpret_cast<unsigned int*>(win + PEER_FLAGS_BYTES + 8);
}
__device__ __forceinline__ double* peer_slot(const PeerArgs& p, int win_rank, int par, int src_rank) {
  return reinterpret_cast<double*>(p.win[win_rank] + PEER_SLOT_OFFSET) + ((size_t)par * p.world + src_rank) * p.n_cap;
}
__device__ __forceinline__ void peer_st_release(unsigned long long* addr, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long peer_ld_acquire(const unsigned long long* addr) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(addr) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long peer_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// signal every window that this rank's slot of collective p.seq is complete (call from ONE thread per peer, after the
// pushing threads have fenced and the block / grid has synchronised)
__device__ __forceinline__ void peer_signal(const PeerArgs& p, int peer) {
  peer_st_release(peer_flag(p.win[peer], (int)(p.seq & 1ull), p.rank), p.seq);
}
// wait until rank `src` has signalled collective p.seq in the OWN window; bounded (30 s of %globaltimer) so that a dead
// peer poisons the result instead of hanging the GPU.  Returns false on time-out.
__device__ __forceinline__ bool peer_wait(const PeerArgs& p, int src) {
  const unsigned long long* f = peer_flag(p.win[p.rank], (int)(p.seq & 1ull), src);
  const unsigned long long t0 = peer_globaltimer();
  while (peer_ld_acquire(f) < p.seq) {
    if (peer_globaltimer() - t0 > 30000000000ull) {
      atomicAdd(peer_timeout_counter(p.win[p.rank]), 1u);
      return false;
    }
  }
  return true;
}

}  // namespace b200rl
