// Gathering the shards of a sharded query over NVLink / NVSwitch peer memory (BASELINE config 4).
//
// Every rank owns rows [lo, hi) of the global `body_pos` / `obs` tensors.  The gathered tensors live in SYMMETRIC
// memory: the same allocation on every GPU, mapped into every process (peer pointers) and -- where the NVSwitch
// fabric offers it -- behind one MULTICAST address, a store to which the switch replicates into all GPUs.
//
// Two ways to fill them, both ending in the same release / acquire hand-shake:
//   direct   the query kernel's own output pointers ARE the multicast addresses of this rank's rows (a store to a
//            multicast address is an ordinary STG in SASS; the address mapping does the replication), followed by
//            parc_peer_barrier;
//   push     the query writes its shard locally, parc_peer_push re-reads it (still L2-resident) with 16-byte
//            loads and stores it to the multicast address -- or, without multicast, to each peer pointer in turn --
//            and signals from the same kernel.
//
// Hand-shake: signal slot s is a uint64 counter at the same offset of every rank's symmetric buffer.  A block that
// has finished its stores adds 1 to slot s on EVERY rank (one multimem.red on the multicast address, or one
// red.release.sys per peer pointer) and then spins on its LOCAL copy until it reaches world * epoch: every rank's
// block s has then finished, and -- because the adds are releases ordered behind the block's stores and the spin is
// an acquire -- their data is visible here.  `epoch` is a per-slot device counter private to the rank, so the same
// launch can be replayed from a CUDA graph.
//
// The reference is single-GPU; this replaces nothing in it.  The NCCL all-gather of parc_b200/sharding.py is the
// baseline these kernels are measured against (bench.py, cfg4).
#include "parc_common.cuh"

namespace parc {

#define PEER_THREADS 256

struct PeerSync {
  uint64_t* mc_signal;                      // multicast address of the signal slots, or nullptr
  uint64_t* peer_signal[PARC_MAX_PEERS];    // per-rank addresses of the signal slots (used when mc_signal is null)
  uint64_t* local_signal;                   // this rank's copy
  uint64_t* epoch;                          // [slots] private per-rank launch counters
  int world;
  int rank;
  long long timeout_ns;                     // > 0: give up waiting after this long (and say so in *timeout_flag)
  int32_t* timeout_flag;
};

__device__ __forceinline__ void st_mc_v4(float4* p, const float4& v) {
  asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_mc_f32(float* p, float v) {
  asm volatile("multimem.st.weak.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// Called by ONE thread of a block after a __syncthreads() that follows the block's stores.
__device__ __forceinline__ void signal_and_wait(const PeerSync& s, int slot) {
  const uint64_t e = s.epoch[slot] + 1;
  asm volatile("fence.acq_rel.sys;" ::: "memory");
  if (s.mc_signal) {
    asm volatile("multimem.red.release.sys.global.add.u64 [%0], %1;" ::"l"(s.mc_signal + slot), "l"((uint64_t)1) : "memory");
  } else {
    for (int r = 0; r < s.world; ++r)
      asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(s.peer_signal[r] + slot), "l"((uint64_t)1) : "memory");
  }
  const uint64_t want = e * (uint64_t)s.world;
  uint64_t seen, t0 = 0, now;
  if (s.timeout_ns > 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(s.local_signal + slot) : "memory");
    if (seen >= want) break;
    if (s.timeout_ns > 0) {                  // a rank that never arrives must not hang this GPU
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (now - t0 > (uint64_t)s.timeout_ns) {
        if (s.timeout_flag) atomicOr(s.timeout_flag, 1);
        break;
      }
    }
  }
  s.epoch[slot] = e;
}

__global__ void __launch_bounds__(32) peer_barrier_kernel(const __grid_constant__ PeerSync s, int slot) {
  if (threadIdx.x == 0) signal_and_wait(s, slot);
}

struct PushParams {
  ParcPeerSegment seg[PARC_MAX_PUSH_SEGMENTS];
  int num_segments;
  int vec16;                                // every segment is 16-byte aligned and a multiple of 16 bytes long
  PeerSync sync;
};

__global__ void __launch_bounds__(PEER_THREADS) peer_push_kernel(const __grid_constant__ PushParams p) {
  const int64_t t0 = (int64_t)blockIdx.x * PEER_THREADS + threadIdx.x;
  const int64_t nt = (int64_t)gridDim.x * PEER_THREADS;
  for (int k = 0; k < p.num_segments; ++k) {
    const ParcPeerSegment& sg = p.seg[k];
    if (p.vec16) {
      const float4* __restrict__ src = reinterpret_cast<const float4*>(sg.src);
      const int64_t n = sg.bytes >> 4;
      if (sg.dst_multicast) {
        float4* dst = reinterpret_cast<float4*>(sg.dst_multicast);
        int64_t i = t0;
        for (; i + 3 * nt < n; i += 4 * nt) {       // four independent 16-byte loads in flight per thread
          const float4 a = src[i], b = src[i + nt], c = src[i + 2 * nt], d = src[i + 3 * nt];
          st_mc_v4(dst + i, a); st_mc_v4(dst + i + nt, b); st_mc_v4(dst + i + 2 * nt, c); st_mc_v4(dst + i + 3 * nt, d);
        }
        for (; i < n; i += nt) st_mc_v4(dst + i, src[i]);
      } else {
        for (int64_t i = t0; i < n; i += nt) {
          const float4 v = src[i];
          // start behind the own rank and go round: at any moment the ranks write into different peers
          for (int k = 1; k <= p.sync.world; ++k) {
            int r = p.sync.rank + k;
            if (r >= p.sync.world) r -= p.sync.world;
            reinterpret_cast<float4*>(sg.dst_peer[r])[i] = v;
          }
        }
      }
    } else {
      const float* __restrict__ src = reinterpret_cast<const float*>(sg.src);
      const int64_t n = sg.bytes >> 2;
      for (int64_t i = t0; i < n; i += nt) {
        const float v = src[i];
        if (sg.dst_multicast) st_mc_f32(reinterpret_cast<float*>(sg.dst_multicast) + i, v);
        else
          for (int k = 1; k <= p.sync.world; ++k) {
            int r = p.sync.rank + k;
            if (r >= p.sync.world) r -= p.sync.world;
            reinterpret_cast<float*>(sg.dst_peer[r])[i] = v;
          }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) signal_and_wait(p.sync, blockIdx.x);
}

static int fill_sync(const ParcPeerSignals* sig, int slots_needed, PeerSync* out) {
  if (!sig) return PARC_E_NULL;
  if (sig->world < 1 || sig->world > PARC_MAX_PEERS || sig->num_slots < slots_needed) return PARC_E_SIZE;
  if (!sig->local_signal || !sig->epoch) return PARC_E_NULL;
  PeerSync s = {};
  s.mc_signal = sig->multicast_signal;
  s.local_signal = sig->local_signal;
  s.epoch = sig->epoch;
  s.world = sig->world;
  s.rank = sig->rank;
  if (sig->rank < 0 || sig->rank >= sig->world) return PARC_E_SIZE;
  s.timeout_ns = sig->timeout_ns;
  s.timeout_flag = sig->timeout_flag;
  if (sig->timeout_ns < 0) return PARC_E_SIZE;
  if (reinterpret_cast<uintptr_t>(s.timeout_flag) & 3u) return PARC_E_ALIGN;
  if (!s.mc_signal)
    for (int r = 0; r < sig->world; ++r) {
      if (!sig->peer_signal[r]) return PARC_E_NULL;
      s.peer_signal[r] = sig->peer_signal[r];
    }
  if ((reinterpret_cast<uintptr_t>(s.local_signal) & 7u) || (reinterpret_cast<uintptr_t>(s.epoch) & 7u)) return PARC_E_ALIGN;
  *out = s;
  return PARC_OK;
}

}  // namespace parc

using namespace parc;

extern "C" int parc_peer_barrier(const ParcPeerSignals* signals, int32_t slot, void* stream) {
  PeerSync s;
  if (slot < 0) return PARC_E_SIZE;
  const int rc = fill_sync(signals, slot + 1, &s);
  if (rc) return rc;
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(s, slot);
  return check_launch();
}

extern "C" int parc_peer_push(const ParcPeerSegment* segments, int32_t num_segments, const ParcPeerSignals* signals,
                              int32_t num_blocks, void* stream) {
  if (!segments) return PARC_E_NULL;
  if (num_segments < 1 || num_segments > PARC_MAX_PUSH_SEGMENTS) return PARC_E_SIZE;
  if (num_blocks < 1) num_blocks = 64;
  PushParams p = {};
  const int rc = fill_sync(signals, num_blocks, &p.sync);
  if (rc) return rc;
  p.num_segments = num_segments;
  p.vec16 = 1;
  for (int k = 0; k < num_segments; ++k) {
    const ParcPeerSegment& sg = segments[k];
    if (sg.bytes < 0 || (sg.bytes & 3)) return PARC_E_SIZE;
    if (sg.bytes > 0 && !sg.src) return PARC_E_NULL;
    uintptr_t bits = reinterpret_cast<uintptr_t>(sg.src) | (uintptr_t)sg.bytes;
    if (sg.dst_multicast) {
      bits |= reinterpret_cast<uintptr_t>(sg.dst_multicast);
    } else {
      for (int r = 0; r < p.sync.world; ++r) {
        if (sg.bytes > 0 && !sg.dst_peer[r]) return PARC_E_NULL;
        bits |= reinterpret_cast<uintptr_t>(sg.dst_peer[r]);
      }
    }
    if (bits & 3u) return PARC_E_ALIGN;
    if (bits & 15u) p.vec16 = 0;
    p.seg[k] = sg;
  }
  peer_push_kernel<<<num_blocks, PEER_THREADS, 0, (cudaStream_t)stream>>>(p);
  return check_launch();
}
