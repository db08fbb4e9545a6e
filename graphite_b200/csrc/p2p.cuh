// p2p.cuh — the multi-GPU exchange step of the point-partitioned path, written over NVLink peer memory.
//
// Every camera-sized reduction of the path (diag(B) and g_c after linearize, the 54 Schur-diagonal sums, S p once per
// PCG iteration, cost and rho scalars) is a sum over ranks of a small vector (128 KB - 6 MB).  At that size an
// all-reduce is pure latency, so it is done as a ONE-SHOT ALL-GATHER + LOCAL SUM fused into the kernels either side:
//   - the PRODUCER (e.g. the per-camera reduction of the partial rows) stores its values straight into every rank's
//     receive area through peer-mapped pointers (cudaIpc), rank r's slot;
//   - the last CTA of the producer publishes the exchange's epoch number to every peer's flag word (release, system scope);
//   - the CONSUMER (e.g. the PCG vector update) waits for the epoch on its own flag words (acquire, system scope) and adds
//     the nranks slots in RANK ORDER, so every rank computes bit-identical sums -> identical PCG scalars and LM
//     decisions on all ranks with no broadcast, and run-to-run reproducible results (NCCL's ring order is not used).
// The receive area has two halves indexed by epoch parity: rank A can push epoch e+1 while B still reads e, and A cannot
// reach e+2 before B has pushed e+1, which B does only after consuming e (same stream).  The epoch is a DEVICE-side
// sequence number advanced by every exchange that actually runs: producers that return early (PCG already stopped -
// uniformly on all ranks, since all ranks hold identical scalars) do not consume one, which keeps the parity argument
// valid whatever the host enqueued.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gb {

constexpr int P2P_MAX_RANKS = 8; // one NVSwitch box

struct P2P {
  int nranks, rank;
  unsigned char *recv[P2P_MAX_RANKS];        // receive area of every rank (peer-mapped; [rank] is local)
  unsigned long long *flags[P2P_MAX_RANKS];  // flag words of every rank: flags[dst][src] = last epoch src pushed to dst
  unsigned long long half_bytes, slot_bytes; // receive area = [2 halves][nranks slots][slot_bytes]
  unsigned int *counter;                     // local: CTAs of the running producer that have finished pushing
  unsigned long long *seq;                   // local: epoch of the last exchange this rank has pushed
  int *error;                                // local (pinned host): set when a wait timed out (a peer died)
  unsigned long long ll_off, ll_half_bytes, ll_slot_bytes; // receive area + ll_off = the LL area [2 halves][nranks][slot]
  unsigned long long timeout_ns;             // how long a consumer waits for a peer before it gives up (GB_P2P_TIMEOUT_S)
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// relaxed system-scope store: after ONE __threadfence_system() the flag words of all peers are written back to back
// (fence + relaxed store is a release pattern); st.release on every store would serialise one NVLink round trip per peer
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// slot of rank `src` inside rank `dst`'s receive area for this epoch
template <typename T> __device__ __forceinline__ T *p2p_slot(const P2P &pp, int dst, int src, unsigned long long epoch) {
  return reinterpret_cast<T *>(pp.recv[dst] + (epoch & 1ull) * pp.half_bytes + (unsigned long long)src * pp.slot_bytes);
}

// epoch of the exchange a producer is about to push (read by every thread BEFORE its CTA signals) / of the exchange a
// consumer has to wait for (the producer before it in the stream has advanced seq)
__device__ __forceinline__ unsigned long long p2p_next_epoch(const P2P &pp) {
  return *reinterpret_cast<volatile unsigned long long *>(pp.seq) + 1ull;
}
__device__ __forceinline__ unsigned long long p2p_current_epoch(const P2P &pp) {
  return *reinterpret_cast<volatile unsigned long long *>(pp.seq);
}

// Producer side, called by EVERY thread of EVERY CTA of the grid after its peer stores: the last CTA to arrive publishes
// the epoch to all peers.  (Pattern of the CUDA threadFenceReduction sample, with system-scope fences.)
__device__ __forceinline__ void p2p_signal(const P2P &pp, unsigned long long epoch, unsigned int nblocks) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(pp.counter, 1u);
    if (prev == nblocks - 1) {
      *pp.counter = 0u; // the next producer starts after this grid has ended (stream order)
      *pp.seq = epoch;  // every CTA has read seq before arriving
      __threadfence_system();
      for (int r = 0; r < pp.nranks; r++)
        if (r != pp.rank) st_relaxed_sys(pp.flags[r] + pp.rank, epoch);
    }
  }
}

// Consumer side, called by every thread of a CTA: returns when all peers have published `epoch`.
__device__ __forceinline__ void p2p_wait(const P2P &pp, unsigned long long epoch) {
  if (threadIdx.x < pp.nranks && (int)threadIdx.x != pp.rank) {
    const unsigned long long *f = pp.flags[pp.rank] + threadIdx.x;
    if (ld_acquire_sys(f) < epoch) {
      const unsigned long long t0 = global_timer_ns();
      while (ld_acquire_sys(f) < epoch) {
        if (global_timer_ns() - t0 > pp.timeout_ns) { // a peer is gone: fail loudly on the host (p2p_check) instead of hanging
          *pp.error = 1;
          break;
        }
      }
    }
  }
  __syncthreads();
}

// sum of the nranks slots in rank order (L2 loads: the slots are written by peers over NVLink)
template <typename T> __device__ __forceinline__ T p2p_sum(const P2P &pp, unsigned long long epoch, long long i) {
  T acc = __ldcg(p2p_slot<T>(pp, pp.rank, 0, epoch) + i);
  for (int r = 1; r < pp.nranks; r++) acc += __ldcg(p2p_slot<T>(pp, pp.rank, r, epoch) + i);
  return acc;
}

// ---------------------------------------------------------------------------------------------------------------------
// LL ("low latency") exchange of the PCG solve kernel.  The flag protocol above costs two serialised NVLink latencies
// (stores, system fence, flag store; measured 16 us per PCG iteration with the waits at 2-4 GPUs).  Here every 32-bit half
// of a value travels in ONE 8-byte word together with the 32-bit epoch of the exchange: an 8-byte store is indivisible, so
// a reader that sees the epoch has the data - no fence, no flag, one one-way latency.  The LL area is separate from the
// flag-protocol area (a plain value must never be mistaken for a stamped word) and starts zeroed (epoch 0 is never used).
// ---------------------------------------------------------------------------------------------------------------------
template <typename T> struct LLW { static constexpr int words = sizeof(T) / 4; }; // 8-byte words per value
__device__ __forceinline__ unsigned long long *ll_slot(const P2P &pp, int dst, int src, unsigned long long epoch) {
  return reinterpret_cast<unsigned long long *>(pp.recv[dst] + pp.ll_off + (epoch & 1ull) * pp.ll_half_bytes +
                                                (unsigned long long)src * pp.ll_slot_bytes);
}
__device__ __forceinline__ void ll_store_word(unsigned long long *p, unsigned int data, unsigned int epoch32) {
  const unsigned long long w = ((unsigned long long)epoch32 << 32) | (unsigned long long)data;
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned int ll_load_word(const P2P &pp, const unsigned long long *p, unsigned int epoch32) {
  unsigned long long w;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  if ((unsigned int)(w >> 32) != epoch32) {
    const unsigned long long t0 = global_timer_ns();
    unsigned int polls = 0;
    do {
      asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
      if ((++polls & 1023u) == 0u && global_timer_ns() - t0 > pp.timeout_ns) { // a peer is gone: fail loudly on the host
        *pp.error = 1;
        break;
      }
    } while ((unsigned int)(w >> 32) != epoch32);
  }
  return (unsigned int)w;
}
__device__ __forceinline__ void ll_store(unsigned long long *base, long long i, double v, unsigned int epoch32) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  ll_store_word(base + 2 * i, (unsigned int)b, epoch32);
  ll_store_word(base + 2 * i + 1, (unsigned int)(b >> 32), epoch32);
}
__device__ __forceinline__ void ll_store(unsigned long long *base, long long i, float v, unsigned int epoch32) {
  ll_store_word(base + i, __float_as_uint(v), epoch32);
}
template <typename T> __device__ __forceinline__ T ll_load(const P2P &pp, const unsigned long long *base, long long i, unsigned int epoch32);
template <> __device__ __forceinline__ double ll_load<double>(const P2P &pp, const unsigned long long *base, long long i, unsigned int epoch32) {
  const unsigned long long lo = ll_load_word(pp, base + 2 * i, epoch32), hi = ll_load_word(pp, base + 2 * i + 1, epoch32);
  return __longlong_as_double((long long)((hi << 32) | lo));
}
template <> __device__ __forceinline__ float ll_load<float>(const P2P &pp, const unsigned long long *base, long long i, unsigned int epoch32) {
  return __uint_as_float(ll_load_word(pp, base + i, epoch32));
}

// Sum over ranks of value i of the current exchange, in rank order.  All nranks x words loads are issued before any of
// them is tested (independent L2 loads, one latency); only words whose epoch has not arrived yet are polled again.
// (A loop of ll_load calls serialises 16 dependent ~0.6 us loads at 8 ranks: measured 9 us per PCG iteration.)
// __noinline__ with scalar arguments: its 16 in-flight words must not raise the register allocation of the product
// pipeline around it, and the P2P struct (kernel parameter) must not be copied to local memory for a reference.
__device__ __forceinline__ unsigned int ll_poll_word(const unsigned long long *p, unsigned int epoch32, unsigned long long timeout_ns, int *error) {
  unsigned long long w;
  const unsigned long long t0 = global_timer_ns();
  unsigned int polls = 0;
  do {
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    if ((++polls & 1023u) == 0u && global_timer_ns() - t0 > timeout_ns) { // a peer is gone: fail loudly on the host
      *error = 1;
      break;
    }
  } while ((unsigned int)(w >> 32) != epoch32);
  return (unsigned int)w;
}
template <typename T>
__device__ __noinline__ T ll_sum_impl(const unsigned char *half_base /*own LL area, this epoch's half*/, int nranks,
                                      unsigned long long slot_bytes, long long i, unsigned int epoch32,
                                      unsigned long long timeout_ns, int *error) {
  constexpr int WPV = LLW<T>::words;
  unsigned long long w[P2P_MAX_RANKS * WPV];
#pragma unroll
  for (int q = 0; q < P2P_MAX_RANKS; q++)
    if (q < nranks) {
      const unsigned long long *p = reinterpret_cast<const unsigned long long *>(half_base + (unsigned long long)q * slot_bytes) + WPV * i;
#pragma unroll
      for (int h = 0; h < WPV; h++) asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w[q * WPV + h]) : "l"(p + h) : "memory");
    }
  T acc = T(0);
#pragma unroll
  for (int q = 0; q < P2P_MAX_RANKS; q++)
    if (q < nranks) {
      const unsigned long long *p = reinterpret_cast<const unsigned long long *>(half_base + (unsigned long long)q * slot_bytes) + WPV * i;
      unsigned int d[WPV];
#pragma unroll
      for (int h = 0; h < WPV; h++)
        d[h] = (unsigned int)(w[q * WPV + h] >> 32) == epoch32 ? (unsigned int)w[q * WPV + h] : ll_poll_word(p + h, epoch32, timeout_ns, error);
      if (WPV == 2) acc += (T)__longlong_as_double((long long)(((unsigned long long)d[WPV - 1] << 32) | d[0]));
      else acc += (T)__uint_as_float(d[0]);
    }
  return acc;
}
template <typename T> __device__ __forceinline__ T ll_sum(const P2P &pp, unsigned long long epoch, long long i, unsigned int epoch32) {
  return ll_sum_impl<T>(pp.recv[pp.rank] + pp.ll_off + (epoch & 1ull) * pp.ll_half_bytes, pp.nranks, pp.ll_slot_bytes, i, epoch32,
                        pp.timeout_ns, pp.error);
}

// Generic pair for buffers that have no fused producer / consumer: push src to every rank, then sum in place.
template <typename T>
__global__ void __launch_bounds__(256) k_p2p_push(P2P pp, const T *__restrict__ src, long long n) {
  const unsigned long long epoch = p2p_next_epoch(pp);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const T v = src[i];
    for (int r = 0; r < pp.nranks; r++) p2p_slot<T>(pp, r, pp.rank, epoch)[i] = v;
  }
  p2p_signal(pp, epoch, gridDim.x);
}
template <typename T>
__global__ void __launch_bounds__(256) k_p2p_sum(P2P pp, T *__restrict__ dst, long long n) {
  const unsigned long long epoch = p2p_current_epoch(pp);
  p2p_wait(pp, epoch);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = p2p_sum<T>(pp, epoch, i);
}

} // namespace gb
