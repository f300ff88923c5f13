// Layer-ahead weight prefetch into L2 for latency-bound decode steps (tensor-parallel shards).
//
// At 8-way tensor parallelism one layer of Llama-3-70B is 27 MB of weights per rank, streamed by five kernels of 3-19 us
// each; measured, a step spends ~35 us per layer streaming and ~60 us in kernel boundaries, ramp-up and all-reduce waits,
// during which HBM idles.  B200's L2 is 126 MB: several layers of one rank's weights fit.  This
// kernel runs on a side stream with a handful of CTAs, next to the step's kernels, and walks the weights one layer ahead
// of the compute with `prefetch.global.L2`, so that the GEMVs find their operands in L2 and HBM works through the
// boundaries.  It never touches data the step depends on: correctness does not depend on it, only timing.
//
// Pacing: the prefetcher must stay a bounded distance ahead (running free it would finish all layers in a fraction of
// the step and evict its own lines).  It reads the step's progress from a word the step already produces: the tag of
// an LL all-reduce unit ({payload, tag} with tag = epoch * L + layer, csrc/linear_decode.cu) — when the tag of layer l
// appears, layer l's attention output phase is running, and the prefetcher may fetch layer l + lookahead.  A wall-clock
// watchdog bounds every wait, after which the kernel simply continues (or stops): a missing producer cannot hang it.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "petals_b200.h"

namespace {

struct L2PrefetchParams {
  const unsigned long long* ranges;   // [n_layers * per_layer] device addresses
  const long long* nbytes;            // [n_layers * per_layer]
  int n_layers, per_layer;
  const uint2* progress;              // LL unit to watch (.y = tag); nullptr = unpaced
  const unsigned long long* epoch;    // the step's epoch counter (already bumped when this kernel starts)
  unsigned int tag_mul;               // L of the tag formula
  int lookahead;                      // fetch layer l once progress >= l - lookahead
  unsigned long long wait_ns;         // watchdog per wait
};

__global__ void __launch_bounds__(256) l2_prefetch_kernel(L2PrefetchParams p) {
  __shared__ int go;
  const unsigned int base = p.progress ? static_cast<unsigned int>(*p.epoch) * p.tag_mul : 0u;
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  for (int l = 0; l < p.n_layers; ++l) {
    const int need = l - p.lookahead;  // progress value that releases layer l
    if (p.progress != nullptr && need >= 0) {
      if (threadIdx.x == 0) {
        const uint64_t t0 = pb::globaltimer_ns();
        int ok = 1;
        for (unsigned int spins = 0;; ++spins) {
          const unsigned int tag = *reinterpret_cast<const volatile unsigned int*>(&p.progress->y);
          const unsigned int rel = tag - base;  // tags of earlier epochs wrap to huge values
          if (rel < p.tag_mul && static_cast<int>(rel) >= need) break;
          __nanosleep(200);
          if ((spins & 0x3f) == 0 && pb::globaltimer_ns() - t0 > p.wait_ns) { ok = 0; break; }
        }
        go = ok;
      }
      __syncthreads();
      const int cont = go;
      __syncthreads();
      if (!cont) return;  // the step is not progressing the way we expect: stop prefetching, harmlessly
    }
    for (int r = 0; r < p.per_layer; ++r) {
      const char* ptr = reinterpret_cast<const char*>(p.ranges[l * p.per_layer + r]);
      const long long lines = (p.nbytes[l * p.per_layer + r] + 127) >> 7;
      for (long long i = tid; i < lines; i += nthreads) pb::prefetch_l2(ptr + (i << 7));
    }
  }
}

}  // namespace

// ranges / nbytes: device arrays of n_layers * per_layer entries. progress / epoch may be null together (unpaced run).
extern "C" int pb_l2_prefetch(const void* ranges, const void* nbytes, int n_layers, int per_layer, const void* progress, const void* epoch,
                              unsigned int tag_mul, int lookahead, int ctas, long long wait_us, void* stream) {
  if (n_layers <= 0 || per_layer <= 0 || ctas <= 0 || ctas > 148 || lookahead < 0) return PB_ERR_SHAPE;
  if ((progress == nullptr) != (epoch == nullptr)) return PB_ERR_SHAPE;
  if (progress != nullptr && tag_mul == 0) return PB_ERR_SHAPE;
  L2PrefetchParams p;
  p.ranges = static_cast<const unsigned long long*>(ranges);
  p.nbytes = static_cast<const long long*>(nbytes);
  p.n_layers = n_layers; p.per_layer = per_layer;
  p.progress = static_cast<const uint2*>(progress);
  p.epoch = static_cast<const unsigned long long*>(epoch);
  p.tag_mul = tag_mul; p.lookahead = lookahead;
  p.wait_ns = static_cast<unsigned long long>(wait_us > 0 ? wait_us : 2000) * 1000ull;
  l2_prefetch_kernel<<<ctas, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return pb_check_launch("l2_prefetch");
}
