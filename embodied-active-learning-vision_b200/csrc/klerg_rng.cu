// Workspace samples of Robot.get_samples on the device, bit-exact with the reference's host draw (sm_100a).
//
// The reference draws   samples = low + torch.rand(N, D) * (high - low)   (klerg.py:173,375; torch.distributions.Uniform)
// with torch's CPU generator: the 32-bit Mersenne Twister MT19937, one tempered output y per float32,
// u = (y & (2^24 - 1)) * 2^-24.  At 1e7 samples x 6 dimensions that is 0.3 s of one host core plus a 240 MB
// host-to-device copy per planner step - the floor of Robot.step() once the pair passes run on the GPU.  This
// kernel continues the SAME generator on the device: the host hands over the generator's 624-word state, the kernel
// produces the same stream, writes the samples of this rank's slice and returns the advanced state, which the host
// puts back into torch's generator (so the memory-buffer randperm that follows draws what it would have drawn).
//
// MT19937 regenerates its state 624 words at a time, and word i of the new block depends on words i, i+1 and
// i+397 of the old / partly new block: three waves of <= 227 independent words.  One CTA runs the recurrence
// (it is sequential from block to block); while its first 227 threads produce block b, the other threads temper,
// convert and store the outputs of block b-1.  ~100k blocks for 6e7 numbers, a few hundred cycles each.
#include <cstdint>

#include "klerg_common.cuh"

namespace klerg {

constexpr int MT_N = 624, MT_M = 397;
constexpr int MT_GEN = 256;                 // generator group: warps 0-7
constexpr int MT_OUT = 768;                 // output group: warps 8-31, one output of a block per thread
constexpr int MT_THREADS = MT_GEN + MT_OUT;

__device__ __forceinline__ uint32_t mt_twist(uint32_t u, uint32_t v) {
  return (((u & 0x80000000u) | (v & 0x7FFFFFFFu)) >> 1) ^ ((v & 1u) ? 0x9908B0DFu : 0u);
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
  y ^= y >> 11;
  y ^= (y << 7) & 0x9D2C5680u;
  y ^= (y << 15) & 0xEFC60000u;
  y ^= y >> 18;
  return y;
}

struct RngArgs {
  const uint32_t* state_in;  // [624]
  int left, next;            // the engine's counters (mt19937_engine: left_, next_)
  int64_t count;             // numbers to draw (= N * D)
  int D;
  float lo[KLERG_MAX_D], span[KLERG_MAX_D];
  int64_t first, last;       // numbers [first, last) are written (this rank's rows), out[i - first]
  float* out;
  uint32_t* state_out;       // [624 + 2]: state, left, next after the draw
};

// value k of the stream (k-th draw of this call) from a tempered output; lo / span of its dimension
__device__ __forceinline__ void mt_emit(const RngArgs& a, int64_t k, float lo, float span, uint32_t y) {
  if (k < a.first || k >= a.last) return;
  const float u = (float)(y & 0x00FFFFFFu) * 5.9604644775390625e-08f;  // exact: 24 bits times 2^-24
  a.out[k - a.first] = __fadd_rn(lo, __fmul_rn(u, span));  // low + rand * (high - low): multiply, then add
}

__global__ void __launch_bounds__(MT_THREADS) mt19937_uniform_kernel(const RngArgs a) {
  __shared__ uint32_t st[2][MT_N];
  __shared__ float s_lo[MT_N + KLERG_MAX_D], s_sp[MT_N + KLERG_MAX_D];  // lo / span by (offset + index) without a modulo
  const int tid = threadIdx.x;
  for (int i = tid; i < MT_N; i += MT_THREADS) st[0][i] = a.state_in[i];
  for (int i = tid; i < MT_N + KLERG_MAX_D; i += MT_THREADS) {
    s_lo[i] = a.lo[i % a.D];
    s_sp[i] = a.span[i % a.D];
  }
  __syncthreads();
  int cur = 0;            // buffer holding the block the outputs are currently read from
  int left = a.left, next = a.next;
  int64_t done = 0;
  // outputs still available in the block the host handed over
  {
    const int avail = left - 1;
    const int take = (int)min((int64_t)avail, a.count);
    for (int i = tid; i < take; i += MT_THREADS) mt_emit(a, i, s_lo[i], s_sp[i], mt_temper(st[0][next + i]));
    done = take;
    next += take;
    left -= take;
  }
  // whole blocks: regenerate into the other buffer while the previous block's outputs are stored
  int64_t prev_base = -1;  // stream index of the first output of the block in st[cur] still to be emitted
  int prev_take = 0;
  int dm = 0;              // prev_base % D
  const int dstep = MT_N % a.D;
  while (done < a.count || prev_take > 0) {
    const bool gen = done < a.count;
    const uint32_t* old = st[cur];
    uint32_t* nw = st[cur ^ 1];
    if (tid < MT_GEN) {
      // generator group (warps 0-7): the three waves, meeting on a named barrier of their own
      if (gen) {
        if (tid < MT_N - MT_M) nw[tid] = old[tid + MT_M] ^ mt_twist(old[tid], old[tid + 1]);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid < MT_N - MT_M) {
          const int i = (MT_N - MT_M) + tid;
          nw[i] = nw[tid] ^ mt_twist(old[i], old[i + 1]);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid < MT_N - 1 - 2 * (MT_N - MT_M)) {
          const int i = 2 * (MT_N - MT_M) + tid;
          nw[i] = nw[i - (MT_N - MT_M)] ^ mt_twist(old[i], old[i + 1]);
        }
        if (tid == 255) nw[MT_N - 1] = nw[MT_M - 1] ^ mt_twist(old[MT_N - 1], nw[0]);  // words 396 and 0: waves 2 and 1
      }
    } else {
      // output group (warps 8-31): temper, convert and store the pending block, one output per thread
      const int i = tid - MT_GEN;
      if (i < prev_take) mt_emit(a, prev_base + i, s_lo[dm + i], s_sp[dm + i], mt_temper(old[i]));
    }
    prev_take = 0;
    if (gen) {
      // the regenerating call reads word 0 of the new block: left = 624, next = 0 before it
      const int take = (int)min((int64_t)MT_N, a.count - done);
      if (prev_base >= 0) dm += dstep;
      else dm = (int)(done % a.D);
      if (dm >= a.D) dm -= a.D;
      prev_base = done;
      prev_take = take;
      done += take;
      next = take;
      left = MT_N + 1 - take;
      cur ^= 1;
    }
    __syncthreads();
  }
  for (int i = tid; i < MT_N; i += MT_THREADS) a.state_out[i] = st[cur][i];
  if (tid == 0) {
    a.state_out[MT_N] = (uint32_t)left;
    a.state_out[MT_N + 1] = (uint32_t)next;
  }
}

}  // namespace klerg

using namespace klerg;

extern "C" int klerg_mt19937_uniform(const uint32_t* state624, int32_t left, int32_t next, int64_t n_rows, int32_t D,
                                     const float* low, const float* high_minus_low, int64_t row_lo, int64_t row_hi,
                                     float* out, uint32_t* state_out, void* stream) {
  if (!state624 || !state_out || !low || !high_minus_low) { set_error("mt19937_uniform: null argument"); return -1; }
  if (D < 1 || D > KLERG_MAX_D || n_rows < 0) { set_error("mt19937_uniform: bad sizes"); return -1; }
  if (left < 1 || left > MT_N + 1 || next < 0 || next > MT_N || (left > 1 && next + left - 1 > MT_N)) {
    set_error("mt19937_uniform: inconsistent generator counters (left=%d next=%d)", left, next);
    return -1;
  }
  if (row_lo < 0 || row_hi < row_lo || row_hi > n_rows) { set_error("mt19937_uniform: bad row range"); return -1; }
  if (row_hi > row_lo && !out) { set_error("mt19937_uniform: null output"); return -1; }
  RngArgs a{};
  a.state_in = state624; a.left = left; a.next = next; a.count = n_rows * D; a.D = D;
  for (int d = 0; d < D; ++d) { a.lo[d] = low[d]; a.span[d] = high_minus_low[d]; }
  a.first = row_lo * D; a.last = row_hi * D; a.out = out; a.state_out = state_out;
  mt19937_uniform_kernel<<<1, MT_THREADS, 0, (cudaStream_t)stream>>>(a);
  return check_launch("mt19937_uniform_kernel");
}
