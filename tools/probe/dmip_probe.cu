// dmip_probe.cu — tcgen05 design probes and building-block self-tests.  NOT part of the product: builds into
// tools/probe/libdmip_probe.so (tools/probe/build.sh), loaded only by tests/ and tools/.  MMA issue rate for single-CTA
// and CTA-pair instructions, A from shared or tensor memory, with an optional concurrent bulk-TMA weight stream; the
// cost of the synchronisation primitives; small GEMMs through the library's own descriptor / layout helpers
// (K-major and MN-major operands, bf16 hi/lo split).  The numbers these print decide the kernels' tiling.
#include <stdarg.h>
#include <stdio.h>

#include "../../diffusion-modelling-for-inverse-problems_b200/csrc/dmip_common.h"
#include "../../diffusion-modelling-for-inverse-problems_b200/csrc/dmip_ptx.cuh"

namespace dmip {
// the product's error / launch-count plumbing lives in dmip_api.cu; the probe library carries its own
static thread_local char g_probe_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_probe_err, sizeof(g_probe_err), fmt, ap);
  va_end(ap);
}
void count_launch(int) {}
void reset_launch_count() {}
}  // namespace dmip

namespace dmip {

namespace {

constexpr int kKb = 16384;        // one 128-row x 64-k bf16 K-block
constexpr int kRing = 2;

struct BenchParams {
  int cg;            // CTAs per MMA (1 or 2)
  int a_tmem;        // 0: A from shared memory, 1: A from tensor memory
  int alt_acc;       // 1: alternate between two accumulators (no D dependence between consecutive chunks)
  int n;             // N of one instruction (whole pair)
  int k;             // K per accumulation chain (multiple of 64, <= 256)
  int iters;         // chains
  int stream_bytes;  // bytes the producer warp copies global -> shared per K-block (0: no stream)
  int rnd;           // 1: operands hold pseudo-random bf16 data instead of zeros (switching power)
  int commit;        // 1: one tcgen05.commit per K-block (4 MMAs), as the sampler does
  int readers;       // 1: warps 4-7 read the accumulators with tcgen05.ld in a loop (epilogue stand-in); 2: ld + st
  int reader_iters;
  int kbwait;        // n: n mbarrier waits on an already-completed phase (+ tcgen05.fence::after) before every K-block
  int math;          // warps 8-23: 1 = MUFU.TANH loop, 2 = FFMA loop, 3 = both (row-warp stand-in without TMEM traffic)
  int math_iters;
  float* sink;
  const uint8_t* gsrc;
  long long* cycles; // [grid][4]: issue, complete, producer cycles, unused
};

// kCg is a template parameter: a kernel that CONTAINS cta_group::2 instructions can only be launched with an even
// cluster dimension (launch fails with "cluster misconfiguration" otherwise), so the single-CTA probe is a separate kernel.
template <int kCg>
__global__ void __launch_bounds__(768, 1) k_mma_bench2(const __grid_constant__ BenchParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                          // 4 K-blocks x 16 KB
  uint8_t* sW = smem + 4 * kKb;                // 4 K-blocks x (rows_per_cta x 128 B), rows_per_cta <= 256
  uint8_t* sR = smem + 4 * kKb + 8 * kKb;      // stream ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(sR + kRing * kKb);
  uint64_t* done = bars;                       // MMA completion
  uint64_t* rfull = bars + 1;                  // [kRing]
  uint64_t* dummy = bars + 1 + kRing;          // per-K-block commits land here (nobody waits)
  uint64_t* ready = bars + 2 + kRing;          // phase 0 completed during setup
  uint32_t* holder = reinterpret_cast<uint32_t*>(bars + 3 + kRing);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t crank = kCg == 2 ? cluster_ctarank() : 0u;
  if (threadIdx.x == 0) {
    mbar_init(done, 1);
    for (int i = 0; i < kRing; ++i) mbar_init(&rfull[i], 1);
    mbar_init(dummy, 1);
    mbar_init(ready, 1);
    fence_barrier_init();
    mbar_arrive(ready);
  }
  for (int i = threadIdx.x; i < (12 + kRing) * kKb / 16; i += 768) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (P.rnd) {   // bf16 values in (-2, 2) with random mantissas: 0x3f80..0x3fff / 0xbf80..0xbfff
      uint32_t h = (i * 2654435761u) ^ (blockIdx.x * 40503u);
      auto nxt = [&]() { h = h * 1664525u + 1013904223u; return ((h >> 9) & 0x807F807Fu) | 0x3F803F80u; };
      v = make_uint4(nxt(), nxt(), nxt(), nxt());
    }
    reinterpret_cast<uint4*>(smem)[i] = v;
  }
  fence_proxy_async_smem();
  if (kCg == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 0) {
    if (kCg == 2) tmem_alloc_pair<512>(holder); else tmem_alloc<512>(holder);
  }
  tc_fence_before();
  if (kCg == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *holder, 0);
  if (P.rnd && warp < 4) {   // A operand image in tensor memory columns 0..127
    const uint32_t lt = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    for (int c0 = 0; c0 < 128; c0 += 16) {
      uint32_t pk[16];
      uint32_t h = (threadIdx.x * 2654435761u) ^ (c0 * 97u);
#pragma unroll
      for (int j = 0; j < 16; ++j) { h = h * 1664525u + 1013904223u; pk[j] = ((h >> 9) & 0x807F807Fu) | 0x3F803F80u; }
      tmem_st16(lt + c0, pk);
    }
    tc_wait_st();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();
  const int rows_b = P.n / kCg;               // B rows held by this CTA
  long long t_issue = 0, t_done = 0, t_prod = 0;
  if (warp == 0 && crank == 0) {
    const uint32_t idesc = umma_idesc_bf16(128u * kCg, static_cast<uint32_t>(P.n));
    const uint64_t desc_hi = umma_smem_desc_sw128(0) & 0xFFFFFFFF00000000ull;
    const uint32_t a_lo0 = (smem_u32(sA) & 0x3FFFFu) >> 4, b_lo0 = (smem_u32(sW) & 0x3FFFFu) >> 4;
    const uint32_t wkb = static_cast<uint32_t>(rows_b) * 128u;
    const long long t0 = clock64();
    for (int it = 0; it < P.iters; ++it) {
      const uint32_t d_tmem = tmem_base + 256 + ((P.alt_acc && (it & 1)) ? static_cast<uint32_t>(P.n <= 128 ? 128 : 0) : 0u);
      for (int kb = 0; kb < P.k / 64; ++kb) {
        for (int w = 0; w < P.kbwait; ++w) {
          mbar_wait(ready, 0, 0x905);
          tc_fence_after();
        }
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t bdesc = desc_hi | (b_lo0 + ((kb * wkb) >> 4) + kk * 2);
            const uint64_t adesc = desc_hi | (a_lo0 + ((kb * kKb) >> 4) + kk * 2);
            const uint32_t a_tm = tmem_base + kb * 32 + kk * 8;
            if (kCg == 2) {
              if (P.a_tmem) umma_ts_pair(d_tmem, a_tm, bdesc, idesc, 1u);
              else umma_ss_pair(d_tmem, adesc, bdesc, idesc, 1u);
            } else {
              if (P.a_tmem) umma_ts(d_tmem, a_tm, bdesc, idesc, 1u);
              else umma_ss(d_tmem, adesc, bdesc, idesc, 1u);
            }
          }
          if (P.commit) {
            if (kCg == 2) tc_commit_pair(dummy, 0x1); else tc_commit(dummy);
          }
        }
        __syncwarp();
      }
    }
    if (elect_one()) {
      if (kCg == 2) tc_commit_pair(done, 0x3); else tc_commit(done);
    }
    __syncwarp();
    t_issue = clock64() - t0;
    mbar_wait(done, 0, 0x901);
    t_done = clock64() - t0;
  } else if (warp == 0) {
    mbar_wait(done, 0, 0x902);                 // peer CTA: its tensor core works for the leader's instructions
  } else if (warp == 1 && P.stream_bytes > 0) {
    // weight-stream stand-in: stream_bytes per K-block into a 2-slot ring, next copy issued when the previous one
    // into that slot has landed (self-paced: as fast as L2 / shared memory allow, capped at the MMA count)
    const long long t0 = clock64();
    const int n_copies = P.iters * (P.k / 64);
    uint32_t ph = 0;
    int s = 0;
    const size_t span = 8u << 20;
    size_t off = (static_cast<size_t>(blockIdx.x) * 65536) % span;
    for (int i = 0; i < n_copies; ++i) {
      if (i >= kRing) mbar_wait(&rfull[s], ph, 0x903);
      if (elect_one()) {
        mbar_arrive_expect_tx(&rfull[s], static_cast<uint32_t>(P.stream_bytes));
        bulk_g2s(sR + s * kKb, P.gsrc + off, static_cast<uint32_t>(P.stream_bytes), &rfull[s]);
      }
      __syncwarp();
      off = (off + P.stream_bytes) % span;
      if (++s == kRing) { s = 0; if (i >= kRing) ph ^= 1u; }
    }
    // drain
    for (int i = 0; i < kRing && i < n_copies; ++i) {
      mbar_wait(&rfull[s], ph, 0x904);
      if (++s == kRing) { s = 0; ph ^= 1u; }
    }
    t_prod = clock64() - t0;
  } else if (warp >= 8) {
    if (P.math > 0) {
      float a0 = threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f;
      for (int it = 0; it < P.math_iters; ++it) {
        if (P.math & 1) {
          asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a0));
          asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a1));
          asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a2));
          asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a3));
        }
        if (P.math & 2) {
#pragma unroll
          for (int r = 0; r < 6; ++r) {
            a0 = fmaf(a0, 1.0001f, 0.5f); a1 = fmaf(a1, 1.0001f, 0.5f); a2 = fmaf(a2, 1.0001f, 0.5f); a3 = fmaf(a3, 1.0001f, 0.5f);
          }
        }
      }
      if (a0 + a1 + a2 + a3 == 123.456f) P.sink[threadIdx.x] = a0;
    }
  } else if (warp >= 4 && P.readers > 0) {
    // epilogue stand-in: lane quarter (warp % 4) reads 2 x 32 accumulator columns per iteration, optionally writes 16
    const uint32_t lt = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    for (int it = 0; it < P.reader_iters; ++it) {
      uint32_t v[32];
      tmem_ld32(lt + 256 + (it & 1) * 64, v);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= v[j];
      if (P.readers > 1) {
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = v[j] ^ v[j + 16];
        tmem_st16(lt + 128 + (it & 3) * 16, pk);
        tc_wait_st();
      }
    }
    if (acc == 0x12345678u) P.cycles[blockIdx.x * 4 + 3] = acc;
  }
  if (lane == 0 && warp == 0 && crank == 0) {
    P.cycles[blockIdx.x * 4] = t_issue;
    P.cycles[blockIdx.x * 4 + 1] = t_done;
  }
  if (lane == 0 && warp == 1) P.cycles[blockIdx.x * 4 + 2] = t_prod;
  tc_fence_before();
  if (kCg == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    if (kCg == 2) tmem_dealloc_pair<512>(tmem_base); else tmem_dealloc<512>(tmem_base);
  }
}


// ------------------------------------------------------------------------------------------------ primitive costs
// One warp, `iters` iterations of a small sequence of the synchronisation primitives the sampler's producer / issuer
// loops are made of; cycles per iteration for each sequence.  out[m] for m = 0..7.
struct PrimParams {
  int iters;
  int a, b, c;       // read inside the loops (constant-bank loads the compiler cannot hoist over "memory" clobbers)
  long long* out;
  unsigned long long* sink;
};

__global__ void __launch_bounds__(64, 1) k_prim_bench(const __grid_constant__ PrimParams P) {
  __shared__ uint64_t bar_done[2];   // phase 0 completed before the loop
  __shared__ uint64_t bar_self;      // count 1: every arrive completes a phase
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar_done[0], 1);
    mbar_init(&bar_done[1], 1);
    mbar_init(&bar_self, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) { mbar_arrive(&bar_done[0]); mbar_arrive(&bar_done[1]); }
  if (warp == 0) tmem_alloc<32>(&holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp != 0) return;
  unsigned long long acc = 0;
  for (int m = 0; m < 10; ++m) {
    __syncwarp();
    const long long t0 = clock64();
    uint32_t ph = 0;
    for (int i = 0; i < P.iters; ++i) {
      switch (m) {
        case 0: mbar_wait(&bar_done[0], 0, 1); break;                              // try_wait on a long-completed phase
        case 1: if (elect_one()) mbar_arrive(&bar_self); __syncwarp(); break;      // elect + arrive
        case 2: if (elect_one()) mbar_arrive(&bar_self); __syncwarp();             // arrive then wait for that phase
                mbar_wait(&bar_self, ph, 2); ph ^= 1u; break;
        case 3: acc += clock64(); break;                                           // CS2R
        case 4: if (elect_one()) tc_commit(&bar_self); __syncwarp(); break;        // tcgen05.commit (nothing pending)
        case 5: if (elect_one()) tc_commit(&bar_self); __syncwarp();               // commit then wait for its arrival
                mbar_wait(&bar_self, ph, 3); ph ^= 1u; break;
        case 6: tc_fence_after(); break;
        case 7: asm volatile("" ::: "memory"); acc += P.a + P.b + P.c; break;      // constant-bank reloads
        case 8: mbar_wait(&bar_done[0], 0, 1); tc_fence_after();                   // the issuer's per-K-block skeleton
                if (elect_one()) tc_commit(&bar_self); __syncwarp(); break;
        case 9: mbar_wait(&bar_done[1], 0, 1);                                     // the producer's per-stage skeleton
                if (elect_one()) mbar_arrive_expect_tx(&bar_self, 0); __syncwarp(); break;
      }
    }
    const long long t1 = clock64();
    if (lane == 0) P.out[m] = t1 - t0;
    // leave bar_self in a known phase: it completes a phase per arrive; re-sync parity for the next sequence
    __syncwarp();
    if (m == 1 || m == 4 || m == 8 || m == 9) {
      // unknown parity: re-initialise
      if (lane == 0) { mbar_init(&bar_self, 1); fence_barrier_init(); }
      __syncwarp();
    }
  }
  if (acc == 0x1234567ull) P.sink[0] = acc;
  tc_fence_before();
  __syncwarp();
  tmem_dealloc<32>(holder);
}


// ------------------------------------------------------------------------------------------------ wake-up latency
// Two warps ping-pong over two mbarriers: cycles per round trip (= 2 x arrive -> the other warp notices).
// mode 0: mbarrier.try_wait (hardware suspend), 1: mbarrier.test_wait spin, 2: try_wait with a 0 ns suspend hint.
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void wait_mode(int mode, uint64_t* bar, uint32_t parity) {
  if (mode == 0) { while (!mbar_try_wait(bar, parity)) {} }
  else if (mode == 1) { while (!mbar_test_wait(bar, parity)) {} }
  else { while (!mbar_try_wait_hint(bar, parity, 0)) {} }
}

__global__ void __launch_bounds__(64, 1) k_pingpong(int iters, long long* out) {
  __shared__ uint64_t ping, pong;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int mode = 0; mode < 3; ++mode) {
    if (threadIdx.x == 0) {
      mbar_init(&ping, 1);
      mbar_init(&pong, 1);
      fence_barrier_init();
    }
    __syncthreads();
    const long long t0 = clock64();
    uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) {
      if (warp == 0) {
        if (elect_one()) mbar_arrive(&ping);
        __syncwarp();
        wait_mode(mode, &pong, ph);
      } else {
        wait_mode(mode, &ping, ph);
        if (elect_one()) mbar_arrive(&pong);
        __syncwarp();
      }
      ph ^= 1u;
    }
    const long long t1 = clock64();
    if (warp == 0 && lane == 0) out[20 + mode] = t1 - t0;
    __syncthreads();
  }
}


// Wake-up latency of a sleeping waiter as a function of how long it has been asleep: warp 1 arrives `delay` cycles after
// warp 0 started waiting; out[24 + 4*k + mode] = cycles between the arrive and warp 0 noticing, for delays 2^k * 250.
__global__ void __launch_bounds__(64, 1) k_wakeup(long long* out) {
  __shared__ uint64_t bar, go;
  __shared__ long long t_arrive;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int mode = 0; mode < 3; ++mode) {
    for (int k = 0; k < 6; ++k) {
      const long long delay = 250LL << k;
      if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_init(&go, 1);
        fence_barrier_init();
      }
      __syncthreads();
      if (warp == 0) {
        if (elect_one()) mbar_arrive(&go);
        __syncwarp();
        wait_mode(mode, &bar, 0);
        const long long t1 = clock64();
        __syncwarp();
        if (lane == 0) out[24 + 4 * k + mode] = t1 - t_arrive;
      } else {
        wait_mode(0, &go, 0);
        const long long t0 = clock64();
        while (clock64() - t0 < delay) {}
        if (lane == 0) {
          t_arrive = clock64();
          mbar_arrive(&bar);
        }
        __syncwarp();
      }
      __syncthreads();
    }
  }
}


// ------------------------------------------------------------------------------------------------ XU pipe rates
// 16 warps per SM each issue `iters` x 8 independent instructions of one kind; out[40 + kind] = cycles.
// kind 0 MUFU.TANH, 1 MUFU.EX2, 2 MUFU.RCP, 3 F2FP.BF16 pack, 4 FFMA, 5 MUFU.TANH + F2FP interleaved (epilogue mix)
template <int kind>
__device__ __forceinline__ long long xu_loop(int iters, float (&a)[8], uint32_t& pk) {
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (kind == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[j]));
      else if (kind == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[j]));
      else if (kind == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[j]));
      else if (kind == 3) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[j]), "f"(a[(j + 1) & 7])); pk ^= r; }
      else if (kind == 4) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[j]));
      else if (kind == 5) {
        asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[j]));
        if (j & 1) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[j]), "f"(a[j - 1])); pk ^= r; }
      } else if (kind == 6) {   // 32 x 32 -> 64 multiply (IMAD.WIDE.U32): the Philox round primitive
        unsigned long long w;
        uint32_t x = __float_as_uint(a[j]);
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(x), "r"(0xD2511F53u));
        a[j] = __uint_as_float(static_cast<uint32_t>(w) ^ static_cast<uint32_t>(w >> 32));
      } else if (kind == 7) {   // mul.hi.u32 alone
        uint32_t x = __float_as_uint(a[j]), h;
        asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(h) : "r"(x), "r"(0xD2511F53u));
        a[j] = __uint_as_float(h);
      } else if (kind == 8) {   // mul.lo.u32 alone
        uint32_t x = __float_as_uint(a[j]), h;
        asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(h) : "r"(x), "r"(0xD2511F53u));
        a[j] = __uint_as_float(h);
      } else if (kind == 9) {   // u32 -> f32 conversion
        uint32_t x = __float_as_uint(a[j]);
        asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(a[j]) : "r"(x));
      } else if (kind == 10) {  // packed half tanh: two activations per MUFU instruction
        uint32_t x = __float_as_uint(a[j]);
        asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(x));
        a[j] = __uint_as_float(x);
      } else if (kind == 11) {  // packed bf16 tanh
        uint32_t x = __float_as_uint(a[j]);
        asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(x));
        a[j] = __uint_as_float(x);
      } else {                  // f32 pair -> f16x2 pack
        uint32_t r;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[j]), "f"(a[(j + 1) & 7]));
        pk ^= r;
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  return t1 - t0;
}

__global__ void __launch_bounds__(512, 1) k_xu_rate(int iters, long long* out, float* sink) {
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = threadIdx.x * 1e-3f + j * 0.01f;
  uint32_t pk = 0;
  long long t[13];
  t[10] = xu_loop<10>(iters, a, pk);
  t[11] = xu_loop<11>(iters, a, pk);
  t[12] = xu_loop<12>(iters, a, pk);
  t[6] = xu_loop<6>(iters, a, pk);
  t[7] = xu_loop<7>(iters, a, pk);
  t[8] = xu_loop<8>(iters, a, pk);
  t[9] = xu_loop<9>(iters, a, pk);
  t[0] = xu_loop<0>(iters, a, pk);
  t[1] = xu_loop<1>(iters, a, pk);
  t[2] = xu_loop<2>(iters, a, pk);
  t[3] = xu_loop<3>(iters, a, pk);
  t[4] = xu_loop<4>(iters, a, pk);
  t[5] = xu_loop<5>(iters, a, pk);
  if (threadIdx.x == 0)
    for (int k = 0; k < 13; ++k) out[40 + k] = t[k];
  if (a[0] + a[1] + a[2] + a[3] + a[4] + a[5] + a[6] + a[7] + __uint_as_float(pk) == 123.f) sink[0] = a[0];
}

// legacy warp-level tensor path: issue rate of mma.sync with 16 warps x 8 independent accumulator tiles
template <int kKind>
__device__ long long mmasync_loop(int iters, float (&c)[8][4], uint32_t a0, uint32_t b0) {
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (kKind == 0) {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                     : "r"(a0), "r"(a0 + j), "r"(a0 ^ 5u), "r"(a0 + 3u), "r"(b0), "r"(b0 + j));
      } else {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                     : "r"(a0), "r"(a0 + j), "r"(a0 ^ 5u), "r"(a0 + 3u), "r"(b0), "r"(b0 + j));
      }
    }
  }
  __syncthreads();
  return clock64() - t0;
}

__global__ void __launch_bounds__(512, 1) k_mmasync_rate(int iters, long long* out, float* sink) {
  float c[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) c[j][e] = 0.f;
  const uint32_t a0 = 0x3c003c00u + threadIdx.x, b0 = 0x3c003c00u ^ threadIdx.x;
  const long long t0 = mmasync_loop<0>(iters, c, a0 & 0x3fffe000u, b0 & 0x3fffe000u);
  const long long t1 = mmasync_loop<1>(iters, c, a0, b0);
  if (threadIdx.x == 0) { out[0] = t0; out[1] = t1; }
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) acc += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  if (acc == 123.f) sink[0] = acc;
}


// ---- constants of the sampler's tile geometry the self-test kernels reuse
constexpr int kStageBytes = 16384;
constexpr uint32_t kTmemH = 0;
constexpr uint32_t kTmemAcc = 256;

// ------------------------------------------------------------------------------------------------ self-test kernel
// D(128 x n) = A(128 x k) W(n x k)^T through the same descriptors / layouts / TMEM access as the sampler.
__global__ void __launch_bounds__(128, 1) k_debug_umma(int mode, const float* __restrict__ a, const float* __restrict__ w,
                                                        float* __restrict__ d, int n, int k) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                       // up to 4 K-blocks
  uint8_t* sW = smem + 4 * kStageBytes;     // up to 4 K-blocks
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8 * kStageBytes);
  uint32_t* holder = reinterpret_cast<uint32_t*>(smem + 8 * kStageBytes + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = threadIdx.x;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(holder);
  for (int i = threadIdx.x; i < 8 * kStageBytes / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  // operand images
  for (int kk = 0; kk < k; ++kk) {
    const unsigned short ha = static_cast<unsigned short>(pack_bf16x2(a[row * k + kk], 0.f) & 0xFFFFu);
    *reinterpret_cast<unsigned short*>(sA + sw128_offset(row, kk, kStageBytes)) = ha;
    if (row < n) {
      const unsigned short hw = static_cast<unsigned short>(pack_bf16x2(w[row * k + kk], 0.f) & 0xFFFFu);
      *reinterpret_cast<unsigned short*>(sW + sw128_offset(row, kk, kStageBytes)) = hw;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *holder;
  const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  if (mode == 1) {
    // A into tensor memory: lane = row, 32-bit column c = elements (2c, 2c+1)
    for (int c0 = 0; c0 < k / 2; c0 += 16) {
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(a[row * k + 2 * (c0 + j)], a[row * k + 2 * (c0 + j) + 1]);
      tmem_st16(lane_taddr + kTmemH + c0, pk);
    }
    tc_wait_st();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, static_cast<uint32_t>(n));
    const uint32_t d_tmem = tmem_base + kTmemAcc;
    for (int kb = 0; kb < k / 64; ++kb)
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t bdesc = umma_smem_desc_sw128(smem_u32(sW) + kb * kStageBytes + kk * 32);
        const uint32_t acc = (kb | kk) != 0 ? 1u : 0u;
        if (mode == 0)
          umma_ss(d_tmem, umma_smem_desc_sw128(smem_u32(sA) + kb * kStageBytes + kk * 32), bdesc, idesc, acc);
        else
          umma_ts(d_tmem, tmem_base + kTmemH + kb * 32 + kk * 8, bdesc, idesc, acc);
      }
    tc_commit(bar);
  }
  mbar_wait(bar, 0, 0x900);
  tc_fence_after();
  for (int pc = 0; pc < n / 8; ++pc) {
    uint32_t v[8];
    tmem_ld8(lane_taddr + kTmemAcc + pc * 8, v);
    tc_wait_ld();
#pragma unroll
    for (int e = 0; e < 8; ++e) d[row * n + pc * 8 + e] = __uint_as_float(v[e]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
  (void)lane;
}

// ------------------------------------------------------------------------------------------------ MMA micro-benchmark
// Issues `iters` x (k/16) back-to-back tcgen05.mma (M=128, N=n, K=16) from one elected lane and reports the
// cycles from first issue to completion.  mode 0: A in shared memory, 1: A in tensor memory.  Operands are whatever
// is in shared / tensor memory (zeros): only the timing matters.
__global__ void __launch_bounds__(128, 1) k_debug_mma_bench(int mode, int n, int k, int iters, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sW = smem + 4 * kStageBytes;   // up to 256 rows x 4 K-blocks
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 12 * kStageBytes);
  uint32_t* holder = reinterpret_cast<uint32_t*>(smem + 12 * kStageBytes + 8);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(holder);
  for (int i = threadIdx.x; i < 12 * kStageBytes / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *holder, 0);
  if (warp == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, static_cast<uint32_t>(n));
    const uint64_t desc_hi = umma_smem_desc_sw128(0) & 0xFFFFFFFF00000000ull;
    const uint32_t a_lo0 = (smem_u32(sA) & 0x3FFFFu) >> 4, b_lo0 = (smem_u32(sW) & 0x3FFFFu) >> 4;
    const uint32_t wkb = static_cast<uint32_t>(n) * 128u;   // bytes per K-block of the B image
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int kb = 0; kb < k / 64; ++kb) {
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t bdesc = desc_hi | (b_lo0 + ((kb * wkb) >> 4) + kk * 2);
            const uint32_t d_tmem = tmem_base + kTmemAcc + (it & 1) * 0;   // same accumulator: worst-case dependence
            if (mode == 0) umma_ss(d_tmem, desc_hi | (a_lo0 + ((kb * kStageBytes) >> 4) + kk * 2), bdesc, idesc, 1u);
            else umma_ts(d_tmem, tmem_base + kTmemH + kb * 32 + kk * 8, bdesc, idesc, 1u);
          }
        }
        __syncwarp();
      }
    }
    if (elect_one()) tc_commit(bar);
    __syncwarp();
    const long long t1 = clock64();
    mbar_wait(bar, 0, 0x901);
    const long long t2 = clock64();
    if (threadIdx.x == 0) {
      cycles[blockIdx.x * 2] = t1 - t0;
      cycles[blockIdx.x * 2 + 1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}



// ------------------------------------------------------------------------------------------------ general GEMM self-test
// D(128 x n) = A(128 x k) B(n x k)^T with either operand K-major or MN-major in shared memory, optionally as the bf16x3
// split product  A_hi B_hi + A_hi B_lo + A_lo B_hi  (what the fused loss kernels run), through umma_smem_desc /
// mn128_offset / sw128_offset.  The image strides and the descriptor fields are passed separately so the host can
// probe the meaning of LBO / SBO for MN-major operands.
struct Umma2Params {
  int a_mn, b_mn, n, k, split, iters;
  uint32_t a_img_lbo, a_img_sbo, b_img_lbo, b_img_sbo;   // MN-major image strides (bytes)
  uint32_t a_fld_lbo, a_fld_sbo, b_fld_lbo, b_fld_sbo;   // descriptor fields (bytes)
  const float* a;
  const float* b;
  float* d;
  long long* cycles;
};

constexpr int kU2A = 32768, kU2B = 65536;   // bytes per image: A hi | A lo | B hi | B lo

__global__ void __launch_bounds__(128, 1) k_umma2(const __grid_constant__ Umma2Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA[2] = {smem, smem + kU2A};
  uint8_t* sB[2] = {smem + 2 * kU2A, smem + 2 * kU2A + kU2B};
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * kU2A + 2 * kU2B);
  uint32_t* holder = reinterpret_cast<uint32_t*>(smem + 2 * kU2A + 2 * kU2B + 8);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(holder);
  for (int i = threadIdx.x; i < (2 * kU2A + 2 * kU2B) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  const int n = P.n, k = P.k;
  auto put = [&](uint8_t* const (&img)[2], uint32_t off, float v) {
    const float hi = bf16_round(v);
    *reinterpret_cast<unsigned short*>(img[0] + off) = static_cast<unsigned short>(pack_bf16x2(hi, 0.f) & 0xFFFFu);
    *reinterpret_cast<unsigned short*>(img[1] + off) = static_cast<unsigned short>(pack_bf16x2(v - hi, 0.f) & 0xFFFFu);
  };
  for (int kk = 0; kk < k; ++kk) {
    const int m = threadIdx.x;
    put(sA, P.a_mn ? mn128_offset(m, kk, P.a_img_lbo, P.a_img_sbo) : sw128_offset(m, kk, 128 * 128), P.a[m * k + kk]);
    for (int r = threadIdx.x; r < n; r += 128)
      put(sB, P.b_mn ? mn128_offset(r, kk, P.b_img_lbo, P.b_img_sbo) : sw128_offset(r, kk, n * 128), P.b[r * k + kk]);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *holder;
  const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16_major(128, static_cast<uint32_t>(n), P.a_mn, P.b_mn);
    t0 = clock64();
    for (int it = 0; it < P.iters; ++it) {
      for (int j = 0; j < k / 16; ++j) {
        // start address of K-slice j of each image
        const uint32_t a_off = P.a_mn ? j * 2 * P.a_img_sbo : (j >> 2) * (128 * 128) + (j & 3) * 32;
        const uint32_t b_off = P.b_mn ? j * 2 * P.b_img_sbo : (j >> 2) * (n * 128) + (j & 3) * 32;
        const uint32_t al = P.a_mn ? P.a_fld_lbo : 16, as = P.a_mn ? P.a_fld_sbo : 1024;
        const uint32_t bl = P.b_mn ? P.b_fld_lbo : 16, bs = P.b_mn ? P.b_fld_sbo : 1024;
        const uint64_t ah = umma_smem_desc(smem_u32(sA[0]) + a_off, al, as), alo = umma_smem_desc(smem_u32(sA[1]) + a_off, al, as);
        const uint64_t bh = umma_smem_desc(smem_u32(sB[0]) + b_off, bl, bs), blo = umma_smem_desc(smem_u32(sB[1]) + b_off, bl, bs);
        umma_ss(tmem_base, ah, bh, idesc, (it | j) != 0 ? 1u : 0u);
        if (P.split) {
          umma_ss(tmem_base, ah, blo, idesc, 1u);
          umma_ss(tmem_base, alo, bh, idesc, 1u);
        }
      }
    }
    tc_commit(bar);
    t1 = clock64();
  }
  mbar_wait(bar, 0, 0x902);
  tc_fence_after();
  if (threadIdx.x == 0 && P.cycles) {
    P.cycles[0] = t1 - t0;
    P.cycles[1] = clock64() - t0;
  }
  for (int pc = 0; pc < n / 8; ++pc) {
    uint32_t v[8];
    tmem_ld8(lane_taddr + pc * 8, v);
    tc_wait_ld();
#pragma unroll
    for (int e = 0; e < 8; ++e) P.d[threadIdx.x * n + pc * 8 + e] = __uint_as_float(v[e]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

int launch_debug_mma_bench(int mode, int n, int k, int iters, int grid, long long* cycles, cudaStream_t s) {
  DMIP_REQUIRE(k % 64 == 0 && k >= 64 && k <= 256 && n % 16 == 0 && n >= 16 && n <= 256 && grid >= 1, "bad bench shape");
  const int smem = 12 * kStageBytes + 64 + 1024;
  DMIP_CHECK_CUDA(cudaFuncSetAttribute(k_debug_mma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k_debug_mma_bench<<<grid, 128, smem, s>>>(mode, n, k, iters, cycles);
  DMIP_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DMIP_OK;
}

int launch_debug_umma(int mode, const float* a, const float* w, float* d, int n, int k, cudaStream_t s) {
  DMIP_REQUIRE(k % 64 == 0 && k >= 64 && k <= 256, "debug_umma: k must be 64, 128, 192 or 256");
  DMIP_REQUIRE(n % 16 == 0 && n >= 16 && n <= 128, "debug_umma: n must be a multiple of 16 in [16,128]");
  const int smem = 8 * kStageBytes + 64 + 1024;
  DMIP_CHECK_CUDA(cudaFuncSetAttribute(k_debug_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k_debug_umma<<<1, 128, smem, s>>>(mode, a, w, d, n, k);
  DMIP_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DMIP_OK;
}


int launch_debug_umma2(const Umma2Params& P, cudaStream_t s) {
  DMIP_REQUIRE(P.k % 16 == 0 && P.k >= 16 && P.k <= 128, "umma2: k must be a multiple of 16 in [16,128]");
  DMIP_REQUIRE(P.n % 8 == 0 && P.n >= 8 && P.n <= 256, "umma2: n must be a multiple of 8 in [8,256]");
  const int smem = 2 * kU2A + 2 * kU2B + 64 + 1024;
  DMIP_CHECK_CUDA(cudaFuncSetAttribute(k_umma2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k_umma2<<<1, 128, smem, s>>>(P);
  DMIP_CHECK_CUDA(cudaGetLastError());
  return DMIP_OK;
}

int launch_debug_mma_bench2(int cg, int mode, int n, int k, int iters, int stream_bytes, int grid, const void* gsrc,
                            long long* cycles, cudaStream_t s) {
  DMIP_REQUIRE(cg == 1 || cg == 2, "cg must be 1 or 2");
  DMIP_REQUIRE(k % 64 == 0 && k >= 64 && k <= 256, "k must be 64..256");
  DMIP_REQUIRE(n % 16 == 0 && n >= 16 && n <= 256 && n / cg <= 256, "bad n");
  DMIP_REQUIRE(stream_bytes >= 0 && stream_bytes <= kKb && stream_bytes % 16 == 0, "stream_bytes must be <= 16384, multiple of 16");
  DMIP_REQUIRE(grid >= cg && grid % cg == 0, "grid must be a multiple of cg");
  DMIP_REQUIRE(stream_bytes == 0 || gsrc != nullptr, "gsrc (>= 8 MB + 64 KB) needed when streaming");
  const int smem = (12 + kRing) * kKb + 256 + 1024;
  DMIP_CHECK_CUDA(cudaFuncSetAttribute(k_mma_bench2<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  DMIP_CHECK_CUDA(cudaFuncSetAttribute(k_mma_bench2<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  BenchParams P = {};
  P.cg = cg; P.a_tmem = mode & 1; P.alt_acc = (mode >> 1) & 1; P.n = n; P.k = k; P.iters = iters;
  P.rnd = (mode >> 2) & 1; P.commit = (mode >> 3) & 1; P.readers = (mode >> 4) & 3;
  P.reader_iters = iters * (k / 16) / 2;
  P.math = (mode >> 6) & 3;
  P.kbwait = (mode >> 8) & 7;
  P.math_iters = iters * (k / 16) * 8;
  P.sink = reinterpret_cast<float*>(cycles);
  P.stream_bytes = stream_bytes;
  P.gsrc = static_cast<const uint8_t*>(gsrc);
  P.cycles = cycles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(768);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(cg);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = cg > 1 ? 1 : 0;
  if (cg == 2) DMIP_CHECK_CUDA(cudaLaunchKernelEx(&cfg, k_mma_bench2<2>, P));
  else DMIP_CHECK_CUDA(cudaLaunchKernelEx(&cfg, k_mma_bench2<1>, P));
  DMIP_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DMIP_OK;
}

int launch_debug_prim_bench(int iters, long long* out, cudaStream_t s) {
  PrimParams P = {};
  P.iters = iters; P.a = 1; P.b = 2; P.c = 3;
  P.out = out;
  P.sink = reinterpret_cast<unsigned long long*>(out + 16);
  k_prim_bench<<<1, 64, 0, s>>>(P);
  DMIP_CHECK_CUDA(cudaGetLastError());
  k_pingpong<<<1, 64, 0, s>>>(iters, out);
  DMIP_CHECK_CUDA(cudaGetLastError());
  k_wakeup<<<1, 64, 0, s>>>(out);
  DMIP_CHECK_CUDA(cudaGetLastError());
  k_xu_rate<<<1, 512, 0, s>>>(iters, out, reinterpret_cast<float*>(out + 60));
  DMIP_CHECK_CUDA(cudaGetLastError());
  k_mmasync_rate<<<1, 512, 0, s>>>(iters, out + 53, reinterpret_cast<float*>(out + 61));
  DMIP_CHECK_CUDA(cudaGetLastError());
  count_launch(5);
  return DMIP_OK;
}

}  // namespace dmip

using namespace dmip;

extern "C" {

const char* dmip_probe_last_error(void) { return g_probe_err; }

int dmip_debug_mma_bench(int32_t mode, int32_t n, int32_t k, int32_t iters, int32_t grid, void* cycles, void* stream) {
  return launch_debug_mma_bench(mode, n, k, iters, grid, static_cast<long long*>(cycles), static_cast<cudaStream_t>(stream));
}

int dmip_debug_mma_bench2(int32_t cg, int32_t mode, int32_t n, int32_t k, int32_t iters, int32_t stream_bytes, int32_t grid,
                          const void* gsrc, void* cycles, void* stream) {
  return launch_debug_mma_bench2(cg, mode, n, k, iters, stream_bytes, grid, gsrc, static_cast<long long*>(cycles),
                                 static_cast<cudaStream_t>(stream));
}

int dmip_debug_prim_bench(int32_t iters, void* out, void* stream) {
  return launch_debug_prim_bench(iters, static_cast<long long*>(out), static_cast<cudaStream_t>(stream));
}

int dmip_debug_umma(int32_t mode, const float* a, const float* w, float* d, int32_t n, int32_t k, void* stream) {
  DMIP_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (A in smem) or 1 (A in tmem)");
  return launch_debug_umma(mode, a, w, d, n, k, static_cast<cudaStream_t>(stream));
}

/* general GEMM self-test: majors[2] = {a_mn, b_mn}; img[4] = {a_lbo, a_sbo, b_lbo, b_sbo} image strides and
 * fld[4] the descriptor fields (bytes); cycles: device int64[2] or NULL */
int dmip_debug_umma2(const int32_t* majors, int32_t n, int32_t k, int32_t split, int32_t iters, const uint32_t* img,
                     const uint32_t* fld, const float* a, const float* b, float* d, void* cycles, void* stream) {
  Umma2Params P = {};
  P.a_mn = majors[0]; P.b_mn = majors[1]; P.n = n; P.k = k; P.split = split; P.iters = iters;
  P.a_img_lbo = img[0]; P.a_img_sbo = img[1]; P.b_img_lbo = img[2]; P.b_img_sbo = img[3];
  P.a_fld_lbo = fld[0]; P.a_fld_sbo = fld[1]; P.b_fld_lbo = fld[2]; P.b_fld_sbo = fld[3];
  P.a = a; P.b = b; P.d = d; P.cycles = static_cast<long long*>(cycles);
  return launch_debug_umma2(P, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
