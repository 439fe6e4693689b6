// Developer probe (not part of the library): how many shared-memory wavefronts one LDS.32/64/128
// costs for a given lane -> address pattern, and what POPC rate an SM sustains with and without
// shared loads in the mix. Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build_variants/lds_probe tools/lds_probe.cu; tools/gpu_evidence.sh runs it under ncu.
//
//   lds_probe lds     : prints cycles per LDS for every (pattern, width); run it under
//                       ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum for the counts
//   lds_probe popc    : POPC throughput (per SM and clock) for 8..32 resident warps
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                 \
  do {                                                                        \
    cudaError_t e_ = (x);                                                     \
    if (e_ != cudaSuccess) {                                                  \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                \
      exit(1);                                                                \
    }                                                                         \
  } while (0)

constexpr int kIters = 4096;

template <int W>  // bytes per lane
__global__ void k_lds(const int* idx_of_lane, unsigned long long* out_cycles, uint32_t* sink) {
  __shared__ __align__(16) uint32_t sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 2654435761u;
  __syncthreads();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + idx_of_lane[threadIdx.x & 31] * W;
  uint32_t acc = 0;
  const long long t0 = clock64();
#pragma unroll 8
  for (int it = 0; it < kIters; ++it) {
    const uint32_t a = base + ((it & 7) * 512);
    if constexpr (W == 16) {
      uint32_t x, y, z, w;
      asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(a));
      acc += x ^ y ^ z ^ w;
    } else if constexpr (W == 8) {
      uint32_t x, y;
      asm volatile("ld.volatile.shared.v2.u32 {%0,%1}, [%2];" : "=r"(x), "=r"(y) : "r"(a));
      acc += x ^ y;
    } else {
      uint32_t x;
      asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(x) : "r"(a));
      acc += x;
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out_cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (acc == 0x12345u) sink[0] = acc;
}

// POPC rate: every thread runs `kIters` rounds of 16 independent (LOP3, LOP3, POPC, IMAD) chains;
// with lds = 1 each round also performs 8 LDS.128 (4 wavefronts each) whose values feed the chains
template <int LDS>
__global__ void __launch_bounds__(256, 4) k_popc(unsigned long long* out_cycles, uint32_t* sink, int iters) {
  __shared__ __align__(16) uint32_t sm[3 * 4096];
  for (int i = threadIdx.x; i < 3 * 4096; i += blockDim.x) sm[i] = i * 2654435761u;
  __syncthreads();
  uint32_t cnt[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) cnt[c] = 0;
  const uint4* st = reinterpret_cast<const uint4*>(sm);
  uint32_t x = threadIdx.x * 0x9E3779B9u + 1, y = x * 31 + 7;
  uint32_t xs[4] = {x, y, x ^ (y >> 5), y * 77u};
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    uint4 v[LDS ? LDS : 1];
    if constexpr (LDS > 0) {
#pragma unroll
      for (int k = 0; k < LDS; ++k) {
        const uint32_t ad = (uint32_t)__cvta_generic_to_shared(st + ((threadIdx.x ^ k) + 256 * (k % 3) + 16 * (it & 1)));
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[k].x), "=r"(v[k].y), "=r"(v[k].z), "=r"(v[k].w) : "r"(ad));
      }
    } else {
      v[0] = make_uint4(x, y, x ^ it, y + it);
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const uint4& q = v[LDS ? c % LDS : 0];
      const uint32_t a = q.x ^ q.z, b = q.y ^ q.w;  // every component is used (ptxas splits a v4 load otherwise)
      const uint32_t m0 = xs[c & 3], m1 = xs[c >> 2];
      const uint32_t t0w = (a & m0) | (b & m1);
      const uint32_t t1w = (a & ~m1) | (b & ~m0);
      cnt[c] += __popc(t0w) + (__popc(t1w) << 16);
    }
    xs[0] += x;
    xs[2] ^= y;
    x = x * 1664525u + 1013904223u;
    y ^= x >> 3;
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int c = 0; c < 16; ++c) s += cnt[c];
  if (threadIdx.x == 0) out_cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (s == 0x12345u) sink[0] = s;
}

// XU peak: 16 dependent POPC chains per thread, one LOP3 per POPC
__global__ void __launch_bounds__(256, 4) k_popc_peak(unsigned long long* out_cycles, uint32_t* sink, int iters) {
  uint32_t x[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) x[c] = threadIdx.x * 0x9E3779B9u + c * 0x85EBCA6Bu;
  const uint32_t m = blockIdx.x * 2246822519u + 1;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 16; ++c) x[c] = __popc(x[c]) ^ m ^ (x[c] << 7);
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int c = 0; c < 16; ++c) s += x[c];
  if (threadIdx.x == 0) out_cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (s == 0x12345u) sink[0] = s;
}

struct Pattern {
  const char* name;
  int idx[32];
};

int main(int argc, char** argv) {
  const char* what = argc > 1 ? argv[1] : "lds";
  unsigned long long* d_cyc;
  uint32_t* d_sink;
  int* d_idx;
  CK(cudaMalloc(&d_cyc, 8 * 4096));
  CK(cudaMalloc(&d_sink, 4));
  CK(cudaMalloc(&d_idx, 32 * 4));
  if (!strcmp(what, "lds")) {
    std::vector<Pattern> ps;
    auto add = [&](const char* name, auto f) {
      Pattern p;
      p.name = name;
      for (int l = 0; l < 32; ++l) p.idx[l] = f(l);
      ps.push_back(p);
    };
    add("distinct", [](int l) { return l; });
    add("ignore_bit4 (abd now)", [](int l) { return l & 15; });
    add("ignore_bit3", [](int l) { return (l & 7) | ((l >> 4) << 3); });
    add("ignore_bit2", [](int l) { return (l & 3) | ((l >> 3) << 2); });
    add("ignore_bit1", [](int l) { return (l & 1) | ((l >> 2) << 1); });
    add("ignore_bit0", [](int l) { return l >> 1; });
    add("ignore_bits01 (bcd now)", [](int l) { return l >> 2; });
    add("ignore_bits23 (acd now)", [](int l) { return (l & 3) | ((l >> 4) << 2); });
    add("ignore_bits34", [](int l) { return l & 7; });
    add("ignore_bits12", [](int l) { return (l & 1) | ((l >> 3) << 1); });
    add("ignore_bits04", [](int l) { return (l >> 1) & 7; });
    add("ignore_bits14", [](int l) { return (l & 1) | (((l >> 2) & 3) << 1); });
    add("ignore_bits24", [](int l) { return (l & 3) | (((l >> 3) & 1) << 2); });
    add("ignore_bits02", [](int l) { return ((l >> 1) & 1) | ((l >> 3) << 1); });
    add("ignore_bits03", [](int l) { return ((l >> 1) & 3) | ((l >> 4) << 2); });
    add("ignore_bits13", [](int l) { return (l & 1) | (((l >> 2) & 1) << 1) | ((l >> 4) << 2); });
    add("ignore_bits012", [](int l) { return l >> 3; });
    add("ignore_bits234", [](int l) { return l & 3; });
    add("all_same", [](int) { return 0; });
    add("stride2 (conflict)", [](int l) { return l * 2; });
    add("pairs_swapped", [](int l) { return l ^ 1; });
    add("reversed", [](int l) { return 31 - l; });
    add("rot8", [](int l) { return (l + 8) & 31; });
    for (auto& p : ps) {
      CK(cudaMemcpy(d_idx, p.idx, 128, cudaMemcpyHostToDevice));
      unsigned long long c16, c8, c4;
      k_lds<16><<<1, 256>>>(d_idx, d_cyc, d_sink);
      CK(cudaMemcpy(&c16, d_cyc, 8, cudaMemcpyDeviceToHost));
      k_lds<8><<<1, 256>>>(d_idx, d_cyc, d_sink);
      CK(cudaMemcpy(&c8, d_cyc, 8, cudaMemcpyDeviceToHost));
      k_lds<4><<<1, 256>>>(d_idx, d_cyc, d_sink);
      CK(cudaMemcpy(&c4, d_cyc, 8, cudaMemcpyDeviceToHost));
      printf("%-28s cycles/LDS: .128 %.2f  .64 %.2f  .32 %.2f\n", p.name, (double)c16 / kIters, (double)c8 / kIters, (double)c4 / kIters);
    }
  } else {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int iters = 4096;
    int khz = 0;
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    auto run = [&](auto kern, const char* name, int ctas_per_sm) {
      kern<<<sms * ctas_per_sm, 256>>>(d_cyc, d_sink, 16);  // warm-up
      CK(cudaEventRecord(e0));
      kern<<<sms * ctas_per_sm, 256>>>(d_cyc, d_sink, iters);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      std::vector<unsigned long long> c(sms * ctas_per_sm);
      CK(cudaMemcpy(c.data(), d_cyc, 8 * c.size(), cudaMemcpyDeviceToHost));
      double mean = 0;
      for (auto v : c) mean += (double)v / c.size();
      // The kernel's elapsed time is what counts (ncu's sm__cycles_elapsed agrees with it). The clock64 span of warp 0
      // is printed only as a warning: the scheduler favours the oldest warp, warp 0 finishes long before its CTA, and
      // its span once made this probe report 25 POPC lanes per clock and SM instead of 16.
      const double clocks = ms * 1e-3 * khz * 1e3;  // at the nominal SM clock
      const double lanes = (double)ctas_per_sm * 256 * iters * 32 / clocks;
      printf("%-22s %d CTAs/SM: %.3f ms = %.0f clk at %d MHz: %.2f POPC lanes/clk/SM (XU peak 16) | warp-0 span %.0f clk (misleading)\n", name,
             ctas_per_sm, ms, clocks, khz / 1000, lanes, mean);
    };
    for (int c = 1; c <= 4; ++c) run(k_popc_peak, "popc chains (peak)", c);
    for (int c = 1; c <= 4; ++c) run(k_popc<0>, "popc only", c);
    for (int c = 1; c <= 4; ++c) run(k_popc<8>, "popc + 8 LDS.128", c);
    for (int c = 1; c <= 4; ++c) run(k_popc<16>, "popc + 16 LDS.128", c);
  }
  return 0;
}
