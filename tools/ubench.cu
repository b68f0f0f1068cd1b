// ubench.cu -- instruction-throughput microbenchmarks for the design of the
// scoring kernels (sm_100a).  Not part of the product; results are summarised
// in profiles/ and DESIGN.md.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench tools/ubench.cu
//
// Every kernel runs `iters` iterations of an unrolled body with ILP
// independent chains per thread and reports thread-ops / clk / SM measured
// with clock64() inside the kernel (max over blocks).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ILP = 8;
__constant__ float cw[256];

struct Res { unsigned long long cyc; };

#define KERNEL_PROLOG \
  __syncthreads(); long long t0 = clock64();
#define KERNEL_EPILOG(val) \
  long long t1 = clock64(); \
  if ((val) == 12345.678f) out[0] = 1; \
  __syncthreads(); \
  if (threadIdx.x == 0) atomicMax(&res->cyc, (unsigned long long)(t1 - t0));

__global__ void k_fadd_rr(int iters, float b, float* out, Res* res) {
  float a[ILP];
  for (int i = 0; i < ILP; i++) a[i] = threadIdx.x + i;
  KERNEL_PROLOG
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < ILP; i++) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
  }
  float s = 0; for (int i = 0; i < ILP; i++) s += a[i];
  KERNEL_EPILOG(s)
}

__global__ void k_fadd_rc(int iters, float* out, Res* res) {
  float a[ILP];
  for (int i = 0; i < ILP; i++) a[i] = threadIdx.x + i;
  KERNEL_PROLOG
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < ILP; i++) a[i] = __fadd_rn(a[i], cw[(u * ILP + i)]);
  }
  float s = 0; for (int i = 0; i < ILP; i++) s += a[i];
  KERNEL_EPILOG(s)
}

__global__ void k_ffma_rrr(int iters, float b, float c, float* out, Res* res) {
  float a[ILP];
  for (int i = 0; i < ILP; i++) a[i] = threadIdx.x + i;
  KERNEL_PROLOG
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < ILP; i++) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(b), "f"(c));
  }
  float s = 0; for (int i = 0; i < ILP; i++) s += a[i];
  KERNEL_EPILOG(s)
}

// s = fma(m, const, s): the "one-hot select-and-add" form
__global__ void k_ffma_rc(int iters, float m0, float* out, Res* res) {
  float a[ILP];
  float m = m0 + threadIdx.x * 0.f;
  for (int i = 0; i < ILP; i++) a[i] = threadIdx.x + i;
  KERNEL_PROLOG
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < ILP; i++) a[i] = __fmaf_rn(m, cw[u * ILP + i], a[i]);
  }
  float s = 0; for (int i = 0; i < ILP; i++) s += a[i];
  KERNEL_EPILOG(s)
}

__global__ void k_fadd2(int iters, float b, float* out, Res* res) {
  unsigned long long a[ILP], bb;
  asm volatile("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
  for (int i = 0; i < ILP; i++) { float x = threadIdx.x + i; asm volatile("mov.b64 %0, {%1, %1};" : "=l"(a[i]) : "f"(x)); }
  KERNEL_PROLOG
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < ILP; i++) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(bb));
  }
  float s = 0;
  for (int i = 0; i < ILP; i++) { float x, y; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a[i])); s += x + y; }
  KERNEL_EPILOG(s)
}

__global__ void k_ffma2(int iters, float b, float c, float* out, Res* res) {
  unsigned long long a[ILP], bb, cc;
  asm volatile("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
  asm volatile("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
  for (int i = 0; i < ILP; i++) { float x = threadIdx.x + i; asm volatile("mov.b64 %0, {%1, %1};" : "=l"(a[i]) : "f"(x)); }
  KERNEL_PROLOG
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < ILP; i++) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a[i]) : "l"(bb), "l"(cc));
  }
  float s = 0;
  for (int i = 0; i < ILP; i++) { float x, y; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a[i])); s += x + y; }
  KERNEL_EPILOG(s)
}

// integer / logic ops
#define INT_KERNEL(name, ASM)                                                   \
  __global__ void name(int iters, unsigned b, unsigned c, float* out, Res* res) { \
    unsigned a[ILP];                                                            \
    for (int i = 0; i < ILP; i++) a[i] = threadIdx.x * 2654435761u + i;         \
    KERNEL_PROLOG                                                               \
    for (int it = 0; it < iters; it++) {                                        \
      _Pragma("unroll") for (int u = 0; u < 4; u++)                              \
      _Pragma("unroll") for (int i = 0; i < ILP; i++)                            \
        asm volatile(ASM : "+r"(a[i]) : "r"(b), "r"(c));                        \
    }                                                                           \
    unsigned s = 0; for (int i = 0; i < ILP; i++) s ^= a[i];                    \
    KERNEL_EPILOG(__uint_as_float(s))                                           \
  }

INT_KERNEL(k_prmt, "prmt.b32 %0, %0, %1, %2;")
INT_KERNEL(k_lop3, "lop3.b32 %0, %0, %1, %2, 0x96;")
INT_KERNEL(k_iadd3, "{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }")
INT_KERNEL(k_iadd, "add.u32 %0, %0, %1;")
INT_KERNEL(k_imad, "mad.lo.u32 %0, %0, %1, %2;")
INT_KERNEL(k_shf, "shf.l.wrap.b32 %0, %0, %1, %2;")
INT_KERNEL(k_shl, "shl.b32 %0, %0, 3;")
INT_KERNEL(k_sel, "{ .reg .pred p; setp.gt.u32 p, %1, %2; selp.u32 %0, %0, %1, p; }")
INT_KERNEL(k_vadd4, "{ .reg .u32 t; vadd4.u32.u32.u32.sat t, %0, %1, %2; mov.u32 %0, t; }")
INT_KERNEL(k_vadd2, "{ .reg .u32 t; vadd2.u32.u32.u32.sat t, %0, %1, %2; mov.u32 %0, t; }")

// half2 add
__global__ void k_hadd2(int iters, unsigned b, float* out, Res* res) {
  unsigned a[ILP];
  for (int i = 0; i < ILP; i++) a[i] = 0x3c003c00u + i;
  KERNEL_PROLOG
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < ILP; i++) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
  }
  unsigned s = 0; for (int i = 0; i < ILP; i++) s ^= a[i];
  KERNEL_EPILOG(__uint_as_float(s))
}

// ---- shared memory look-ups.  pattern: each lane has a fixed byte offset
// (one of `ndist` distinct rows); each step reads smem[off + step*stride].
template <int VEC>
__global__ void k_lds(int iters, int ndist, int rowbytes, float* out, Res* res) {
  extern __shared__ float sm[];
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = i * 1e-9f;
  unsigned off[ILP];
  for (int i = 0; i < ILP; i++)
    off[i] = (unsigned)(((threadIdx.x * 7 + i * 3 + (threadIdx.x >> 3)) % ndist) * rowbytes);
  unsigned base = (unsigned)__cvta_generic_to_shared(sm);
  float acc[ILP];
  for (int i = 0; i < ILP; i++) acc[i] = 0;
  KERNEL_PROLOG
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < ILP; i++) {
        unsigned addr = base + off[i] + (u * 16 * VEC / 4);
        if (VEC == 1) {
          float v; asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
          acc[i] += v;
        } else if (VEC == 2) {
          float v, w; asm volatile("ld.volatile.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v), "=f"(w) : "r"(addr));
          acc[i] += v; acc[(i + 1) % ILP] += w;
        } else {
          float v, w, x, y; asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v), "=f"(w), "=f"(x), "=f"(y) : "r"(addr));
          acc[i] += v; acc[(i + 1) % ILP] += w; acc[(i + 2) % ILP] += x; acc[(i + 3) % ILP] += y;
        }
      }
  }
  float s = 0; for (int i = 0; i < ILP; i++) s += acc[i];
  KERNEL_EPILOG(s)
}

// random gather from a table of `entries` floats (prefix-memo pattern)
__global__ void k_lds_rand(int iters, int entries, float* out, Res* res) {
  extern __shared__ float sm[];
  for (int i = threadIdx.x; i < entries; i += blockDim.x) sm[i] = i * 1e-9f;
  unsigned x[ILP];
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 2654435761u + i * 40503u + blockIdx.x;
  unsigned base = (unsigned)__cvta_generic_to_shared(sm);
  float acc[ILP];
  for (int i = 0; i < ILP; i++) acc[i] = 0;
  unsigned mask = (unsigned)(entries - 1) << 2;
  KERNEL_PROLOG
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < ILP; i++) {
        x[i] = x[i] * 1664525u + 1013904223u;
        unsigned addr = base + ((x[i] >> 8) & mask);
        float v; asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
        acc[i] += v;
      }
  }
  float s = 0; for (int i = 0; i < ILP; i++) s += acc[i];
  KERNEL_EPILOG(s)
}

// constant-bank indexed (divergent) loads
__global__ void k_ldc(int iters, int ndist, float* out, Res* res) {
  unsigned off[ILP];
  for (int i = 0; i < ILP; i++) off[i] = (threadIdx.x * 7 + i * 3 + (threadIdx.x >> 3)) % ndist;
  float acc[ILP];
  for (int i = 0; i < ILP; i++) acc[i] = 0;
  KERNEL_PROLOG
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < ILP; i++) acc[i] += cw[off[i] + u * 8];
  }
  float s = 0; for (int i = 0; i < ILP; i++) s += acc[i];
  KERNEL_EPILOG(s)
}

__global__ void k_shfl(int iters, float* out, Res* res) {
  float a[ILP];
  int src[ILP];
  for (int i = 0; i < ILP; i++) { a[i] = threadIdx.x + i; src[i] = (threadIdx.x * 5 + i) & 31; }
  KERNEL_PROLOG
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < ILP; i++) a[i] = __shfl_sync(0xffffffffu, a[i], src[i]);
  }
  float s = 0; for (int i = 0; i < ILP; i++) s += a[i];
  KERNEL_EPILOG(s)
}

// mixed: one PRMT + one FADD per step (ALU + FMA pipes in parallel?)
__global__ void k_mix_prmt_fadd(int iters, unsigned b, unsigned c, float fb, float* out, Res* res) {
  unsigned a[ILP]; float f[ILP];
  for (int i = 0; i < ILP; i++) { a[i] = threadIdx.x * 2654435761u + i; f[i] = i; }
  KERNEL_PROLOG
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < ILP; i++) {
        asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
        asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fb));
      }
  }
  float s = 0; for (int i = 0; i < ILP; i++) s += f[i] + __uint_as_float(a[i]);
  KERNEL_EPILOG(s)
}

// mixed: PRMT + IMAD (add on the FMA pipe)
__global__ void k_mix_prmt_imad(int iters, unsigned b, unsigned c, float* out, Res* res) {
  unsigned a[ILP], f[ILP];
  for (int i = 0; i < ILP; i++) { a[i] = threadIdx.x * 2654435761u + i; f[i] = i; }
  KERNEL_PROLOG
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < ILP; i++) {
        asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
        asm volatile("mad.lo.u32 %0, %1, 1, %0;" : "+r"(f[i]) : "r"(a[i]));
      }
  }
  unsigned s = 0; for (int i = 0; i < ILP; i++) s ^= f[i] ^ a[i];
  KERNEL_EPILOG(__uint_as_float(s))
}

// mixed: the exact inner loop LDS.32 + FADD (look-up then add), 5 distinct rows
__global__ void k_mix_lds_fadd(int iters, int ndist, int rowbytes, float* out, Res* res) {
  extern __shared__ float sm[];
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = i * 1e-9f;
  unsigned off[ILP];
  for (int i = 0; i < ILP; i++)
    off[i] = (unsigned)(((threadIdx.x * 7 + i * 3 + (threadIdx.x >> 3)) % ndist) * rowbytes);
  unsigned base = (unsigned)__cvta_generic_to_shared(sm);
  float acc[ILP];
  for (int i = 0; i < ILP; i++) acc[i] = 0;
  KERNEL_PROLOG
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < ILP; i++) {
        float v; asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + off[i] + u * 4));
        asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(acc[i]) : "f"(v));
      }
  }
  float s = 0; for (int i = 0; i < ILP; i++) s += acc[i];
  KERNEL_EPILOG(s)
}

template <typename F>
void run(const char* name, int threads, int blocks_per_sm, int nsm, double ops_per_thread_iter,
         int iters, size_t smem, F launch) {
  Res* res; float* out;
  CK(cudaMalloc(&res, sizeof(Res))); CK(cudaMalloc(&out, 64));
  CK(cudaMemset(res, 0, sizeof(Res)));
  int grid = nsm * blocks_per_sm;
  launch(grid, threads, smem, iters / 8 + 1, out, res);  // warm-up
  CK(cudaDeviceSynchronize());
  CK(cudaMemset(res, 0, sizeof(Res)));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  launch(grid, threads, smem, iters, out, res);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  Res h; CK(cudaMemcpy(&h, res, sizeof(Res), cudaMemcpyDeviceToHost));
  double ops_sm = (double)threads * blocks_per_sm * ops_per_thread_iter * iters;
  printf("%-28s thr=%4d bps=%d  %8.2f thread-ops/clk/SM  (%.3f ms, %llu cyc, eff clk %.0f MHz)\n",
         name, threads, blocks_per_sm, ops_sm / (double)h.cyc, ms, h.cyc, h.cyc / (ms * 1e3));
  cudaFree(res); cudaFree(out);
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int nsm = p.multiProcessorCount;
  printf("device %s, %d SMs, clock %d kHz\n", p.name, nsm, p.clockRate);
  float hw[256]; for (int i = 0; i < 256; i++) hw[i] = 1.0f + i * 1e-3f;
  CK(cudaMemcpyToSymbol(cw, hw, sizeof(hw)));
  const int IT = 4096;
  const double OPS = 4.0 * ILP;
  for (int threads : {512, 1024}) {
    int bps = 1;
    run("fadd reg,reg", threads, bps, nsm, OPS, IT, 0, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_fadd_rr<<<g, t, s>>>(it, 1.0001f, o, r); });
    run("fadd reg,const", threads, bps, nsm, OPS, IT, 0, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_fadd_rc<<<g, t, s>>>(it, o, r); });
    run("ffma reg,reg,reg", threads, bps, nsm, OPS, IT, 0, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_ffma_rrr<<<g, t, s>>>(it, 1.0001f, 0.5f, o, r); });
    run("ffma reg,const,acc", threads, bps, nsm, OPS, IT, 0, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_ffma_rc<<<g, t, s>>>(it, 1.0f, o, r); });
    run("fadd2 (f32x2) [x2 adds]", threads, bps, nsm, OPS, IT, 0, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_fadd2<<<g, t, s>>>(it, 1.0001f, o, r); });
    run("ffma2 (f32x2) [x2 fmas]", threads, bps, nsm, OPS, IT, 0, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_ffma2<<<g, t, s>>>(it, 1.0001f, 0.5f, o, r); });
    run("hadd2", threads, bps, nsm, OPS, IT, 0, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_hadd2<<<g, t, s>>>(it, 0x3c003c00u, o, r); });
    run("prmt", threads, bps, nsm, OPS, IT, 0, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_prmt<<<g, t, s>>>(it, 0x12345678u, 0x3210u, o, r); });
    run("lop3", threads, bps, nsm, OPS, IT, 0, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_lop3<<<g, t, s>>>(it, 0x12345678u, 0x0f0f3210u, o, r); });
    run("iadd", threads, bps, nsm, OPS, IT, 0, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_iadd<<<g, t, s>>>(it, 0x12345678u, 0x3210u, o, r); });
    run("iadd3 (2 adds)", threads, bps, nsm, OPS, IT, 0, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_iadd3<<<g, t, s>>>(it, 0x12345678u, 0x3210u, o, r); });
    run("imad", threads, bps, nsm, OPS, IT, 0, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_imad<<<g, t, s>>>(it, 3u, 0x3210u, o, r); });
    run("shf", threads, bps, nsm, OPS, IT, 0, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_shf<<<g, t, s>>>(it, 0x12345678u, 5u, o, r); });
    run("shl imm", threads, bps, nsm, OPS, IT, 0, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_shl<<<g, t, s>>>(it, 0, 0, o, r); });
    run("setp+selp", threads, bps, nsm, OPS, IT, 0, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_sel<<<g, t, s>>>(it, 5u, 3u, o, r); });
    run("vadd4.sat", threads, bps, nsm, OPS, IT, 0, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_vadd4<<<g, t, s>>>(it, 0x01020304u, 0u, o, r); });
    run("vadd2.sat", threads, bps, nsm, OPS, IT, 0, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_vadd2<<<g, t, s>>>(it, 0x01020304u, 0u, o, r); });
    run("shfl idx", threads, bps, nsm, OPS, IT, 0, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_shfl<<<g, t, s>>>(it, o, r); });
    run("mix prmt+fadd (2 ops)", threads, bps, nsm, 2 * OPS, IT, 0, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_mix_prmt_fadd<<<g, t, s>>>(it, 0x12345678u, 0x3210u, 1.0001f, o, r); });
    run("mix prmt+imad (2 ops)", threads, bps, nsm, 2 * OPS, IT, 0, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_mix_prmt_imad<<<g, t, s>>>(it, 0x12345678u, 0x3210u, o, r); });
    size_t SM = 65536;
    CK(cudaFuncSetAttribute(k_lds<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM));
    CK(cudaFuncSetAttribute(k_lds<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM));
    CK(cudaFuncSetAttribute(k_lds<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM));
    CK(cudaFuncSetAttribute(k_lds_rand, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM));
    CK(cudaFuncSetAttribute(k_mix_lds_fadd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM));
    int bl = 1;
    run("lds.32 5-row bcast", threads, bl, nsm, OPS, IT, SM, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_lds<1><<<g, t, s>>>(it, 5, 96, o, r); });
    run("lds.32 1-row bcast", threads, bl, nsm, OPS, IT, SM, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_lds<1><<<g, t, s>>>(it, 1, 96, o, r); });
    run("lds.32 21-row (protein)", threads, bl, nsm, OPS, IT, SM, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_lds<1><<<g, t, s>>>(it, 21, 132, o, r); });
    run("lds.64 5-row [x2 values]", threads, bl, nsm, OPS, IT, SM, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_lds<2><<<g, t, s>>>(it, 5, 96, o, r); });
    run("lds.128 5-row [x4 values]", threads, bl, nsm, OPS, IT, SM, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_lds<4><<<g, t, s>>>(it, 5, 112, o, r); });
    run("lds.32 16 entries x 4 B", threads, bl, nsm, OPS, IT, SM, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_lds<1><<<g, t, s>>>(it, 16, 4, o, r); });
    run("lds.64 16 entries x 8 B [x2 values]", threads, bl, nsm, OPS, IT, SM, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_lds<2><<<g, t, s>>>(it, 16, 8, o, r); });
    run("lds.64 25 entries x 8 B [x2 values]", threads, bl, nsm, OPS, IT, SM, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_lds<2><<<g, t, s>>>(it, 25, 8, o, r); });
    run("lds.128 8 entries x 16 B [x4 values]", threads, bl, nsm, OPS, IT, SM, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_lds<4><<<g, t, s>>>(it, 8, 16, o, r); });
    run("lds.128 16 entries x 16 B [x4 values]", threads, bl, nsm, OPS, IT, SM, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_lds<4><<<g, t, s>>>(it, 16, 16, o, r); });
    run("lds.128 1-row [x4 values]", threads, bl, nsm, OPS, IT, SM, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_lds<4><<<g, t, s>>>(it, 1, 112, o, r); });
    run("lds.32 random 4096", threads, bl, nsm, OPS, IT, SM, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_lds_rand<<<g, t, s>>>(it, 4096, o, r); });
    run("lds.32 random 16384", threads, bl, nsm, OPS, IT, SM, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_lds_rand<<<g, t, s>>>(it, 16384, o, r); });
    run("ldc divergent 5", threads, bps, nsm, OPS, IT / 4, 0, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_ldc<<<g, t, s>>>(it, 5, o, r); });
    run("ldc uniform 1", threads, bps, nsm, OPS, IT / 4, 0, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_ldc<<<g, t, s>>>(it, 1, o, r); });
    run("mix lds.32+fadd (2 ops)", threads, bl, nsm, 2 * OPS, IT, SM, [&](int g, int t, size_t s, int it, float* o, Res* r) { k_mix_lds_fadd<<<g, t, s>>>(it, 5, 96, o, r); });
  }
  return 0;
}
