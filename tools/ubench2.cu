// ubench2.cu -- pipe-concurrency microbenchmarks for the prune filter (sm_100a):
// can the byte-lane adds run as IMADs on the FMA pipe while the PRMT look-ups keep the ALU pipe busy?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/ubench2 tools/ubench2.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
constexpr int ILP = 8;
struct Res { unsigned long long cyc; };
#define PROLOG __syncthreads(); long long t0 = clock64();
#define EPILOG(val) long long t1 = clock64(); if ((val) == 0x12345u) out[0] = 1; __syncthreads(); \
  if (threadIdx.x == 0) atomicMax(&res->cyc, (unsigned long long)(t1 - t0));

// MODE 0: PRMT + IADD (ptxas pairs them into IADD3)   1: PRMT + IMAD(p, one, acc)   2: PRMT only
// 3: IMAD only   4: 2 PRMT + 1 IMAD + (1 IADD3 for the other)   5: PRMT + IMAD + SHF every 8
template <int MODE>
__global__ void k_mix(int iters, unsigned tl, unsigned th, unsigned one, unsigned* out, Res* res) {
  unsigned acc[ILP], sel[ILP];
  for (int i = 0; i < ILP; i++) { acc[i] = i; sel[i] = (threadIdx.x * 2654435761u + i * 77u) & 0x7777u; }
  // bounce the tables through shared memory so that they live in plain registers
  __shared__ unsigned sh[4];
  if (threadIdx.x == 0) { sh[0] = tl; sh[1] = th; sh[2] = one; }
  __syncthreads();
  unsigned t_lo = *(volatile unsigned*)&sh[0], t_hi = *(volatile unsigned*)&sh[1];
  PROLOG
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < ILP; i++) {
        unsigned p;
        if (MODE != 3) { asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(p) : "r"(t_lo), "r"(t_hi), "r"(sel[i])); sel[i] = p; }
        else p = sel[i];
        if (MODE == 0) acc[i] += p;
        else if (MODE == 1 || MODE == 3 || MODE == 5) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(p), "r"(one));
        else if (MODE == 2) acc[i] ^= p;  // folded into LOP3? keep a consumer
        else if (MODE == 4) { if (i & 1) acc[i] += p; else asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(p), "r"(one)); }
        if (MODE == 5 && (i == 0)) asm volatile("shf.r.clamp.b32 %0, %0, %1, %2;" : "+r"(sel[u]) : "r"(sel[u + 4]), "r"(one));
      }
  }
  unsigned s = 0; for (int i = 0; i < ILP; i++) s ^= acc[i];
  EPILOG(s)
}

// MODE 0: dp2a only   1: dp2a + PRMT
template <int MODE>
__global__ void k_dp2a(int iters, unsigned tl, unsigned th, unsigned w, unsigned* out, Res* res) {
  unsigned acc[ILP], sel[ILP];
  for (int i = 0; i < ILP; i++) { acc[i] = i; sel[i] = (threadIdx.x * 2654435761u + i * 77u) & 0x03030303u; }
  __shared__ unsigned sh[4];
  if (threadIdx.x == 0) { sh[0] = tl; sh[1] = th; }
  __syncthreads();
  unsigned t_lo = *(volatile unsigned*)&sh[0], t_hi = *(volatile unsigned*)&sh[1];
  PROLOG
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < ILP; i++) {
        unsigned d;
        asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(sel[i]), "r"(w), "r"(acc[i]));
        acc[i] = d;
        if (MODE == 1) { unsigned p; asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(p) : "r"(t_lo), "r"(t_hi), "r"(sel[i])); sel[i] = p; }
      }
  }
  unsigned s = 0; for (int i = 0; i < ILP; i++) s ^= acc[i] ^ sel[i];
  EPILOG(s)
}

// Table look-ups with per-lane indices that change every step (nothing can be hoisted):
// MODE 0: LDS.32   1: LDC (constant bank, indexed)   2: one LDS + one LDC per step   3: index update only
__constant__ float ctab[64 * 8];
template <int MODE>
__global__ void k_lookup(int iters, int ndist, unsigned* out, Res* res) {
  __shared__ float stab[64 * 8];
  for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) stab[i] = i * 0.001f;
  unsigned idx[ILP]; float acc[ILP];
  for (int i = 0; i < ILP; i++) { idx[i] = (threadIdx.x * 7 + i * 3 + (threadIdx.x >> 3)) % ndist; acc[i] = 0.f; }
  PROLOG
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < ILP; i++) {
        // next symbol: cheap LCG step folded into 0..7, kept < ndist by the caller's choice of 4/8
        idx[i] = (idx[i] * 5u + 1u) & (unsigned)(ndist - 1);
        const unsigned a = idx[i] + (u * ILP + i) % 8 * 8;   // row * 8 + symbol
        if (MODE == 0) acc[i] += stab[a];
        else if (MODE == 1) acc[i] += ctab[a];
        else if (MODE == 2) acc[i] += ((i & 1) ? stab[a] : ctab[a]);
        else acc[i] += __uint_as_float(a);
      }
  }
  float s = 0; for (int i = 0; i < ILP; i++) s += acc[i];
  EPILOG(__float_as_uint(s))
}

template <typename F>
void run(const char* name, int threads, int bps, int nsm, double ops, int iters, F launch) {
  Res* res; unsigned* out;
  CK(cudaMalloc(&res, sizeof(Res))); CK(cudaMalloc(&out, 64));
  launch(nsm * bps, threads, iters / 8 + 1, out, res);
  CK(cudaDeviceSynchronize());
  CK(cudaMemset(res, 0, sizeof(Res)));
  launch(nsm * bps, threads, iters, out, res);
  CK(cudaDeviceSynchronize());
  Res h; CK(cudaMemcpy(&h, res, sizeof(Res), cudaMemcpyDeviceToHost));
  double tot = (double)threads * bps * ops * iters;
  printf("%-44s thr=%4d bps=%d  %8.2f thread-instr/clk/SM\n", name, threads, bps, tot / (double)h.cyc);
  cudaFree(res); cudaFree(out);
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int nsm = p.multiProcessorCount;
  const int IT = 4096; const double OPS = 4.0 * ILP;
  for (int threads : {256, 768}) {
    int bps = threads == 256 ? 3 : 1;
#define RUN(M, nm, mult) run(nm, threads, bps, nsm, OPS * mult, IT, [&](int g, int t, int it, unsigned* o, Res* r) { k_mix<M><<<g, t>>>(it, 0x03020100u, 0x07060504u, 1u, o, r); });
    RUN(0, "PRMT + IADD (-> 2 PRMT : 1 IADD3), 1.5 instr", 1.5)
    RUN(1, "PRMT + IMAD(real), 2 instr", 2.0)
    RUN(2, "PRMT + LOP3 xor, 2 instr", 2.0)
    RUN(3, "IMAD only, 1 instr", 1.0)
    RUN(4, "2 PRMT + 1 IMAD + 1 IADD, 2 instr/pair", 2.0)
    RUN(5, "PRMT + IMAD + SHF/8", 2.125)
    run("lookup LDS.32 (4 symbols)", threads, bps, nsm, OPS, IT, [&](int g, int t, int it, unsigned* o, Res* r) { k_lookup<0><<<g, t>>>(it, 4, o, r); });
    run("lookup LDC indexed (4 symbols)", threads, bps, nsm, OPS, IT, [&](int g, int t, int it, unsigned* o, Res* r) { k_lookup<1><<<g, t>>>(it, 4, o, r); });
    run("lookup LDS + LDC alternating", threads, bps, nsm, OPS, IT, [&](int g, int t, int it, unsigned* o, Res* r) { k_lookup<2><<<g, t>>>(it, 4, o, r); });
    run("index update only", threads, bps, nsm, OPS, IT, [&](int g, int t, int it, unsigned* o, Res* r) { k_lookup<3><<<g, t>>>(it, 4, o, r); });
    run("dp2a only, 1 instr", threads, bps, nsm, OPS, IT, [&](int g, int t, int it, unsigned* o, Res* r) { k_dp2a<0><<<g, t>>>(it, 0x03020100u, 0x07060504u, 0x00001001u, o, r); });
    run("dp2a + PRMT, 2 instr", threads, bps, nsm, OPS * 2, IT, [&](int g, int t, int it, unsigned* o, Res* r) { k_dp2a<1><<<g, t>>>(it, 0x03020100u, 0x07060504u, 0x00001001u, o, r); });
  }
  return 0;
}
