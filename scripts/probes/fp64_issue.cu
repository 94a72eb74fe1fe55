// FP64 issue-rate probes for sm_100a: how many SMSP cycles does a warp-wide DFMA / DMUL / DADD cost, as a
// function of operand pattern (constant vs distinct register operands), dependency-chain count and resident warps?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_issue fp64_issue.cu && ./fp64_issue
#include <cstdio>
#include <cuda_runtime.h>

#define FMA(d, a, b, c) asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a), "d"(b), "d"(c))
#define MUL(d, a, b) asm volatile("mul.rn.f64 %0, %1, %2;" : "=d"(d) : "d"(a), "d"(b))
#define ADD(d, a, b) asm volatile("add.rn.f64 %0, %1, %2;" : "=d"(d) : "d"(a), "d"(b))
#define MULF(d, a, b) asm volatile("fma.rn.f64 %0, %1, %2, 0d8000000000000000;" : "=d"(d) : "d"(a), "d"(b))
#define ADDF(d, a, b) asm volatile("fma.rn.f64 %0, %1, 0d3FF0000000000000, %2;" : "=d"(d) : "d"(a), "d"(b))

template <int MODE>
__global__ void probe(double* out, int iters, double seed) {
  double a[8], x[8], y[8];
  int n[8] = {0, 1, 2, 3, 4, 5, 6, 7};
#pragma unroll
  for (int k = 0; k < 8; ++k) { a[k] = threadIdx.x * 1e-9 + k; x[k] = 1.0 + seed * (k + 1); y[k] = seed * (k + 2); }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (MODE == 0) FMA(a[k], a[k], x[0], y[0]);             // one varying operand, two shared
      if (MODE == 1) FMA(a[k], a[k], x[k], y[k]);             // three distinct register operands
      if (MODE == 2) FMA(a[k], x[k], y[k], a[k]);             // accumulate form
      if (MODE == 3) MUL(a[k], a[k], x[k]);
      if (MODE == 4) ADD(a[k], a[k], y[k]);
      if (MODE == 5) { if (k & 1) MUL(a[k], a[k], x[k]); else FMA(a[k], a[k], x[k], y[k]); }
      if (MODE == 6) FMA(a[k], a[(k + 1) & 7], x[k], y[(k + 3) & 7]);   // operands spread over many registers
      if (MODE == 7) { if (k % 3 == 0) FMA(a[k], a[k], x[k], y[k]); else if (k % 3 == 1) MUL(a[k], a[k], x[k]); else ADD(a[k], a[k], y[k]); }
      if (MODE == 8) { FMA(a[k], a[k], x[k], y[k]); asm volatile("xor.b32 %0, %0, %1;" : "+r"(n[k]) : "r"(i)); }  // 1 ALU op per DFMA
      if (MODE == 9) { if (k & 1) MUL(a[k], a[k], x[k]); else FMA(a[k], a[k], x[k], y[k]);
                       asm volatile("xor.b32 %0, %0, %1;" : "+r"(n[k]) : "r"(i)); }
      if (MODE == 10) FMA(a[0], a[0], x[k], y[k]);              // ONE dependent chain: latency
      if (MODE == 11) { if (k < 4) FMA(a[k], a[k], x[k], y[k]); else MUL(a[k], a[k], x[k]); }   // 4 DFMA then 4 DMUL
      // round 2: is the mixing penalty a property of the OPCODE?  DMUL as DFMA(a, x, -RZ), DADD as DFMA(a, 1.0, y):
      // bit-identical results (x*y + -0.0 == x*y, a*1 + y == a + y in round-to-nearest), one opcode on the pipe.
      if (MODE == 12) { if (k & 1) MULF(a[k], a[k], x[k]); else FMA(a[k], a[k], x[k], y[k]); }
      if (MODE == 13) { if (k % 3 == 0) FMA(a[k], a[k], x[k], y[k]); else if (k % 3 == 1) MULF(a[k], a[k], x[k]); else ADDF(a[k], a[k], y[k]); }
      // the headline kernel's own mix (456 DFMA : 275 DMUL : 139 DADD per 870): 4 : 3 : 1 of every 8
      if (MODE == 14) { if (k < 4) FMA(a[k], a[k], x[k], y[k]); else if (k < 7) MUL(a[k], a[k], x[k]); else ADD(a[k], a[k], y[k]); }
      if (MODE == 15) { if (k < 4) FMA(a[k], a[k], x[k], y[k]); else if (k < 7) MULF(a[k], a[k], x[k]); else ADDF(a[k], a[k], y[k]); }
      if (MODE == 16) { if (k & 1) ADD(a[k], a[k], y[k]); else FMA(a[k], a[k], x[k], y[k]); }   // DFMA / DADD alternating
      if (MODE == 17) { if (k & 1) ADD(a[k], a[k], y[k]); else MUL(a[k], a[k], x[k]); }         // DMUL / DADD alternating
      if (MODE == 18) { FMA(a[k], a[k], x[k], y[k]);                                              // 1 DSETP per 4 DFMA
                        if ((k & 3) == 3) asm volatile("{.reg .pred p; setp.gt.f64 p, %1, %2; selp.b32 %0, %0, %3, p;}" : "+r"(n[k]) : "d"(a[k]), "d"(x[k]), "r"(i)); }
      if (MODE == 19) { FMA(a[k], a[k], x[k], y[k]); if (k & 1) asm volatile("xor.b32 %0, %0, %1;" : "+r"(n[k]) : "r"(i)); }  // 1 ALU per 2 DFMA
      if (MODE == 20) { FMA(a[k], a[k], x[k], y[k]);                                              // 1 SHFL per 4 DFMA
                        if ((k & 3) == 3) asm volatile("shfl.sync.up.b32 %0, %0, 1, 0, 0xffffffff;" : "+r"(n[k])); }
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k] + n[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int threads, int blocks_per_sm) {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  int clock_khz = 0;
  cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
  const int blocks = prop.multiProcessorCount * blocks_per_sm, iters = 20000;
  double* d;
  cudaMalloc(&d, (size_t)blocks * threads * 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0);
    probe<MODE><<<blocks, threads>>>(d, iters, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  const double inst = (double)blocks * threads * iters * 8;
  const double rate = inst / (best * 1e-3);
  // SMSP cycles per warp instruction at the nominal max clock
  const double warp_inst_per_smsp = inst / 32.0 / (prop.multiProcessorCount * 4.0);
  const double cyc = best * 1e-3 * clock_khz * 1e3 / warp_inst_per_smsp;
  printf("%-34s warps/SM %2d  %.2f T inst/s  %.2f SMSP-cycles per warp instruction (at %d MHz)\n", name,
         threads * blocks_per_sm / 32, rate / 1e12, cyc, clock_khz / 1000);
  cudaFree(d);
}

int main(int argc, char** argv) {
  if (argc > 1) {                       // round-2 probes only
    for (int bps : {2, 4}) {
      run<1>("DFMA (distinct regs)", 128, bps);
      run<5>("DFMA/DMUL alternating", 128, bps);
      run<12>("DFMA/DMUL-as-DFMA alternating", 128, bps);
      run<7>("DFMA/DMUL/DADD round-robin", 128, bps);
      run<13>("the same, all as DFMA", 128, bps);
      run<14>("kernel mix 4 DFMA:3 DMUL:1 DADD", 128, bps);
      run<15>("kernel mix, all as DFMA", 128, bps);
      run<16>("DFMA/DADD alternating", 128, bps);
      run<17>("DMUL/DADD alternating", 128, bps);
      run<18>("4 DFMA + 1 DSETP+SEL", 128, bps);
      run<19>("2 DFMA + 1 ALU", 128, bps);
      run<20>("4 DFMA + 1 SHFL", 128, bps);
    }
    return 0;
  }
  for (int bps : {1, 2, 3}) {
    run<7>("DFMA/DMUL/DADD round-robin", 128, bps);
    run<8>("DFMA + 1 ALU op each", 128, bps);
    run<9>("DFMA/DMUL alternating + ALU op", 128, bps);
    run<10>("DFMA single dependent chain", 128, bps);
    run<11>("4 DFMA then 4 DMUL", 128, bps);
  }
  for (int bps : {3}) {
    run<0>("DFMA a=a*X+Y (shared X,Y)", 128, bps);
    run<1>("DFMA a=a*x_k+y_k (distinct regs)", 128, bps);
    run<2>("DFMA a=x_k*y_k+a", 128, bps);
    run<3>("DMUL a=a*x_k", 128, bps);
    run<4>("DADD a=a+y_k", 128, bps);
    run<5>("DFMA/DMUL alternating", 128, bps);
    run<6>("DFMA operands spread", 128, bps);
  }
  return 0;
}
