// Round-2 probe: what does a non-DFMA instruction cost inside an FP64-bound instruction stream on sm_100a?
// Each loop iteration issues 8 independent DFMAs (8 chains) plus N copies of one other instruction kind, for
// N = 0, 2, 4, 8, 16.  Reported: SMSP cycles per loop iteration per warp; the slope over N is the issue cost of that
// instruction kind next to a saturated FP64 pipe (0 = it co-issues for free).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_coissue fp64_coissue.cu && ./fp64_coissue
#include <cstdio>
#include <cuda_runtime.h>

#define FMA(d, a, b, c) asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a), "d"(b), "d"(c))

enum { K_NONE, K_DSETP, K_FSEL, K_SHFL, K_LDS, K_STS, K_IMAD, K_ISETP, K_MUFU, K_DADD, K_MOV, K_XOR, K_DSETP_ONLY };

template <int KIND>
__device__ __forceinline__ void other(int j, double* a, double* x, int* n, int i, double* sm) {
  const int k = j & 7, kx = 4 + (j & 3), kx1 = 4 + ((j + 1) & 3);
  if (KIND == K_DSETP || KIND == K_DSETP_ONLY)  // DSETP + predicated integer add
    asm volatile("{.reg .pred p; setp.gt.f64 p, %1, %2; @p add.s32 %0, %0, 1;}" : "+r"(n[k]) : "d"(a[k]), "d"(x[kx1]));
  if (KIND == K_FSEL)                           // one 32-bit select on an integer predicate
    asm volatile("{.reg .pred p; setp.gt.s32 p, %1, 0; selp.b32 %0, %0, %2, p;}" : "+r"(n[k]) : "r"(n[(k + 1) & 7]), "r"(i));
  if (KIND == K_SHFL) asm volatile("shfl.sync.up.b32 %0, %0, 1, 0, 0xffffffff;" : "+r"(n[k]));
  if (KIND == K_LDS) { double v; asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"(n[k] & 0x7f8)); x[kx] = v; }
  if (KIND == K_STS) asm volatile("st.volatile.shared.f64 [%0], %1;" ::"r"((n[k] & 0x7f8)), "d"(x[kx]));
  if (KIND == K_IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(n[k]) : "r"(i), "r"(n[(k + 1) & 7]));
  if (KIND == K_ISETP) asm volatile("{.reg .pred p; setp.gt.s32 p, %1, %2; @p add.s32 %0, %0, 1;}" : "+r"(n[k]) : "r"(n[(k + 1) & 7]), "r"(i));
  if (KIND == K_MUFU) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(x[kx]));
  if (KIND == K_DADD) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x[kx]) : "d"(x[kx1]));
  if (KIND == K_MOV) asm volatile("mov.b32 %0, %1;" : "=r"(n[k]) : "r"(n[(k + 3) & 7]));
  if (KIND == K_XOR) asm volatile("xor.b32 %0, %0, %1;" : "+r"(n[k]) : "r"(i));
}

template <int KIND, int N>
__global__ void probe(double* out, int iters, double seed) {
  __shared__ double sm[256 + 32];
  double a[8], x[8], y[8];
  int n[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { a[k] = threadIdx.x * 1e-9 + k; x[k] = 1.0 + seed * (k + 1); y[k] = seed * (k + 2); n[k] = threadIdx.x * 8 + k; }
  sm[threadIdx.x] = seed; sm[threadIdx.x + 128] = seed;
  __syncthreads();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (KIND != K_DSETP_ONLY) FMA(a[k], a[k], x[k & 3], y[k]);
      // spread the N other instructions evenly between the DFMAs
#pragma unroll
      for (int j = (k * N) / 8; j < ((k + 1) * N) / 8; ++j) other<KIND>(j, a, x, n, i, sm);
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k] + n[k] + x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND, int N>
double run(int blocks_per_sm) {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  int clock_khz = 0;
  cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
  const int threads = 128, blocks = prop.multiProcessorCount * blocks_per_sm, iters = 20000;
  double* d;
  cudaMalloc(&d, (size_t)blocks * threads * 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0);
    probe<KIND, N><<<blocks, threads>>>(d, iters, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  cudaFree(d);
  // SMSP cycles per loop iteration per warp: each SMSP runs blocks_per_sm warps (128 threads = 4 warps, one per SMSP)
  return best * 1e-3 * clock_khz * 1e3 / ((double)iters * blocks_per_sm);
}

template <int KIND>
void sweep(const char* name) {
  for (int bps : {2, 4}) {
    const double c0 = run<KIND, 0>(bps), c2 = run<KIND, 2>(bps), c4 = run<KIND, 4>(bps), c8 = run<KIND, 8>(bps), c16 = run<KIND, 16>(bps);
    printf("%-28s warps/SMSP %d  cycles/iter (8 DFMA + N):  N=0 %.2f  N=2 %.2f  N=4 %.2f  N=8 %.2f  N=16 %.2f   slope(8..16) %.2f cyc/instr\n",
           name, bps, c0, c2, c4, c8, c16, (c16 - c8) / 8.0);
  }
}

int main() {
  sweep<K_DSETP>("DSETP + @P IADD");
  sweep<K_DSETP_ONLY>("DSETP + @P IADD, no DFMA");
  sweep<K_FSEL>("ISETP + SELP");
  sweep<K_ISETP>("ISETP + @P IADD");
  sweep<K_SHFL>("SHFL");
  sweep<K_LDS>("LDS.64");
  sweep<K_STS>("STS.64");
  sweep<K_IMAD>("IMAD");
  sweep<K_XOR>("LOP3");
  sweep<K_MOV>("MOV");
  sweep<K_MUFU>("MUFU.RCP64H");
  sweep<K_DADD>("DADD");
  return 0;
}
