// umma_rate.cu — development probe (not part of the library), run on a real B200:
//   (1) correctness of tcgen05.mma with the A operand in TENSOR MEMORY (written with tcgen05.st) for
//       the feature-backward shape  D[128 ch, 64 g] = V^T[128 ch, K=128 px] x W[128 px, 64 g];
//   (2) issue-to-completion cycles per batch of the MMA patterns the blend kernels use:
//       fwd  : 6 x (M=128, N=256, K=16), A and B in shared memory
//       bwdS : 8 k-steps x [ (N=64) + (N=32) ], A (V tile) and B (weights) in shared memory
//       bwdT : the same with A in tensor memory
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/build/umma_rate tools/umma_rate.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  for (long long i = 0; i < (1ll << 24); ++i) if (mbar_try(bar, parity)) return;
  __trap();
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
               ::"r"(d), "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ uint32_t swz(uint32_t off) { return off ^ (((off >> 7) & 7u) << 4); }
__device__ __forceinline__ void split2(float x, __nv_bfloat16 &h, __nv_bfloat16 &l) {
  h = __float2bfloat16_rn(x);
  l = __float2bfloat16_rn(x - __bfloat162float(h));
}
constexpr uint32_t IDESC_BASE = (1u << 4) | (1u << 7) | (1u << 10);

// mode 0: A from smem (MN-major), mode 1: A from TMEM.  reps > 1 times the MMA stream.
__global__ void __launch_bounds__(128) probe(const float *V, const float *W, float *D, int mode, int reps,
                                             long long *cycles) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char *sV1 = smem, *sV2 = smem + 32768, *sW1 = smem + 65536, *sW2 = smem + 81920;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base;
  {
    const int p = tid;
    for (int c = 0; c < 16; ++c) {
      __nv_bfloat16 h[8], l[8];
      for (int k = 0; k < 8; ++k) split2(V[p * 128 + c * 8 + k], h[k], l[k]);
      const int j = c >> 3, cc = c & 7;
      const uint32_t off = j * 16384 + (p >> 3) * 1024 + swz((p & 7) * 128 + cc * 16);
      *reinterpret_cast<uint4 *>(sV1 + off) = *reinterpret_cast<uint4 *>(h);
      *reinterpret_cast<uint4 *>(sV2 + off) = *reinterpret_cast<uint4 *>(l);
    }
    for (int c = 0; c < 8; ++c) {
      __nv_bfloat16 h[8], l[8];
      for (int k = 0; k < 8; ++k) split2(W[p * 64 + c * 8 + k], h[k], l[k]);
      const uint32_t off = (p >> 3) * 1024 + swz((p & 7) * 128 + c * 16);
      *reinterpret_cast<uint4 *>(sW1 + off) = *reinterpret_cast<uint4 *>(h);
      *reinterpret_cast<uint4 *>(sW2 + off) = *reinterpret_cast<uint4 *>(l);
    }
    // A in TMEM: thread = channel = TMEM lane; column j of a part holds pixels (2j, 2j+1)
    const int ch = tid;
    for (int j0 = 0; j0 < 64; j0 += 8) {
      uint32_t rh[8], rl[8];
      for (int j = 0; j < 8; ++j) {
        __nv_bfloat16 h0, l0, h1, l1;
        split2(V[(2 * (j0 + j)) * 128 + ch], h0, l0);
        split2(V[(2 * (j0 + j) + 1) * 128 + ch], h1, l1);
        rh[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        rl[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
      }
      tmem_st8(tb + ((uint32_t)(warp * 32) << 16) + 64 + j0, rh);
      tmem_st8(tb + ((uint32_t)(warp * 32) << 16) + 128 + j0, rl);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    const uint32_t i64 = IDESC_BASE | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t i32 = IDESC_BASE | (1u << 15) | (1u << 16) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t t64 = IDESC_BASE | (0u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t t32 = IDESC_BASE | (0u << 15) | (1u << 16) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t f256 = IDESC_BASE | (0u << 15) | (1u << 16) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (mode == 2) {
        // forward pattern: 6 MMAs N=256 (operands: any valid smem; results unused)
        for (int ks = 0; ks < 2; ++ks)
          for (int q = 0; q < 3; ++q)
            umma_ss(tb, make_desc(smem_u32(sW1) + ks * 32, 16, 1024),
                    make_desc(smem_u32(sV1) + ks * 2048, 4096, 1024), f256, 1);
        continue;
      }
      if (mode == 3 || mode == 4) {
        // transposed forward: 12 MMAs N=128; mode 3 = MN-major A (feature tile), mode 4 = K-major A
        const uint32_t id = IDESC_BASE | ((mode == 3 ? 1u : 0u) << 15) | ((128u >> 3) << 17) |
                            ((128u >> 4) << 24);
        for (int ks = 0; ks < 2; ++ks)
          for (int mb = 0; mb < 2; ++mb)
            for (int q = 0; q < 3; ++q) {
              const uint64_t a = mode == 3 ? make_desc(smem_u32(sV1) + mb * 8192 + ks * 2048, 4096, 1024)
                                           : make_desc(smem_u32(sV1) + mb * 16384 + ks * 32, 16, 1024);
              umma_ss(tb + mb * 128, a, make_desc(smem_u32(sW1) + ks * 32, 16, 1024), id, 1);
            }
        continue;
      }
      // W tile = [hi(32 g) | lo(32 g)] per pixel row: N=64 covers both halves, N=32 the hi half
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t bw = make_desc(smem_u32(sW1) + ks * 2048, 16, 1024);
        const uint32_t acc = (r > 0 || ks > 0) ? 1u : 0u;
        if (mode == 0) {
          umma_ss(tb, make_desc(smem_u32(sV1) + ks * 2048, 16384, 1024), bw, i64, acc);
          umma_ss(tb, make_desc(smem_u32(sV2) + ks * 2048, 16384, 1024), bw, i32, 1);
        } else {
          umma_ts(tb, tb + 64 + ks * 8, bw, t64, acc);
          umma_ts(tb, tb + 128 + ks * 8, bw, t32, 1);
        }
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    cycles[0] = clock64() - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (mode < 2) {
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + c0, r);
      for (int k = 0; k < 32; ++k) D[tid * 64 + c0 + k] = __uint_as_float(r[k]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "n"(256));
}

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

int main() {
  srand(1);
  // W here is "hi only" for the check: columns 0..31 of the 64 hold g 0..31, columns 32..63 g 32..63
  std::vector<float> V(128 * 128), W(128 * 64), D(128 * 64);
  for (auto &x : V) x = frand();
  for (auto &x : W) x = fabsf(frand());
  float *dV, *dW, *dD; long long *dC;
  CK(cudaMalloc(&dV, V.size() * 4)); CK(cudaMalloc(&dW, W.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMalloc(&dC, 8));
  CK(cudaMemcpy(dV, V.data(), V.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304 + 1024));
  for (int mode = 0; mode < 2; ++mode) {
    CK(cudaMemset(dD, 0, D.size() * 4));
    probe<<<1, 128, 98304 + 1024>>>(dV, dW, dD, mode, 1, dC);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: kernel error %s\n", mode, cudaGetErrorString(e)); return 1; }
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    // expected: D[c][g] = sum_p (Vhi+Vlo_if_g<32)[p][c] * Whi[p][g]  ~  V*W for g<32 (minus lo*lo terms),
    // and Vhi*Whi only for g >= 32.  Compare g < 32 against the fp64 product within bf16 tolerance.
    double maxerr = 0, maxref = 0;
    for (int c = 0; c < 128; ++c)
      for (int g = 0; g < 32; ++g) {
        double ref = 0;
        for (int p = 0; p < 128; ++p) {
          __nv_bfloat16 h = __float2bfloat16_rn(W[p * 64 + g]);
          ref += (double)V[p * 128 + c] * (double)__bfloat162float(h);
        }
        maxerr = fmax(maxerr, fabs(ref - D[c * 64 + g]));
        maxref = fmax(maxref, fabs(ref));
      }
    printf("mode %d (%s): max abs err %.3e (max |ref| %.3f, rel %.3e)\n", mode, mode ? "A in TMEM" : "A in smem",
           maxerr, maxref, maxerr / maxref);
  }
  const char *names[5] = {"bwdS (A smem)", "bwdT (A tmem)", "fwd (6 x N=256)", "fwdT A mn (12xN128)",
                          "fwdT A k (12xN128)"};
  for (int mode = 0; mode < 5; ++mode)
    for (int reps : {1, 16, 64}) {
      probe<<<1, 128, 98304 + 1024>>>(dV, dW, dD, mode, reps, dC);
      CK(cudaDeviceSynchronize());
      long long c; CK(cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost));
      printf("%-16s reps %3d: %8lld cycles  -> %.0f per batch\n", names[mode], reps, c, (double)c / reps);
    }
  printf("rate probe done\n");
  return 0;
}
