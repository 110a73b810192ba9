// umma_probe.cu — development probe (not part of the library): checks, on a real B200, the
// tcgen05 conventions the blend kernels rely on:
//   * instruction descriptor for kind::f16 (bf16 x bf16 -> f32), M=128, A K-major / MN-major,
//     B MN-major;
//   * shared-memory matrix descriptors for SWIZZLE_128B canonical layouts, K-major and MN-major,
//     including which of LBO / SBO strides which dimension;
//   * the store-address formulas used to fill those layouts from registers;
//   * TMEM alloc / tcgen05.commit -> mbarrier / tcgen05.ld 32x32b.
// Test 1 = forward shape:  D[128 px, 256 ch] = [b1|b2][128, 64] x {f1,f2}[32, 256]  (3 products)
// Test 2 = backward shape: D[128 ch, 64 g]   = V^T[128 ch, K=128 px] x W[128 px, 64 g]
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/build/umma_probe tools/umma_probe.cu
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
  for (long long i = 0; i < (1ll << 26); ++i) if (mbar_try(bar, parity)) return;
  printf("mbar_wait timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
  __trap();
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
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
__device__ __forceinline__ uint32_t swz(uint32_t off) { return off ^ (((off >> 7) & 7u) << 4); }

// split x into bf16 hi (round-to-nearest) and bf16 lo = rn(x - hi)
__device__ __forceinline__ void split2(float x, __nv_bfloat16 &h, __nv_bfloat16 &l) {
  h = __float2bfloat16_rn(x);
  l = __float2bfloat16_rn(x - __bfloat162float(h));
}

constexpr uint32_t IDESC_BASE = (1u << 4) | (1u << 7) | (1u << 10);   // f32 accum, bf16 A, bf16 B

// ------------------------------------------------------------------------------------------------
// Test 1.  W[128][32] fp32, F[32][256] fp32 -> D[128][256]
// variant bit0: swap LBO/SBO of the MN-major B descriptor
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) probe_fwd(const float *W, const float *F, float *D, int variant) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char *sA = smem;                 // 128 rows x 128 B
  unsigned char *sF1 = smem + 16384;        // [4 n-atoms][4 k-groups][8 rows][128 B]
  unsigned char *sF2 = smem + 32768;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  // A: thread = pixel row
  {
    const int r = tid;
    for (int c = 0; c < 8; ++c) {
      __nv_bfloat16 v[8];
      for (int k = 0; k < 8; ++k) {
        const int g = (c & 3) * 8 + k;
        __nv_bfloat16 h, l;
        split2(W[r * 32 + g], h, l);
        v[k] = (c < 4) ? h : l;
      }
      *reinterpret_cast<uint4 *>(sA + swz(r * 128 + c * 16)) = *reinterpret_cast<uint4 *>(v);
    }
  }
  // B: warp per Gaussian row; lane -> 8 channels
  for (int g = warp; g < 32; g += 4) {
    const int lane = tid & 31;
    const int n0 = lane * 8, j = n0 >> 6, c = (n0 & 63) >> 3;
    __nv_bfloat16 h[8], l[8];
    for (int k = 0; k < 8; ++k) split2(F[g * 256 + n0 + k], h[k], l[k]);
    const uint32_t off = j * 4096 + (g >> 3) * 1024 + swz((g & 7) * 128 + c * 16);
    *reinterpret_cast<uint4 *>(sF1 + off) = *reinterpret_cast<uint4 *>(h);
    *reinterpret_cast<uint4 *>(sF2 + off) = *reinterpret_cast<uint4 *>(l);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = IDESC_BASE | (0u << 15) | (1u << 16) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t lbo = (variant & 1) ? 1024 : 4096, sbo = (variant & 1) ? 4096 : 1024;
    uint32_t acc = 0;
    for (int ks = 0; ks < 2; ++ks) {
      const uint64_t a1 = make_desc(smem_u32(sA) + ks * 32, 16, 1024);
      const uint64_t a2 = make_desc(smem_u32(sA) + 64 + ks * 32, 16, 1024);
      const uint64_t b1 = make_desc(smem_u32(sF1) + ks * 2048, lbo, sbo);
      const uint64_t b2 = make_desc(smem_u32(sF2) + ks * 2048, lbo, sbo);
      umma_bf16(tb, a1, b1, idesc, acc); acc = 1;
      umma_bf16(tb, a1, b2, idesc, 1);
      umma_bf16(tb, a2, b1, idesc, 1);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < 256; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + c0, r);
    for (int k = 0; k < 32; ++k) D[tid * 256 + c0 + k] = __uint_as_float(r[k]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "n"(256));
}

// ------------------------------------------------------------------------------------------------
// Test 2.  V[128 px][128 ch] fp32, W[128 px][64 g] fp32 -> D[128 ch][64 g] = V^T W
// A = V^T: MN-major (M = ch contiguous), atoms of 64 ch x 8 px; B = W: MN-major (N = g contiguous).
// variant bit0: swap LBO/SBO on A.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) probe_bwd(const float *V, const float *W, float *D, int variant) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char *sV1 = smem;               // [2 m-atoms][16 k-groups][8 rows][128 B] = 32 KB
  unsigned char *sV2 = smem + 32768;
  unsigned char *sW1 = smem + 65536;       // [16 k-groups][8 rows][128 B] = 16 KB
  unsigned char *sW2 = smem + 81920;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  {
    const int p = tid;   // thread = pixel = k row
    for (int c = 0; c < 16; ++c) {          // 16 chunks of 8 channels
      __nv_bfloat16 h[8], l[8];
      for (int k = 0; k < 8; ++k) split2(V[p * 128 + c * 8 + k], h[k], l[k]);
      const int j = c >> 3, cc = c & 7;
      const uint32_t off = j * 16384 + (p >> 3) * 1024 + swz((p & 7) * 128 + cc * 16);
      *reinterpret_cast<uint4 *>(sV1 + off) = *reinterpret_cast<uint4 *>(h);
      *reinterpret_cast<uint4 *>(sV2 + off) = *reinterpret_cast<uint4 *>(l);
    }
    for (int c = 0; c < 8; ++c) {           // 8 chunks of 8 gaussians
      __nv_bfloat16 h[8], l[8];
      for (int k = 0; k < 8; ++k) split2(W[p * 64 + c * 8 + k], h[k], l[k]);
      const uint32_t off = (p >> 3) * 1024 + swz((p & 7) * 128 + c * 16);
      *reinterpret_cast<uint4 *>(sW1 + off) = *reinterpret_cast<uint4 *>(h);
      *reinterpret_cast<uint4 *>(sW2 + off) = *reinterpret_cast<uint4 *>(l);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = IDESC_BASE | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t albo = (variant & 1) ? 1024 : 16384, asbo = (variant & 1) ? 16384 : 1024;
    uint32_t acc = 0;
    for (int ks = 0; ks < 8; ++ks) {       // K = 128 px, 16 per MMA = 2 k-groups
      const uint64_t a1 = make_desc(smem_u32(sV1) + ks * 2048, albo, asbo);
      const uint64_t a2 = make_desc(smem_u32(sV2) + ks * 2048, albo, asbo);
      const uint64_t b1 = make_desc(smem_u32(sW1) + ks * 2048, 16, 1024);
      const uint64_t b2 = make_desc(smem_u32(sW2) + ks * 2048, 16, 1024);
      umma_bf16(tb, a1, b1, idesc, acc); acc = 1;
      umma_bf16(tb, a1, b2, idesc, 1);
      umma_bf16(tb, a2, b1, idesc, 1);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + c0, r);
    for (int k = 0; k < 32; ++k) D[tid * 64 + c0 + k] = __uint_as_float(r[k]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "n"(64));
}

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

int main() {
  srand(1);
  // ---- test 1
  std::vector<float> W(128 * 32), F(32 * 256), D(128 * 256);
  for (auto &x : W) x = fabsf(frand());
  for (auto &x : F) x = frand();
  float *dW, *dF, *dD;
  CK(cudaMalloc(&dW, W.size() * 4)); CK(cudaMalloc(&dF, F.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dF, F.data(), F.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(probe_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 1024));
  CK(cudaFuncSetAttribute(probe_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304 + 1024));
  for (int variant = 0; variant < 2; ++variant) {
    CK(cudaMemset(dD, 0, D.size() * 4));
    probe_fwd<<<1, 128, 49152 + 1024>>>(dW, dF, dD, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("test1 variant %d: kernel error %s\n", variant, cudaGetErrorString(e)); return 1; }
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (int i = 0; i < 128; ++i)
      for (int j = 0; j < 256; ++j) {
        double ref = 0;
        for (int k = 0; k < 32; ++k) ref += (double)W[i * 32 + k] * F[k * 256 + j];
        maxerr = fmax(maxerr, fabs(ref - D[i * 256 + j]));
        maxref = fmax(maxref, fabs(ref));
      }
    printf("test1 (fwd shape) variant %d: max abs err %.3e  (max |ref| %.3f, rel %.3e)\n", variant, maxerr, maxref, maxerr / maxref);
  }
  // ---- test 2
  std::vector<float> V(128 * 128), W2(128 * 64), D2(128 * 64);
  for (auto &x : V) x = frand();
  for (auto &x : W2) x = fabsf(frand());
  float *dV, *dW2, *dD2;
  CK(cudaMalloc(&dV, V.size() * 4)); CK(cudaMalloc(&dW2, W2.size() * 4)); CK(cudaMalloc(&dD2, D2.size() * 4));
  CK(cudaMemcpy(dV, V.data(), V.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW2, W2.data(), W2.size() * 4, cudaMemcpyHostToDevice));
  for (int variant = 0; variant < 2; ++variant) {
    CK(cudaMemset(dD2, 0, D2.size() * 4));
    probe_bwd<<<1, 128, 98304 + 1024>>>(dV, dW2, dD2, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("test2 variant %d: kernel error %s\n", variant, cudaGetErrorString(e)); return 1; }
    CK(cudaMemcpy(D2.data(), dD2, D2.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (int c = 0; c < 128; ++c)
      for (int g = 0; g < 64; ++g) {
        double ref = 0;
        for (int p = 0; p < 128; ++p) ref += (double)V[p * 128 + c] * W2[p * 64 + g];
        maxerr = fmax(maxerr, fabs(ref - D2[c * 64 + g]));
        maxref = fmax(maxref, fabs(ref));
      }
    printf("test2 (bwd shape) variant %d: max abs err %.3e  (max |ref| %.3f, rel %.3e)\n", variant, maxerr, maxref, maxerr / maxref);
  }
  printf("probe done\n");
  return 0;
}
