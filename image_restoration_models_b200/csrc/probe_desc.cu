// Hardware probe (bring-up tool, not on the product path): does a SWIZZLE_128B K-major tcgen05 operand descriptor whose
// start address is shifted by whole 128-byte rows inside a TMA-written box address the rows the same way TMA wrote
// them, and what must the descriptor's base-offset field hold?  One CTA: TMA-load A[160][32] and W[32][32] (fp32,
// SWIZZLE_128B boxes), D[128][32] = A[shift : shift + 128] . W^T with kind::tf32, store D.
#ifdef IRB200_TESTING
#include "common.cuh"
#include "sm100.cuh"
#include "tmap.cuh"

namespace irb {
namespace {
using namespace sm100;

__global__ void __launch_bounds__(128, 1)
probe_desc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, float* d, int shift,
                  int base_off) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* g = smem_raw + (base - smem_u32(smem_raw));
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(g);          // [0] load, [1] mma
  uint32_t* tm = reinterpret_cast<uint32_t*>(g + 64);
  const uint32_t sA = base + 1024, sW = base + 1024 + 160 * 128 + 512;         // W box 1024-aligned: 1024 + 20480 + 512
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(smem_u32(&bar[0]), 1);
    mbar_init(smem_u32(&bar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tm)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tm;
  if (tid == 0) {
    mbar_expect_tx(smem_u32(&bar[0]), 160 * 128 + 32 * 128);
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(sA), "l"(reinterpret_cast<uint64_t>(&tmA)), "r"(smem_u32(&bar[0])), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(sW), "l"(reinterpret_cast<uint64_t>(&tmW)), "r"(smem_u32(&bar[0])), "r"(0), "r"(0) : "memory");
    mbar_wait(smem_u32(&bar[0]), 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc<float>(32);
    for (int kk = 0; kk < 4; ++kk) {
      uint64_t ad = sw128_desc(sA + (uint32_t)shift * 128u + kk * 32);
      ad |= (uint64_t)(base_off & 7) << 49;
      umma<float>(tmem, ad, sw128_desc(sW + kk * 32), idesc, kk > 0 ? 1u : 0u);
    }
    umma_commit(smem_u32(&bar[1]));
  }
  mbar_wait(smem_u32(&bar[1]), 0);
  tc_fence_after();
  float v[32];
  tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  for (int e = 0; e < 32; ++e) d[(size_t)tid * 32 + e] = v[e];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}
}  // namespace

int probe_shifted_descriptor(const float* a, const float* w, float* d, int shift, int base_off, cudaStream_t s) {
  CUtensorMap tA, tW;
  cuuint64_t da[2] = {32, 160}, dw[2] = {32, 32}, st[1] = {128};
  cuuint32_t ba[2] = {32, 160}, bw[2] = {32, 32};
  IRB_TRY(make_tmap(&tA, a, false, 2, da, st, ba, true));
  IRB_TRY(make_tmap(&tW, w, false, 2, dw, st, bw, true));
  static SmemOptIn optin;
  IRB_TRY(opt_in_smem(probe_desc_kernel, optin, 64 * 1024));
  probe_desc_kernel<<<1, 128, 40 * 1024, s>>>(tA, tW, d, shift, base_off);
  IRB_LAUNCH_CHECK();
  return IR_OK;
}
}  // namespace irb

#endif  // IRB200_TESTING
