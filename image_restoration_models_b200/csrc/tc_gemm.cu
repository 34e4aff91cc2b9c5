// tcgen05 (5th-gen tensor core) contraction for the 1x1 convolutions of the Restormer block:
//
//     y[pixel, n] = sum_k A(pixel, k) * W[n, k]      M = 128 pixels per tile, N <= 256 per CTA, fp32 accum in TMEM
//
// Replaces (paths relative to /root/reference/src/restormer/restormer.py): Attention.qkv :105 with the
// LayerNorm :37-39/:54-57 fused as a prologue, the attention output (attn@v -> project_out :127-131, folded
// into one per-image matrix) with the residual add :147, FeedForward.project_in :82 (+LayerNorm),
// FeedForward.project_out :86 with the residual add :148, reduce_chan_level{2,3} :223,228 with the channel
// concat :260,:265 as a two-source K loop, and skip_conv :240.
//
// Data movement.  Everything on this path is HBM-bound (K is 48..384), so the kernel is organised around
// streaming, not around the MMA:
//   * persistent CTAs; each owns one N-chunk of the weights, which it stages ONCE into shared memory in the
//     UMMA canonical K-major no-swizzle layout [K/4][NC][4 x tf32] (the weights are pre-packed in HBM in exactly
//     that order, tf32-rounded), and a contiguous range of 128-pixel tiles;
//   * the A tile is read with coalesced 16-byte loads, normalised (LayerNorm statistics in registers via
//     shuffles), rounded to tf32 and written to shared memory as [K/4][128][4 x tf32];
//   * one thread issues tcgen05.mma (kind::tf32, M=128, N=NC, K=8 per instruction); completion is signalled
//     with tcgen05.commit on an mbarrier;
//   * each warp drains its 32 TMEM lanes with tcgen05.ld, transposes through a private padded smem tile and
//     writes 128-byte-coalesced rows (bias / residual fused).
#include "common.cuh"
#include "tc_gemm.cuh"

namespace irb {

namespace {

constexpr int TM = 128;                 // pixels per tile == UMMA M
constexpr int NTHREADS = 128;
constexpr int STG_LD = 36;              // floats per staging row (32 + 4 pad -> conflict-free)
constexpr int STG_BYTES = 4 * 32 * STG_LD * 4;
constexpr int HDR_BYTES = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

// Spin on an mbarrier phase.  A wrong phase would hang the GPU box, so the spin is bounded and traps.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t it = 0; !done; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (it > (1u << 24)) __trap();
  }
}

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // UMMA shared-memory descriptor, SWIZZLE_NONE, K-major: LBO = byte step between the two 16-byte K-chunks of one
  // MMA, SBO = byte step between 8-row groups; version = 1 (sm_100).
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  return d;
}

__device__ __forceinline__ uint32_t make_idesc_tf32(int n) {
  // c_format F32 (bits 4-5 = 1), a/b format TF32 (= 2), both K-major, N>>3 at bit 17, M>>4 at bit 24
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
        "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
        "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct Header {           // first HDR_BYTES of dynamic shared memory
  unsigned long long bar; // mbarrier: MMA group complete
  uint32_t tmem_base;
};

// ---------------------------------------------------------------------------------------------------
// A-tile producer: rows [p0, p0+128) of image b, channels [k0, k0+kc) -> smem [kc/4][128 (+pad)][4] tf32.
// LPP lanes cooperate on one pixel row (LPP consecutive lanes), each holding VPL float4.
// ---------------------------------------------------------------------------------------------------
template <int VPL>
__device__ __forceinline__ void load_a_tile(const TcGemmParams& p, float* __restrict__ sA, int a_rows_ld,
                                            long long rowbase, int p0, int valid, int k0, int kc, int lpp, bool do_ln) {
  const int tid = threadIdx.x;
  const int q = tid % lpp;
  const int rsub = tid / lpp;
  const int pp = NTHREADS / lpp;        // pixels per pass
  const int f4n = kc >> 2;              // float4 per row in this chunk
  for (int r0 = 0; r0 < TM; r0 += pp) {
    const int r = r0 + rsub;
    const bool live = r < valid;
    float4 v[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int f = q + i * lpp;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live && f < f4n) {
        const int k = k0 + 4 * f;
        const float* src = (k < p.k1) ? p.a1 + (rowbase + p0 + r) * (long long)p.lda1 + k
                                      : p.a2 + (rowbase + p0 + r) * (long long)p.lda2 + (k - p.k1);
        v[i] = __ldg(reinterpret_cast<const float4*>(src));
      }
    }
    if (do_ln) {
      // population variance about the mean, eps inside the sqrt (restormer.py:38,55-56); two-pass in registers
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      for (int o = lpp >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mu = s / (float)kc;
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        if (q + i * lpp < f4n) {
          const float dx = v[i].x - mu, dy = v[i].y - mu, dz = v[i].z - mu, dw = v[i].w - mu;
          ss += (dx * dx + dy * dy) + (dz * dz + dw * dw);
        }
      }
      for (int o = lpp >> 1; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float rstd = 1.0f / sqrtf(ss / (float)kc + 1e-5f);
      const float sub = (p.ln_mode == LN_WITHBIAS) ? mu : 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int f = q + i * lpp;
        if (f < f4n) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(p.ln_w + 4 * f));
          float4 o4;
          o4.x = (v[i].x - sub) * rstd * g.x; o4.y = (v[i].y - sub) * rstd * g.y;
          o4.z = (v[i].z - sub) * rstd * g.z; o4.w = (v[i].w - sub) * rstd * g.w;
          if (p.ln_mode == LN_WITHBIAS) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(p.ln_b + 4 * f));
            o4.x += bb.x; o4.y += bb.y; o4.z += bb.z; o4.w += bb.w;
          }
          v[i] = o4;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int f = q + i * lpp;
      if (f < f4n) {
        float4 o4;
        o4.x = to_tf32(v[i].x); o4.y = to_tf32(v[i].y); o4.z = to_tf32(v[i].z); o4.w = to_tf32(v[i].w);
        *reinterpret_cast<float4*>(sA + ((size_t)f * a_rows_ld + r) * 4) = o4;
      }
    }
  }
}

__global__ void __launch_bounds__(NTHREADS) tc_gemm_kernel(const TcGemmParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  Header* hdr = reinterpret_cast<Header*>(smem);
  float* stg = reinterpret_cast<float*>(smem + HDR_BYTES);
  float* sA = reinterpret_cast<float*>(smem + HDR_BYTES + STG_BYTES);
  const int a_rows_ld = TM + p.a_pad;                       // rows per 16-byte K-chunk slab of A (pad breaks conflicts)
  float* sW = sA + (size_t)(p.KC >> 2) * a_rows_ld * 4;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.y * p.NC;
  const int nc = min(p.NC, p.N - n0);
  const uint32_t bar = smem_u32(&hdr->bar);

  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&hdr->tmem_base)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_base;

  const long long t_begin = (long long)p.ntiles * blockIdx.x / gridDim.x;
  const long long t_end = (long long)p.ntiles * (blockIdx.x + 1) / gridDim.x;
  const uint32_t idesc = make_idesc_tf32(nc);
  const uint32_t a_lbo = (uint32_t)a_rows_ld * 16u;
  const uint32_t w_lbo = (uint32_t)nc * 16u;
  const int nchunks = (p.K + p.KC - 1) / p.KC;
  const bool do_ln = p.ln_mode != LN_NONE;
  uint32_t phase = 0;
  int loaded_b = -1;

  for (long long tile = t_begin; tile < t_end; ++tile) {
    const int b = (int)(tile / p.tiles_per_img);
    const int p0 = (int)(tile - (long long)b * p.tiles_per_img) * TM;
    const int valid = min(TM, p.HW - p0);
    const long long rowbase = (long long)b * p.HW;

    if (loaded_b < 0 || (p.w_bstride != 0 && b != loaded_b)) {
      // stage this CTA's weight chunk: global [K/4][N][4] (+ image stride) -> smem [K/4][nc][4]
      const float4* wg = reinterpret_cast<const float4*>(p.w + (long long)b * p.w_bstride);
      float4* ws = reinterpret_cast<float4*>(sW);
      const int total = (p.K >> 2) * nc;
      for (int idx = tid; idx < total; idx += NTHREADS) {
        const int kq = idx / nc, n = idx - kq * nc;
        ws[idx] = __ldg(wg + (size_t)kq * p.N + n0 + n);
      }
      loaded_b = b;
    }

    for (int ch = 0; ch < nchunks; ++ch) {
      const int k0 = ch * p.KC;
      const int kc = min(p.KC, p.K - k0);
      if (ch > 0) {                       // previous chunk's MMAs must have finished reading sA
        mbar_wait(bar, phase);
        phase ^= 1;
      }
      if (p.vpl <= 3) load_a_tile<3>(p, sA, a_rows_ld, rowbase, p0, valid, k0, kc, p.lpp, do_ln);
      else            load_a_tile<4>(p, sA, a_rows_ld, rowbase, p0, valid, k0, kc, p.lpp, do_ln);
      fence_async_smem();                 // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sA);
        const uint32_t w_addr = smem_u32(sW) + (uint32_t)(k0 >> 2) * w_lbo;
        for (int ks = 0; ks < (kc >> 3); ++ks) {
          const uint64_t adesc = make_smem_desc(a_addr + (uint32_t)(2 * ks) * a_lbo, a_lbo, 128);
          const uint64_t bdesc = make_smem_desc(w_addr + (uint32_t)(2 * ks) * w_lbo, w_lbo, 128);
          umma_tf32(tmem_base, adesc, bdesc, idesc, (ch > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(bar);
      }
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();

    // epilogue: warp w owns TMEM lanes [32w, 32w+32) == tile rows
    float* mystg = stg + warp * 32 * STG_LD;
    for (int c0 = 0; c0 < nc; c0 += 32) {
      const int ncols = min(32, nc - c0);     // 32 or 16
      float v[32];
      __syncwarp();                           // tcgen05.ld is .sync.aligned: the warp must be converged
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      if (ncols == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j * 4 < ncols)
          *reinterpret_cast<float4*>(mystg + lane * STG_LD + j * 4) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      __syncwarp();
      const int cpr = ncols >> 2;             // float4 per row
      const int rpi = 32 / cpr;               // rows per iteration
      for (int it = 0; it < cpr; ++it) {
        const int row = it * rpi + lane / cpr;
        const int c4 = (lane % cpr) * 4;
        const int prow = warp * 32 + row;
        if (prow < valid) {
          float4 o = *reinterpret_cast<const float4*>(mystg + row * STG_LD + c4);
          const int n = n0 + c0 + c4;
          if (p.bias) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + n));
            o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
          }
          const long long grow = rowbase + p0 + prow;
          if (p.r) {
            const float4 rr = *reinterpret_cast<const float4*>(p.r + grow * p.ldr + n);
            o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
          }
          *reinterpret_cast<float4*>(p.y + grow * p.ldy + n) = o;
        }
      }
      __syncwarp();
    }
    tc_fence_before();      // TMEM reads of this tile are ordered before the next tile's MMA (after the next bar.sync)
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

}  // namespace

static int next_pow2_cols(int n) { int c = 32; while (c < n) c <<= 1; return c; }

// Choose the N-chunk, K-chunk and lane mapping; returns dynamic shared memory bytes (0 if unsupported).
size_t tc_gemm_configure(TcGemmParams& p) {
  if (p.K % 8 != 0 || p.N % 16 != 0 || p.K <= 0 || p.N <= 0) return 0;
  if (p.k1 % 4 != 0 || (p.k2 != 0 && p.k2 % 4 != 0) || p.k1 + p.k2 != p.K) return 0;
  p.KC = p.K <= 128 ? p.K : 64;
  if (p.ln_mode != LN_NONE && (p.KC != p.K || p.k2 != 0)) return 0;    // LayerNorm needs the whole row in registers
  const int f4 = p.KC / 4;
  int lpp = 1;
  while (lpp < 32 && (f4 + lpp - 1) / lpp > 3) lpp <<= 1;              // <= 3 float4 per lane when possible
  if ((f4 + lpp - 1) / lpp > 4) return 0;
  p.lpp = lpp;
  p.vpl = (f4 + lpp - 1) / lpp;
  const size_t fixed = HDR_BYTES + STG_BYTES + (size_t)(p.KC / 4) * (TM + p.a_pad) * 16;
  const size_t budget2 = 113 * 1024, budget1 = 227 * 1024;
  // fewest N-chunks whose weights fit next to the A tile; two resident CTAs per SM are preferred unless that
  // more than doubles the number of chunks (every chunk re-reads the A tile through L2)
  int nc_for[2] = {0, 0}, chunks_for[2] = {0, 0};
  for (int pass = 0; pass < 2; ++pass) {
    const size_t budget = pass == 0 ? budget2 : budget1;
    for (int chunks = 1; chunks <= p.N / 16; ++chunks) {
      int nc = ((p.N + chunks - 1) / chunks + 15) / 16 * 16;
      if (nc > 256) continue;
      if (fixed + (size_t)nc * p.K * 4 <= budget) { nc_for[pass] = nc; chunks_for[pass] = chunks; break; }
    }
  }
  int best = 0;
  if (nc_for[0] && (!nc_for[1] || chunks_for[0] <= 2 * chunks_for[1])) best = nc_for[0];
  else best = nc_for[1];
  if (!best) return 0;
  p.NC = best;
  p.tmem_cols = next_pow2_cols(best);
  return fixed + (size_t)best * p.K * 4;
}

int launch_gemm_tc(TcGemmParams p, cudaStream_t s) {
  size_t smem = tc_gemm_configure(p);
  IRB_REQUIRE(smem != 0, "tc_gemm: unsupported shape");
  IRB_REQUIRE(p.lda1 % 4 == 0 && (p.k2 == 0 || p.lda2 % 4 == 0) && p.ldy % 4 == 0 && (p.r == nullptr || p.ldr % 4 == 0),
              "tc_gemm: leading dimensions must be multiples of 4");
  p.tiles_per_img = cdiv(p.HW, TM);
  p.ntiles = p.tiles_per_img * p.B;
  const int nchunks_n = cdiv(p.N, p.NC);
  // occupancy: TMEM columns (512 per SM) and shared memory; pad smem so the hardware cannot over-subscribe TMEM
  int occ_tmem = 512 / p.tmem_cols;
  int occ = (int)std::min<size_t>((size_t)occ_tmem, (228 * 1024) / (smem + 1024));
  if (occ < 1) occ = 1;
  if (occ > 4) occ = 4;
  const size_t min_smem = (228 * 1024) / (occ + 1) + 1;    // more than occ CTAs can no longer fit
  if (smem < min_smem) smem = std::min<size_t>(min_smem, 227 * 1024);
  static size_t configured = 0;
  if (smem > configured) {
    IRB_CUDA(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = 227 * 1024;
  }
  int gx = std::max(1, (148 * occ) / nchunks_n);
  if (gx > p.ntiles) gx = p.ntiles;
  dim3 grid(gx, nchunks_n);
  const double rows = (double)p.B * p.HW;
  ProfScope prof(p.tag, 4.0 * rows * (p.K + p.N * (p.r ? 2.0 : 1.0)), 2.0 * rows * p.N * p.K, s);
  tc_gemm_kernel<<<grid, NTHREADS, smem, s>>>(p);
  IRB_LAUNCH_CHECK();
  return IR_OK;
}

}  // namespace irb
