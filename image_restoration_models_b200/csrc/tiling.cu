// Device-side tiling and Gaussian-window blending of the reference's inference harness
// (run_model_inference, /root/reference/src/utils.py:353-454): SURVEY.md §8(f) rows 1-2.
//
//   tile_gather : uint8 / uint16 / fp32 HWC image -> normalised fp32 NCHW tiles, reflect-padded on the right/bottom to
//                 a multiple of 8 (normalize :159-171, patch cut :405, HWC->CHW :412, pad :174-181,414-416); optional
//                 synthetic degradation of the patch before the pad (add_gaussian_noise :29-36, :408-409): the
//                 reference reseeds numpy with 0 for EVERY tile, so the noise field is one constant [th][tw][C] float64
//                 array per tile shape and sigma, which the host generates once; patch = clip(f32(f64(patch) + noise))
//   tile_blend  : fp32 NCHW tile predictions -> HWC image: crop :417, window-weighted accumulate :433-434 in the
//                 reference's tile order with separate fp32 multiply and add (numpy has no FMA), divide by
//                 max(wsum, 1e-8) :440, clip / round-half-even / cast :443-450
// Both reproduce the reference arithmetic bit for bit given the same tile predictions.
#include "common.cuh"

#include <algorithm>

namespace irb {

namespace {

template <typename T>
__global__ void __launch_bounds__(256) tile_gather_kernel(const T* __restrict__ img, float divisor, int H, int W, int C,
                                                          const int* __restrict__ tile_xy, int T_, int th, int tw,
                                                          int TH, int TW, const double* __restrict__ noise,
                                                          float* __restrict__ out) {
  const long long total = (long long)T_ * C * TH * TW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % TW);
    const int y = (int)((idx / TW) % TH);
    const int c = (int)((idx / ((long long)TW * TH)) % C);
    const int t = (int)(idx / ((long long)TW * TH * C));
    // F.pad(..., (0, padw, 0, padh), 'reflect'): index th + k maps to th - 2 - k
    const int sy = y < th ? y : 2 * (th - 1) - y;
    const int sx = x < tw ? x : 2 * (tw - 1) - x;
    const int h0 = tile_xy[2 * t], w0 = tile_xy[2 * t + 1];
    const float v = (float)img[((long long)(h0 + sy) * W + (w0 + sx)) * C + c];
    float o = divisor == 1.0f ? v : __fdiv_rn(v, divisor);
    if (noise != nullptr) {
      // numpy: float32 array += float64 array is computed in double and rounded back to float32; then np.clip(., 0, 1)
      o = (float)__dadd_rn((double)o, noise[((long long)sy * tw + sx) * C + c]);
      o = fminf(fmaxf(o, 0.f), 1.f);
    }
    out[idx] = o;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) tile_blend_kernel(const float* __restrict__ pred, const int* __restrict__ tile_xy,
                                                         int T_, int th, int tw, int TH, int TW,
                                                         const float* __restrict__ window, int win_ld, int H, int W,
                                                         int C, T* __restrict__ out, float scale, float lo, float hi,
                                                         int round_out) {
  const long long total = (long long)H * W * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const int x = (int)((idx / C) % W);
    const int y = (int)(idx / ((long long)C * W));
    float acc = 0.f, wsum = 0.f;
    for (int t = 0; t < T_; ++t) {            // reference loop order: h_idx outer, w_idx inner
      const int ry = y - tile_xy[2 * t], rx = x - tile_xy[2 * t + 1];
      if (ry >= 0 && ry < th && rx >= 0 && rx < tw) {
        const float wv = window[ry * win_ld + rx];
        const float pv = pred[(((long long)t * C + c) * TH + ry) * TW + rx];
        acc = __fadd_rn(acc, __fmul_rn(pv, wv));
        wsum = __fadd_rn(wsum, wv);
      }
    }
    float v = __fdiv_rn(acc, fmaxf(wsum, 1e-8f));
    v = fminf(fmaxf(__fmul_rn(v, scale), lo), hi);
    if (round_out) v = rintf(v);               // np.round: half to even
    out[idx] = (T)v;
  }
}

int nblocks(long long total) { return (int)std::max<long long>(1, std::min<long long>(cdivll(total, 256), 148LL * 16)); }

}  // namespace

int launch_tile_gather(const void* img, int dtype, float divisor, int H, int W, int C, const int* tile_xy, int T_, int th,
                       int tw, int TH, int TW, const double* noise, float* out, cudaStream_t s) {
  IRB_REQUIRE(th >= 1 && tw >= 1 && TH >= th && TW >= tw && TH - th < th && TW - tw < tw,
              "tile_gather: reflect padding must be smaller than the tile");
  const long long total = (long long)T_ * C * TH * TW;
  if (dtype == 0) tile_gather_kernel<uint8_t><<<nblocks(total), 256, 0, s>>>((const uint8_t*)img, divisor, H, W, C, tile_xy, T_, th, tw, TH, TW, noise, out);
  else if (dtype == 1) tile_gather_kernel<uint16_t><<<nblocks(total), 256, 0, s>>>((const uint16_t*)img, divisor, H, W, C, tile_xy, T_, th, tw, TH, TW, noise, out);
  else if (dtype == 2) tile_gather_kernel<float><<<nblocks(total), 256, 0, s>>>((const float*)img, divisor, H, W, C, tile_xy, T_, th, tw, TH, TW, noise, out);
  else IRB_REQUIRE(false, "tile_gather: dtype must be 0 (uint8), 1 (uint16) or 2 (float32)");
  IRB_LAUNCH_CHECK();
  return IR_OK;
}

int launch_tile_blend(const float* pred, const int* tile_xy, int T_, int th, int tw, int TH, int TW, const float* window,
                      int win_ld, int H, int W, int C, void* out, int dtype, float scale, float lo, float hi,
                      cudaStream_t s) {
  const long long total = (long long)H * W * C;
  if (dtype == 0) tile_blend_kernel<uint8_t><<<nblocks(total), 256, 0, s>>>(pred, tile_xy, T_, th, tw, TH, TW, window, win_ld, H, W, C, (uint8_t*)out, scale, lo, hi, 1);
  else if (dtype == 1) tile_blend_kernel<uint16_t><<<nblocks(total), 256, 0, s>>>(pred, tile_xy, T_, th, tw, TH, TW, window, win_ld, H, W, C, (uint16_t*)out, scale, lo, hi, 1);
  else if (dtype == 2) tile_blend_kernel<float><<<nblocks(total), 256, 0, s>>>(pred, tile_xy, T_, th, tw, TH, TW, window, win_ld, H, W, C, (float*)out, scale, lo, hi, 0);
  else IRB_REQUIRE(false, "tile_blend: dtype must be 0 (uint8), 1 (uint16) or 2 (float32)");
  IRB_LAUNCH_CHECK();
  return IR_OK;
}

}  // namespace irb
