// PSNR / SSIM of a restored image against its target on the device: SURVEY.md §8(f) row 4,
// calculate_metrics (/root/reference/src/utils.py:134-156).  The arithmetic lives in scikit-image
// (requirements.txt: scikit-image>=0.18.1, not vendored in the reference), restated from its published algorithm:
//   peak_signal_noise_ratio : both images to float64, mse = mean((a - b)^2), 10 log10(R^2 / mse)
//   structural_similarity   : defaults win_size 7, uniform window, K1 0.01, K2 0.03, sample covariance (49/48), float64
//                             for integer images; per-channel 2-D SSIM map, borders of (win-1)/2 = 3 pixels cropped, mean
//                             over the map, then the mean over the channels (channel_axis=2 for HWC colour, :149-150)
// For uint8 / uint16 images the five window sums (x, y, x^2, y^2, xy over 49 pixels) are exact integers here, so the only
// differences to scikit-image's float64 running-sum filter are at the 1e-13 level; the squared-error sum is an exact
// 64-bit integer.  Partials are reduced in a fixed order: results are bit-reproducible run to run.
#include "common.cuh"

#include <algorithm>

namespace irb {

namespace {

constexpr int TX = 32, TY = 32, R = 3, WIN = 7;
constexpr int PX = TX + 2 * R, PY = TY + 2 * R;

template <typename T> struct MetricAcc;
template <> struct MetricAcc<uint8_t> { using S = int; using A = int; using E = unsigned long long; };
template <> struct MetricAcc<uint16_t> { using S = int; using A = long long; using E = unsigned long long; };
template <> struct MetricAcc<float> { using S = float; using A = double; using E = double; };

// grid = (tiles_x, tiles_y, C); one 32 x 32 pixel tile of one channel per block, 256 threads, 4 pixels per thread
template <typename T>
__global__ void __launch_bounds__(256) metrics_tile_kernel(const T* __restrict__ pred, const T* __restrict__ target, int H,
                                                           int W, int C, double c1, double c2,
                                                           typename MetricAcc<T>::E* __restrict__ err_part,
                                                           double* __restrict__ ssim_part) {
  using S = typename MetricAcc<T>::S;
  using A = typename MetricAcc<T>::A;
  using E = typename MetricAcc<T>::E;
  __shared__ S sx[PY][PX + 1], sy[PY][PX + 1];
  __shared__ double red_s[256];
  __shared__ E red_e[256];
  const int c = blockIdx.z;
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  for (int i = threadIdx.x; i < PX * PY; i += 256) {
    const int py = i / PX, px = i % PX;
    const int gy = y0 + py - R, gx = x0 + px - R;
    S a = 0, b = 0;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
      const long long o = ((long long)gy * W + gx) * C + c;
      a = (S)target[o];            // im1 = target, im2 = pred (utils.py:144, :148-154); the formulas are symmetric
      b = (S)pred[o];
    }
    sx[py][px] = a; sy[py][px] = b;
  }
  __syncthreads();
  const double inv_np = 1.0 / (WIN * WIN), cov_norm = (double)(WIN * WIN) / (WIN * WIN - 1);
  double ssum = 0.0;
  E esum = 0;
  for (int k = 0; k < 4; ++k) {
    const int i = threadIdx.x + 256 * k;
    const int ty = i / TX, tx = i % TX;
    const int gy = y0 + ty, gx = x0 + tx;
    if (gy >= H || gx >= W) continue;
    {
      const A d = (A)sx[ty + R][tx + R] - (A)sy[ty + R][tx + R];
      esum += (E)(d * d);
    }
    if (gy < R || gy >= H - R || gx < R || gx >= W - R) continue;     // crop(S, pad)
    A s1 = 0, s2 = 0, s11 = 0, s22 = 0, s12 = 0;
#pragma unroll
    for (int dy = 0; dy < WIN; ++dy) {
#pragma unroll
      for (int dx = 0; dx < WIN; ++dx) {
        const A a = (A)sx[ty + dy][tx + dx], b = (A)sy[ty + dy][tx + dx];
        s1 += a; s2 += b; s11 += a * a; s22 += b * b; s12 += a * b;
      }
    }
    const double ux = (double)s1 * inv_np, uy = (double)s2 * inv_np;
    const double uxx = (double)s11 * inv_np, uyy = (double)s22 * inv_np, uxy = (double)s12 * inv_np;
    const double vx = cov_norm * (uxx - ux * ux), vy = cov_norm * (uyy - uy * uy), vxy = cov_norm * (uxy - ux * uy);
    const double a1 = 2.0 * ux * uy + c1, a2 = 2.0 * vxy + c2, b1 = ux * ux + uy * uy + c1, b2 = vx + vy + c2;
    ssum += (a1 * a2) / (b1 * b2);
  }
  red_s[threadIdx.x] = ssum; red_e[threadIdx.x] = esum;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) { red_s[threadIdx.x] += red_s[threadIdx.x + st]; red_e[threadIdx.x] += red_e[threadIdx.x + st]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const long long blk = ((long long)c * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    ssim_part[blk] = red_s[0];
    err_part[blk] = red_e[0];
  }
}

// one block: per-channel SSIM means, their mean, and PSNR from the summed squared error
template <typename E>
__global__ void __launch_bounds__(256) metrics_final_kernel(const E* __restrict__ err_part, const double* __restrict__ ssim_part,
                                                            int blocks_per_channel, int C, double n_elems, double area,
                                                            double data_range, double* __restrict__ out) {
  __shared__ double red_s[256];
  __shared__ E red_e[256];
  double ssim_mean = 0.0;
  E err = 0;
  for (int c = 0; c < C; ++c) {
    double s = 0.0; E e = 0;
    for (int i = threadIdx.x; i < blocks_per_channel; i += 256) {
      s += ssim_part[(long long)c * blocks_per_channel + i];
      e += err_part[(long long)c * blocks_per_channel + i];
    }
    red_s[threadIdx.x] = s; red_e[threadIdx.x] = e;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
      if (threadIdx.x < st) { red_s[threadIdx.x] += red_s[threadIdx.x + st]; red_e[threadIdx.x] += red_e[threadIdx.x + st]; }
      __syncthreads();
    }
    ssim_mean += red_s[0] / area;
    err += red_e[0];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double mse = (double)err / n_elems;
    out[0] = 10.0 * log10((data_range * data_range) / mse);       // mse == 0 -> +inf, as numpy
    out[1] = ssim_mean / C;
    out[2] = mse;
  }
}

template <typename T>
int run_metrics(const void* pred, const void* target, int H, int W, int C, double data_range, double* out, void* ws,
                cudaStream_t s) {
  using E = typename MetricAcc<T>::E;
  const dim3 grid(cdiv(W, TX), cdiv(H, TY), C);
  const long long nblk = (long long)grid.x * grid.y * grid.z;
  E* err_part = (E*)ws;
  double* ssim_part = (double*)((char*)ws + align_up((size_t)nblk * sizeof(E), 256));
  const double c1 = (0.01 * data_range) * (0.01 * data_range), c2 = (0.03 * data_range) * (0.03 * data_range);
  metrics_tile_kernel<T><<<grid, 256, 0, s>>>((const T*)pred, (const T*)target, H, W, C, c1, c2, err_part, ssim_part);
  IRB_LAUNCH_CHECK();
  metrics_final_kernel<E><<<1, 256, 0, s>>>(err_part, ssim_part, (int)(grid.x * grid.y), C, (double)H * W * C,
                                           (double)(H - 2 * R) * (W - 2 * R), data_range, out);
  IRB_LAUNCH_CHECK();
  return IR_OK;
}

}  // namespace

size_t image_metrics_workspace_bytes(int H, int W, int C) {
  if (H <= 0 || W <= 0 || C <= 0) return 0;
  const size_t nblk = (size_t)cdiv(W, TX) * cdiv(H, TY) * C;
  return 2 * align_up(nblk * 8, 256);
}

int launch_image_metrics(const void* pred, const void* target, int dtype, int H, int W, int C, double data_range,
                         double* out, void* ws, size_t ws_bytes, cudaStream_t s) {
  IRB_REQUIRE(C == 1 || C == 3, "image_metrics: HWC image with 1 or 3 channels (calculate_metrics' 2-D / channel_axis=2 cases)");
  IRB_REQUIRE(H >= WIN && W >= WIN, "image_metrics: win_size 7 exceeds the image extent (scikit-image raises ValueError)");
  IRB_REQUIRE(data_range > 0.0, "image_metrics: data_range must be positive");
  if (ws_bytes < image_metrics_workspace_bytes(H, W, C)) { set_error("workspace too small"); return IR_ERR_WORKSPACE; }
  if (dtype == 0) return run_metrics<uint8_t>(pred, target, H, W, C, data_range, out, ws, s);
  if (dtype == 1) return run_metrics<uint16_t>(pred, target, H, W, C, data_range, out, ws, s);
  if (dtype == 2) return run_metrics<float>(pred, target, H, W, C, data_range, out, ws, s);
  IRB_REQUIRE(false, "image_metrics: dtype must be 0 (uint8), 1 (uint16) or 2 (float32)");
  return IR_OK;
}

}  // namespace irb
