// Depthwise 3x3 taps straight out of tensor memory.
//
// The fused kernels compute their 1x1 contraction TRANSPOSED: D^T[channel][patch pixel] = W . xn_patch^T, i.e. the weights
// are the MMA's A operand (M = 128 channels = 128 TMEM lanes) and the (8+2) x (16+2) halo patch of the activation is the B
// operand (N = 192 >= 180 patch pixels = TMEM columns, fp32).  A thread of a depthwise warp then OWNS one channel (its TMEM
// lane) and reads the patch pixels it needs with tcgen05.ld -- fp32, in registers, already in the order the packed FFMA2
// taps want them -- with no shared-memory round trip, no fp16 rounding of the conv input and no conversion instruction.
// Two warps share a lane quarter: warp `h` computes output columns 8h .. 8h+7 of all 8 output rows of the tile.
//
// Per patch row the thread loads E = columns 8h .. 8h+9 (pairs E[j] = columns 2j, 2j+1) and O = columns 8h+1 .. 8h+8
// (pairs O[j] = columns 2j+1, 2j+2): with out pair j = output columns (2j, 2j+1) the three taps of a kernel row are
//     w[ky][0] * E[j]  +  w[ky][1] * O[j]  +  w[ky][2] * E[j+1]
// (packed fp32 pairs, the weight broadcast to both halves), so every operand is an aligned 64-bit register pair.
#pragma once
#include <stdint.h>

namespace irb {
namespace dwt {

typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t pack2(float lo, float hi) { f2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f2_t pack2u(uint32_t lo, uint32_t hi) { f2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ void unpack2(f2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2_t fma2(f2_t a, f2_t b, f2_t c) { f2_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2_t mul2(f2_t a, f2_t b) { f2_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2_t add2(f2_t a, f2_t b) { f2_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

constexpr int PATCH_W = 18;       // patch row pitch in TMEM columns (16 + 2 halo)
constexpr int PATCH_ROWS = 10;    // 8 + 2 halo
constexpr int OUT_ROWS = 8;

struct Row {
  uint32_t e[10];
  uint32_t o[8];
};

// issue the three TMEM loads of one patch row (asynchronous: the registers are valid after wait_row)
__device__ __forceinline__ void ld_row(uint32_t taddr, Row& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r.e[0]), "=r"(r.e[1]), "=r"(r.e[2]), "=r"(r.e[3]), "=r"(r.e[4]), "=r"(r.e[5]), "=r"(r.e[6]), "=r"(r.e[7])
               : "r"(taddr));
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r.e[8]), "=r"(r.e[9]) : "r"(taddr + 8u));
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r.o[0]), "=r"(r.o[1]), "=r"(r.o[2]), "=r"(r.o[3]), "=r"(r.o[4]), "=r"(r.o[5]), "=r"(r.o[6]), "=r"(r.o[7])
               : "r"(taddr + 1u));
}
// wait for every outstanding tcgen05.ld of this thread; the registers are operands so that no use can move above it
__device__ __forceinline__ void wait_row(Row& r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r.e[0]), "+r"(r.e[1]), "+r"(r.e[2]), "+r"(r.e[3]), "+r"(r.e[4]), "+r"(r.e[5]), "+r"(r.e[6]), "+r"(r.e[7]),
                 "+r"(r.e[8]), "+r"(r.e[9]), "+r"(r.o[0]), "+r"(r.o[1]), "+r"(r.o[2]), "+r"(r.o[3]), "+r"(r.o[4]), "+r"(r.o[5]),
                 "+r"(r.o[6]), "+r"(r.o[7])
               :
               : "memory");
}

// taps of patch row `R` (compile time) into the live output rows; acc[oy % 3] is output row oy
template <int R>
__device__ __forceinline__ void row_taps(const Row& r, const f2_t (&w)[9], f2_t (&acc)[3][4]) {
  f2_t E[5], O[4];
#pragma unroll
  for (int j = 0; j < 5; ++j) E[j] = pack2u(r.e[2 * j], r.e[2 * j + 1]);
#pragma unroll
  for (int j = 0; j < 4; ++j) O[j] = pack2u(r.o[2 * j], r.o[2 * j + 1]);
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int oy = R - ky;
    if (oy < 0 || oy >= OUT_ROWS) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      f2_t a = ky == 0 ? mul2(w[0], E[j]) : fma2(w[ky * 3], E[j], acc[oy % 3][j]);
      a = fma2(w[ky * 3 + 1], O[j], a);
      acc[oy % 3][j] = fma2(w[ky * 3 + 2], E[j + 1], a);
    }
  }
}

// One unit: this thread's channel, output columns 8h .. 8h+7, output rows 0 .. 7.
//   taddr : TMEM address of patch pixel (row 0, column 8h) of this warp's lane quarter in the accumulator to read
//   w     : the channel's nine taps, each broadcast to both halves of a pair
//   emit  : emit(oy, acc) with acc[j] = output pixels (8h + 2j, 8h + 2j + 1) of row oy, called for oy = 0 .. 7 in order
// The loads of patch row R+1 are in flight while row R's taps execute.
template <typename Emit>
__device__ __forceinline__ void unit(uint32_t taddr_in, const f2_t (&w)[9], Emit&& emit) {
  // tcgen05.ld takes its address from a UNIFORM register.  The address is warp-uniform but derived from the warp index, so
  // ptxas keeps it in a vector register and emits one R2UR per load (two per LDTM in the first build: 8 % of the depthwise
  // instruction stream).  A warp reduction returns its result in a uniform register; OR of equal values is the value.
  const uint32_t taddr = __reduce_or_sync(0xffffffffu, taddr_in);
  f2_t acc[3][4];
  Row ra, rb;
  ld_row(taddr, ra);
#define IRB_DWT_STEP(R, CUR, NXT)                                              \
  wait_row(CUR);                                                               \
  if (R + 1 < PATCH_ROWS) ld_row(taddr + (uint32_t)((R + 1) * PATCH_W), NXT);  \
  row_taps<R>(CUR, w, acc);                                                    \
  if (R >= 2) emit(R - 2, acc[(R - 2) % 3]);
  IRB_DWT_STEP(0, ra, rb)
  IRB_DWT_STEP(1, rb, ra)
  IRB_DWT_STEP(2, ra, rb)
  IRB_DWT_STEP(3, rb, ra)
  IRB_DWT_STEP(4, ra, rb)
  IRB_DWT_STEP(5, rb, ra)
  IRB_DWT_STEP(6, ra, rb)
  IRB_DWT_STEP(7, rb, ra)
  IRB_DWT_STEP(8, ra, rb)
  IRB_DWT_STEP(9, rb, ra)
#undef IRB_DWT_STEP
}

}  // namespace dwt
}  // namespace irb
