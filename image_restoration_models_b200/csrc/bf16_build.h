// Force-included (nvcc -include) into the SECOND compilation of every source: the bf16 flavour of the library.
//
// IR_MODE_BF16 is IR_MODE_HALF with bfloat16 wherever the half mode has float16: the 16-bit intermediates in HBM and every
// tensor-core operand (kind::f16 takes both formats).  Nothing in the kernels computes IN 16-bit arithmetic -- they convert
// to fp32, accumulate in fp32 and convert back -- so the whole difference is the storage type, three conversion intrinsics,
// the cvt instruction of the saturating pack and the operand-format field of the MMA instruction descriptor.  Instead of
// threading a second template parameter through thirty kernels, the sources are compiled twice: this header renames the
// types, moves everything into namespace irb_bf16 and renames the mode-taking entry points, which the primary build's
// entry points forward to (api.cu).  Shared process state (last-error string, launch profiler) stays in the primary build.
#pragma once
#include <cuda_fp16.h>
#include <cuda_bf16.h>

#define IRB_BF16_BUILD 1
#define irb irb_bf16

#define __half __nv_bfloat16
#define __half2 __nv_bfloat162
#define __half22float2 __bfloat1622float2
#define __float2half_rn __float2bfloat16_rn

// extern "C" entry points that take an IrMode (or own per-flavour state): the bf16 flavour lives beside the primary one
#define ir_restormer_packed_bytes ir_restormer_packed_bytes__bf16
#define ir_restormer_pack_weights ir_restormer_pack_weights__bf16
#define ir_restormer_workspace_bytes ir_restormer_workspace_bytes__bf16
#define ir_restormer_forward ir_restormer_forward__bf16
#define ir_restormer_graph_workspace_bytes ir_restormer_graph_workspace_bytes__bf16
#define ir_restormer_forward_graph ir_restormer_forward_graph__bf16
#define ir_dncnn_packed_bytes ir_dncnn_packed_bytes__bf16
#define ir_dncnn_pack_weights ir_dncnn_pack_weights__bf16
#define ir_dncnn_workspace_bytes ir_dncnn_workspace_bytes__bf16
#define ir_dncnn_forward ir_dncnn_forward__bf16
#define ir_dncnn_graph_workspace_bytes ir_dncnn_graph_workspace_bytes__bf16
#define ir_dncnn_forward_graph ir_dncnn_forward_graph__bf16
#define ir_block_workspace_bytes ir_block_workspace_bytes__bf16
#define ir_block_packed_bytes ir_block_packed_bytes__bf16
#define ir_block_pack_weights ir_block_pack_weights__bf16
#define ir_block_forward ir_block_forward__bf16
#define ir_graph_cache_clear ir_graph_cache_clear__bf16
#define ir_graph_cache_stats ir_graph_cache_stats__bf16
