#!/usr/bin/env python
"""profiles/fused_gdfn_pipes.json (the `limiter_evidence` object of bench.py's roofline) from the raw page of the
`ncu --set full` capture of the fused kernels (scripts/ncu_blocks.sh).  Usage: make_gdfn_pipes.py gpurun_out/ncu_raw_<tag>.csv <tag>"""
import csv, json, os, sys

src, tag = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(src)))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
out = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    which = "ffn_fused_kernel" if "ffn_fused" in name else "attn_fused_kernel" if "attn_fused" in name else None
    if which is None:
        continue
    f = lambda k: float(r[col[k]])
    out[which] = {
        "kernel": name[:70],
        "duration_us_under_ncu": f("gpu__time_duration.sum"),
        "pipe_fmaheavy_cycles_active_pct": f("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        "pipe_fma_cycles_active_pct": f("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        "issue_slots_busy_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "pipe_tensor_cycles_active_pct": f("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        "pipe_xu_pct": f("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        "pipe_lsu_pct": f("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        "dram_read_mbytes": f("dram__bytes_read.sum"), "dram_write_mbytes": f("dram__bytes_write.sum"),
        "dram_read_pct_of_peak": f("dram__bytes_read.sum.pct_of_peak_sustained_elapsed"),
        "dram_write_pct_of_peak": f("dram__bytes_write.sum.pct_of_peak_sustained_elapsed"),
        "registers_per_thread": f("launch__registers_per_thread"),
        "dynamic_smem_kbytes": f("launch__shared_mem_per_block_dynamic"),
        "inst_executed": f("smsp__inst_executed.sum"),
        "shared_bank_conflicts": f("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        "stall_cycles_per_issue": {k.split("issue_stalled_")[1].split("_per_issue")[0]: round(float(r[i]), 3)
                                   for k, i in col.items() if k.startswith("smsp__average_warps_issue_stalled_") and float(r[i] or 0) >= 0.1},
    }
res = {"source": f"ncu --set full --clock-control none over one C = 96 TransformerBlock of scripts/bench_kernels.py --ncu "
                 f"(scripts/ncu_blocks.sh, pass {tag}; summary in profiles/r02_fused_ncu_summary.csv, details in "
                 f"profiles/{tag}_fused_ncu_details.csv)",
       **out.get("ffn_fused_kernel", {}),
       "mdta_fused_front": out.get("attn_fused_kernel"),
       "reading": "the FP32 FMA pipe (fmaheavy: FFMA2 / FMUL2 issue there only) is the busiest unit of both fused kernels, then the "
                  "issue slots; the tensor pipe is a quarter busy and DRAM a quarter of its peak: they are bound by CUDA-core "
                  "arithmetic (depthwise taps + GELU gate), not by HBM or the tensor pipe"}
json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "fused_gdfn_pipes.json"), "w"), indent=1)
print(json.dumps(res, indent=1)[:1500])
