#!/usr/bin/env python
"""Copy the evidence of one GPU pass (scripts/r02_gpu_final.sh <tag>) from gpurun_out/ into profiles/ and fill the
measured numbers into DESIGN.md (from scripts/DESIGN.md.tmpl).  Usage: collect_profiles.py <tag> [<bench tag>]
(<bench tag>: a later short refresh, scripts/r02_gpu_k2.sh, whose bench / parity records supersede those of <tag>)"""
import json, os, shutil, subprocess, sys

tag = sys.argv[1]
btag = sys.argv[2] if len(sys.argv) > 2 else tag
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(R, "gpurun_out"), os.path.join(R, "profiles")


def cp(src, dst):
    s = os.path.join(G, src)
    if os.path.exists(s):
        shutil.copy(s, os.path.join(P, dst))
    else:
        print("missing", src)


cp(f"bench_{tag}.json", f"{tag}_bench_fp32.json")
cp(f"bench_half_{tag}.json", f"{tag}_bench_half.json")
cp(f"bench_ref_{tag}.json", f"{tag}_bench_reference_arm.json")
cp(f"parity_{tag}.json", "r02_parity.json")
cp(f"latency_{tag}.json", f"{tag}_latency.json")
cp(f"configs_{tag}.json", f"{tag}_configs.json")
cp(f"blocks_{tag}.log", f"{tag}_block_kernels.log")
cp(f"status_{tag}.txt", f"{tag}_status.txt")
cp(f"gpu_{tag}.txt", f"{tag}_gpu.txt")
cp(f"launches_{tag}_fp32.csv", f"{tag}_ncu_launches_fp32.csv")
cp(f"traffic_{tag}_fp32.json", "traffic_fp32.json")
cp(f"launch_fp32_{tag}.csv", f"{tag}_launches_fp32_events.csv")
cp(f"ncu_summary_{tag}.csv", "r02_fused_ncu_summary.csv")
cp(f"ncu_details_{tag}.csv", f"{tag}_fused_ncu_details.csv")
if os.path.exists(os.path.join(G, f"pytest_{tag}.log")):
    open(os.path.join(P, f"{tag}_pytest_gpu.txt"), "w").write(open(os.path.join(G, f"pytest_{tag}.log")).read()[-600:])

if btag != tag:
    cp(f"bench_{btag}.json", f"{btag}_bench_fp32.json")
    cp(f"bench_half_{btag}.json", f"{btag}_bench_half.json")
    cp(f"parity_{btag}.json", "r02_parity.json")
    cp(f"launch_fp32_{btag}.csv", f"{btag}_launches_fp32_events.csv")
    cp(f"blocks_{btag}.log", f"{btag}_block_kernels.log")
d = json.load(open(os.path.join(G, f"bench_{btag}.json")))
lat = json.load(open(os.path.join(G, f"latency_{tag}.json")))
fam = []
peaks = {"fp32_pipe": "FP32 pipe", "hbm": "HBM"}
for k in d["kernels"]:
    name, ms = k["name"], k["ms"]
    if name == d["roofline"]["kernel"]:
        fam.append(f"{name} {ms:.1f} ms ({d['roofline']['frac']:.2f} of the FP32 pipe, {k['hbm_frac']:.2f} HBM)")
    else:
        which = "HBM" if k["hbm_frac"] >= k["tensor_frac"] else "tensor"
        fam.append(f"{name} {ms:.1f} ms ({k['best_frac']:.2f} {which})")
t = d["tiled_config4"]
par = json.load(open(os.path.join(G, f"parity_{btag}.json")))
e = lambda k: f"{par[k]['max_abs']:.1e}" if k in par else "n/a"
p512 = lambda m: " / ".join(e(f"size512_{t}[{m}]") for t in ("gray_denoise", "motion_deblur", "defocus_dual"))
b8 = max((v["max_abs"] for k, v in par.items() if k.startswith("config2_batch8_element") and "[fp32]" in k), default=float("nan"))
gold = lambda m: max(v["max_abs"] for k, v in par.items() if k.startswith("restormer_") and k.endswith(f"[{m}]"))
sub = {
    "@GOLDF@": f"{gold('fp32'):.1e}", "@GOLDH@": f"{gold('half'):.1e}",
    "@P512F@": p512("fp32"), "@P512H@": p512("half"), "@PB8@": f"≤ {b8:.1e}",
    "@FP32@": f"{d['value']:.1f}", "@FP32MS@": f"{d['ms_per_step']:.1f}", "@HALF@": f"{d['other_mode']['value']:.1f}",
    "@BF16@": f"{d['bf16_mode']['value']:.1f}", "@E2E@": f"{d['e2e']['value']:.1f}",
    "@TILED@": f"{t['value']:.1f}", "@TILEDC@": f"{t['computed_tile_mpix_per_s']:.1f}",
    "@LAT256@": f"{lat['restormer_color_denoise_1x3x256x256']['graph']['device_ms']:.2f}",
    "@LAT512@": f"{lat['restormer_color_denoise_1x3x512x512']['graph']['device_ms']:.2f}",
    "@FAMILIES@": "; ".join(fam) + f".  Step: {d['step_algorithmic_GB']:.0f} GB algorithmic, {d['step_TFLOP']:.2f} TFLOP; "
                  f"{100 * d['roofline']['step_share_at_0p6']:.0f} % of the step time is spent in kernels at >= 0.6 of their roofline.",
    "r02z_*": f"{btag}_*`, `profiles/{tag}_*" if btag != tag else f"{tag}_*",
}
s = open(os.path.join(R, "scripts", "DESIGN.md.tmpl")).read()
for a, b in sub.items():
    s = s.replace(a, b)
open(os.path.join(R, "DESIGN.md"), "w").write(s)
print("DESIGN.md written;", d["value"], "Mpix/s")
