#!/usr/bin/env python
"""Bring-up harness for the tcgen05 contraction: runs tests/tc_cases.py case by case, restarting a fresh process
(fresh CUDA context) after a crash so that one faulting shape does not hide the others.
Writes gpurun_out/tc_unit.json."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


TMA = "--tma" in sys.argv      # bring-up of the TMA-fed kernel (engine 3) on TMA_CASES


def all_cases():
    from tc_cases import CASES, CONV3_CASES, HALF_CASES, TMA_CASES
    if TMA:
        return list(TMA_CASES), len(TMA_CASES)
    return CASES + HALF_CASES + CONV3_CASES, len(CASES) + len(HALF_CASES)


def child(start):
    import numpy as np
    from tc_cases import run_case, run_conv3_case, tolerance
    CASES, n1 = all_cases()
    for i in range(start, len(CASES)):
        print(json.dumps({"begin": i}), flush=True)
        if i >= n1:
            y, y_ref = run_conv3_case(CASES[i], 0, seed=i)
        else:
            y, y_ref = run_case(CASES[i], 3 if TMA else 0, seed=i)
        err = float(np.abs(y - y_ref).max()) if np.isfinite(y).all() else float("inf")
        bad = int((~np.isfinite(y)).sum())
        # where is the error? (row / column of the worst element) helps decode layout mistakes
        d = np.abs(np.nan_to_num(y, nan=1e9) - y_ref)
        r, c = np.unravel_index(int(d.argmax()), d.shape)
        colerr = d.max(axis=0)
        rowerr = d.max(axis=1)
        tol = tolerance(CASES[i], y_ref) if i < n1 else 4e-3 * float(np.abs(y_ref).max()) + 1e-5
        print(json.dumps({"case": i, "cfg": CASES[i], "err": err, "tol": tol, "ok": bool(err <= tol), "nonfinite": bad,
                          "worst": [int(r), int(c)], "bad_cols": int((colerr > tol).sum()),
                          "bad_rows": int((rowerr > tol).sum()),
                          "first_bad_cols": [int(x) for x in np.nonzero(colerr > tol)[0][:12]],
                          "first_bad_rows": [int(x) for x in np.nonzero(rowerr > tol)[0][:12]],
                          "ref_max": float(np.abs(y_ref).max())}), flush=True)


def main():
    if "--from" in sys.argv:
        child(int(sys.argv[sys.argv.index("--from") + 1]))
        return 0
    CASES, _ = all_cases()
    results, start = [], 0
    while start < len(CASES):
        p = subprocess.Popen([sys.executable, os.path.abspath(__file__), "--from", str(start)] + (["--tma"] if TMA else []),
                             stdout=subprocess.PIPE,
                             stderr=subprocess.PIPE, text=True)
        try:
            out, err = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            p.kill()
            out, err = p.communicate()
            err += "\nTIMEOUT"
        last_begin = start - 1
        for line in out.splitlines():
            try:
                j = json.loads(line)
            except Exception:
                continue
            if "begin" in j:
                last_begin = j["begin"]
            else:
                results.append(j)
        done = {r["case"] for r in results}
        if p.returncode != 0 or last_begin not in done:
            results.append({"case": last_begin, "cfg": CASES[last_begin] if last_begin >= 0 else None, "ok": False,
                            "crash": True, "stderr": err[-1500:]})
            start = last_begin + 1
        else:
            start = len(CASES)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "tma_unit.json" if TMA else "tc_unit.json"), "w"), indent=1)
    n_ok = sum(1 for r in results if r.get("ok"))
    print(f"tc_unit: {n_ok}/{len(CASES)} ok")
    for r in results:
        if not r.get("ok"):
            print(json.dumps(r)[:1200])
    return 0 if n_ok == len(CASES) else 1


if __name__ == "__main__":
    sys.exit(main())
