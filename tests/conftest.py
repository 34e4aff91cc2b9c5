import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    import numpy as np
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    return meta, z


def golden_names(kind):
    out = []
    for f in sorted(os.listdir(GOLDEN)):
        if f.endswith(".npz"):
            import numpy as np
            z = np.load(os.path.join(GOLDEN, f), allow_pickle=False)
            if json.loads(str(z["meta"]))["kind"] == kind:
                out.append(f[:-4])
    return out


_REPORT = {}


def record(case, **kv):
    """Parity numbers land in gpurun_out/parity.json so a GPU run can be read back here."""
    _REPORT.setdefault(case, {}).update({k: (v if isinstance(v, str) else float(v)) for k, v in kv.items()})


def pytest_sessionfinish(session, exitstatus):
    if _REPORT:
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        path = os.path.join(out, "parity.json")
        old = {}
        if os.path.exists(path):
            try:
                old = json.load(open(path))
            except Exception:
                old = {}
        old.update(_REPORT)
        json.dump(old, open(path, "w"), indent=1, sort_keys=True)
