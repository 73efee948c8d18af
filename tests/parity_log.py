"""Measured-deviation log of the GPU parity tests.  Every `-m gpu` parity test records what it MEASURED (flip rates,
relative errors, cosines) next to the bound it asserts; the file lands in gpurun_out/ on the GPU box (merged back by
gpurun) and a copy is committed as profiles/parity_r02.json."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "gpurun_out", "parity_r02.json")


def record(test, **vals):
    try:
        os.makedirs(os.path.dirname(PATH), exist_ok=True)
        data = {}
        if os.path.exists(PATH):
            with open(PATH) as f:
                data = json.load(f)
        cur = data.setdefault(test, {})
        for k, v in vals.items():
            cur[k] = v if isinstance(v, (str, list, dict)) else float(v)
        with open(PATH, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
    except OSError:
        pass
    print("[parity] %s: %s" % (test, ", ".join("%s=%s" % (k, ("%.3e" % v) if isinstance(v, float) else v) for k, v in vals.items())))
