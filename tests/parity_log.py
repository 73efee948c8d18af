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


class reassociated_pointwise_convs:
    """Context manager: the oracle's 1x1 convolutions accumulate their input channels in a different (seeded random) order.
    Same arithmetic, same operands, another fp32 summation order -- i.e. a second, equally valid evaluation of the
    reference.  Tests use the deviation between the two oracle runs as the yardstick for quantities that the quantised
    block amplifies (single +-1 code moves change STE masks and rounding residuals): a tensor on which the reference
    disagrees with ITSELF by x cannot be held to less than x."""

    def __init__(self, seed=1234):
        self.seed = seed

    def __enter__(self):
        import torch
        import torch.nn.functional as F
        self._F, self._orig = F, F.conv1d
        gen = torch.Generator().manual_seed(self.seed)
        orig = self._orig

        def conv1d(x, w, bias=None, stride=1, padding=0, dilation=1, groups=1):
            if w.shape[-1] == 1 and groups == 1 and w.shape[1] > 1:
                perm = torch.randperm(w.shape[1], generator=gen)
                return orig(x[:, perm], w[:, perm], bias, stride, padding, dilation, groups)
            return orig(x, w, bias, stride, padding, dilation, groups)
        F.conv1d = conv1d
        return self

    def __exit__(self, *exc):
        self._F.conv1d = self._orig
        return False
