"""Pin oracle/fqss_oracle_music.py to the UNMODIFIED reference ConvTasNetMusicQ on CPU (only where /root/reference
exists): state after 2 observer passes, forward output, and every parameter gradient of a scalar loss.  With `--golden`
also writes tests/golden/music_small.npz for tests/test_oracle_music_golden.py."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", "tests", "golden"))
import _ref_import as R  # noqa: E402

R.install()
from quantization.qat.models.convtasnetq_music import ConvTasNetMusicQ  # noqa: E402
import fqss_oracle as O  # noqa: E402
import fqss_oracle_music as M  # noqa: E402

KW = dict(sources=["a", "b", "c"], audio_channels=2, n_filters=16, kernel=20, stride=10, bn_chan=8, hid_chan=16, conv_kernel=3,
          n_blocks=3, n_repeats=2)
CFG = M.MusicConfig(n_src=3, audio_channels=2, n_filters=16, kernel=20, stride=10, bn_chan=8, hid_chan=16, conv_kernel=3,
                    n_blocks=3, n_repeats=2)


def main(golden=False):
    torch.manual_seed(0)
    torch.set_num_threads(4)
    ref = ConvTasNetMusicQ(**KW)
    ref.set_splitter_combiner(2, 2)
    ref.quantize_model()
    with torch.no_grad():
        for p in ref.parameters():                        # non-trivial affine / slope parameters
            if p.dim() == 1 and p.numel() > 1:
                p.add_(0.1 * torch.randn_like(p))
    init = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 2, 1210, generator=g) * 0.1
    P = O.Params({k: v.clone() for k, v in init.items()})
    st = M.calibrate_music(P, x, CFG, passes=2)
    with torch.no_grad():
        ref(x)
        ref(x)
    for m in ref.modules():
        if hasattr(m, "enable_observer"):
            m.enable_observer(False)
    calib = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    worst = max((calib[k] - P[k]).abs().max().item() for k in calib)
    print("ranges/params after calibration: max |ref-oracle| = %.3e" % worst)
    out_r = ref(x)
    coeff = torch.randn(out_r.shape, generator=g)
    (out_r * coeff).sum().backward()
    P.leafify()
    out_o = M.music_forward(P, x, CFG, st)
    (out_o * coeff).sum().backward()
    print("out   max|d| %.3e  shape %s" % ((out_r - out_o).abs().max().item(), tuple(out_r.shape)))
    gw, n_none = 0.0, 0
    for k, p in ref.named_parameters():
        go = P[k].grad
        if p.grad is None:
            n_none += 1
            assert go is None or float(go.abs().max()) == 0.0, k
            continue
        gw = max(gw, (p.grad - go).abs().max().item() / (p.grad.abs().max().item() + 1e-30))
    print("grads: worst max-normalised diff %.3e ; params without grad in reference: %d" % (gw, n_none))
    ok = worst == 0.0 and torch.equal(out_r, out_o) and gw == 0.0
    print("MUSIC ORACLE == REFERENCE" if ok else "MUSIC ORACLE DIFFERS")
    if golden and ok:
        d = {"x": x.numpy(), "coeff": coeff.numpy(), "out": out_r.detach().numpy()}
        for k, v in init.items():
            d["init/" + k] = v.numpy()
        for k, v in calib.items():
            d["calib/" + k] = v.numpy()
        for k, p in ref.named_parameters():
            if p.grad is not None:
                d["grad/" + k] = p.grad.numpy()
        path = os.path.join(HERE, "..", "tests", "golden", "music_small.npz")
        np.savez_compressed(path, **d)
        print("wrote", os.path.abspath(path), os.path.getsize(path), "bytes")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main("--golden" in sys.argv))
