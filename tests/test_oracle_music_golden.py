"""CPU: pin the oracle of the NEXT scope row (ConvTasNetMusicQ: cLN, skip-less blocks, linear
decoder + overlap-add, un-normalised splitter) to golden vectors produced by the unmodified reference
(oracle/check_music_against_reference.py --golden).  There is no CUDA path for this model yet; this is its first gate."""
import numpy as np
import torch

import fqss_oracle as O
import fqss_oracle_music as M


def T(a):
    return torch.from_numpy(np.asarray(a))


CFG = M.MusicConfig(n_src=3, audio_channels=2, n_filters=16, kernel=20, stride=10, bn_chan=8, hid_chan=16, conv_kernel=3,
                    n_blocks=3, n_repeats=2)


def test_music_oracle_calibration_forward_backward(golden):
    g = golden("music_small.npz")
    nthr = torch.get_num_threads()
    torch.set_num_threads(4)            # the generator's setting: ATen chunks its reductions per thread
    try:
        _check(g)
    finally:
        torch.set_num_threads(nthr)


def _check(g):
    x = T(g["x"])
    P = O.Params({k[5:]: T(g[k]).clone() for k in g.files if k.startswith("init/")})
    st = M.calibrate_music(P, x, CFG, passes=2)
    for k in P:
        assert torch.equal(P[k], T(g["calib/" + k])), k                 # observer EMA + first-call weight ranges
    P.leafify()
    out = M.music_forward(P, x, CFG, st)
    assert torch.equal(out, T(g["out"]))
    (out * T(g["coeff"])).sum().backward()
    n = 0
    for k in g.files:
        if k.startswith("grad/"):
            # bit-identical with the generator's thread count; ATen's LayerNorm / GroupNorm weight-gradient
            # reductions are chunked per thread, so allow summation-order round-off on other hosts
            assert torch.allclose(P[k[5:]].grad, T(g[k]), rtol=1e-4, atol=1e-6), k
            n += 1
    assert n > 100


def test_music_pieces():
    # overlap-add of frames with hop < length, and the un-normalised splitter's reconstruction identity
    frames = torch.arange(24.0).reshape(1, 3, 8)
    y = M.overlap_and_add(frames, 4)
    want = torch.zeros(1, 16)
    for i in range(3):
        want[0, 4 * i:4 * i + 8] += frames[0, i]
    assert torch.equal(y, want)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 2, 500, generator=g) * 3.0
    s = M.split_input_unnormalised(x, 2)
    assert s.shape == (2, 4, 500)
    thr = x.abs().max()
    msb, lsb = s[:, :2], s[:, 2:]
    rec = msb + (lsb + thr) * (thr / 128) / (2 * thr)                       # invert x <- 2 (x - q) thr / delta - thr
    assert (rec - x).abs().max() <= thr / 128 / 128 * 1.01
