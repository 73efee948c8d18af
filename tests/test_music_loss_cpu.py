"""CPU: pin the oracle of the music recipe's training loss (oracle/fqss_oracle_music.py: music_kd_loss, center_trim) to a
golden vector computed with the reference's own helpers in the order of its training loop (musdbhq_train.py:83-109;
tests/golden/make_golden.py music_loss)."""
import numpy as np
import torch

import fqss_oracle_music as M


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_music_kd_loss_oracle_matches_reference(golden):
    g = golden("music_loss.npz")
    wavs = T(g["wavs"]).clone().requires_grad_(True)
    src = M.center_trim(T(g["sources_full"]), wavs)
    loss, kd, task = M.music_kd_loss(wavs, T(g["fwavs"]), src, float(g["kd_lambda"]))
    assert torch.equal(loss.detach(), T(g["loss"])) and torch.equal(kd.detach(), T(g["kd"])) and torch.equal(task.detach(), T(g["task"]))
    loss.backward()
    assert torch.equal(wavs.grad, T(g["g"]))
    wavs.grad = None
    loss0, _, _ = M.music_kd_loss(wavs, None, src, 0.0)
    loss0.backward()
    assert torch.equal(loss0.detach(), T(g["loss0"])) and torch.equal(wavs.grad, T(g["g0"]))
