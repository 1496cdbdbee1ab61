"""Harness formats (fincflow_b200/harness.py): the reference's checkpoint dict
(fastflow/train/experiment.py:400-427) and its sample-grid PNGs (experiment.py:342-345)."""
import io
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from fincflow_b200 import harness


def test_sample_grid_png_matches_torchvision(tmp_path):
    tv = pytest.importorskip("torchvision")
    from PIL import Image

    torch.manual_seed(0)
    for shape in [(23, 3, 32, 32), (7, 1, 28, 28), (10, 3, 8, 8)]:
        x = torch.rand(*shape) * 256
        ours = str(tmp_path / "ours.png")
        theirs = str(tmp_path / "theirs.png")
        harness.save_image_grid(x, ours)
        tv.utils.save_image(x / 256., theirs, nrow=10, padding=2, normalize=False)   # the reference's call
        a, b = np.asarray(Image.open(ours)), np.asarray(Image.open(theirs))
        assert a.shape == b.shape and np.array_equal(a, b)


def test_checkpoint_round_trip_in_the_reference_format(tmp_path):
    from fincflow_b200.flows import FastFlow

    torch.manual_seed(1)
    m = FastFlow(n_blocks=2, block_size=1, image_size=(1, 8, 8), actnorm=True, width=32)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    path = str(tmp_path / "ckpt.pt")
    ck = harness.save_checkpoint(path, m, opt, summary={"Epoch": 3}, config={"lr": 1e-3})
    assert tuple(ck.keys()) == harness.CHECKPOINT_KEYS
    # the reference's own key names (conv_tl.conv.weight ..., coupling.net.4.logs)
    keys = list(ck["model_state_dict"].keys())
    assert any(k.endswith("fastflow_unit.conv_tl.conv.weight") for k in keys)
    assert any(k.endswith("coupling.net.4.logs") for k in keys)
    m2 = FastFlow(n_blocks=2, block_size=1, image_size=(1, 8, 8), actnorm=True, width=32)
    summary, config = harness.load_checkpoint(path, m2, torch.optim.Adam(m2.parameters()))
    assert summary == {"Epoch": 3} and config == {"lr": 1e-3}
    for (n1, p1), (n2, p2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert n1 == n2 and torch.equal(p1, p2)
    # a DataParallel-trained checkpoint (`module.` prefix, experiment.py:174-176) and the reference's extra keys
    sd = {"module." + k: v for k, v in m.state_dict().items()}
    sd["module.preprocess.layers.0.distribution.empty"] = torch.zeros(0)
    m3 = FastFlow(n_blocks=2, block_size=1, image_size=(1, 8, 8), actnorm=True, width=32)
    harness.load_checkpoint({"model_state_dict": sd, "summary": {}, "config": {}}, m3)
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m3.state_dict().values()))
    with pytest.raises(RuntimeError):
        harness.load_checkpoint({"model_state_dict": {"nope": torch.zeros(1)}}, m3)
