"""CPU checks of the assumption and the protocol behind the rectangle transport of the observation all-gather
(csrc/obs_gather_kernel.cu) -- against the canonical oracle, no GPU:

1. a rendered frame equals the flat image (the frame of an env without contact) bit for bit OUTSIDE the rectangle the Taxim
   kernel reports (contact bounding box + 63 px of blur reach + 1 px of central differences, columns aligned to 4 pixels,
   border rows / columns following their inner neighbour) -- restated here from taxim_kernel.cu;
2. pushing only the rectangles and restoring "old rectangle minus new rectangle" from the flat image reproduces whole frames
   over a sequence of steps with moving, growing and disappearing contacts (the tx_obs_push / tx_obs_fill protocol in numpy).
"""
import numpy as np
import pytest
import torch

from conftest import H, W

GROW = 30 + 16 + 8 + 4 + 2 + 1 + 2


def kernel_rect(hm: np.ndarray, press: float):
    """(row0, row1, col0, col1) inclusive, or None without contact: the rectangle of taxim_kernel.cu (whole frame, both halves)."""
    if not press > 0.0:
        return None
    h = (hm - hm.min()) - np.float32(press)
    ys, xs = np.nonzero(h < 0)
    if ys.size == 0:
        return None
    fr0, fr1 = max(ys.min() - GROW, 0), min(ys.max() + GROW, H - 1)
    fc0, fc1 = max(xs.min() - GROW, 0), min(xs.max() + GROW, W - 1)
    a0 = 0 if fr0 - 1 <= 1 else fr0 - 1
    a1 = H - 1 if fr1 + 1 >= H - 2 else fr1 + 1
    xa = 0 if fc0 - 1 <= 1 else (fc0 - 1) & ~3
    xb = W - 1 if fc1 + 1 >= W - 2 else (fc1 + 1) | 3
    return int(a0), int(a1), int(xa), int(xb)


def outside(rect) -> np.ndarray:
    m = np.ones((H, W), bool)
    if rect is not None:
        m[rect[0]:rect[1] + 1, rect[2]:rect[3] + 1] = False
    return m


@pytest.fixture(scope="module")
def flat(canon_taxim):
    hm = np.full((1, H, W), 29.0, np.float32)
    return canon_taxim.render(hm, np.zeros(1, np.float32), want=("rgb",))["rgb"][0]


@pytest.mark.parametrize("name", ["config0", "config1_sub", "config2_sub"])
def test_frame_equals_flat_image_outside_the_kernel_rectangle(canon_taxim, inputs, flat, name):
    hm = inputs[name].numpy()
    press = canon_taxim.indentation_depth(hm)
    rgb = canon_taxim.render(hm, press, want=("rgb",))["rgb"]
    n_in = 0
    for i in range(hm.shape[0]):
        rect = kernel_rect(hm[i], press[i])
        o = outside(rect)
        assert np.array_equal(rgb[i][o], flat[o]), f"{name}[{i}]: pixels outside the rectangle differ from the flat image"
        n_in += (~o).sum()
    assert n_in < 0.9 * hm.shape[0] * H * W  # the rectangles are a real saving on these inputs


def test_rectangle_at_the_image_border(canon_taxim, flat):
    """Contacts whose rectangle is clipped by the image border, incl. the replicate-padded border rows / columns."""
    from tacex_b200 import synth

    d = [synth.depth_map(0, 2e-3, cx, cy, 0.0, 8e-4) for cx, cy in [(-9.2e-3, -6.9e-3), (9.2e-3, 6.9e-3), (0.0, -6.9e-3), (9.3e-3, 0.0)]]
    hm = synth.height_map_mm(torch.stack(d)).numpy()
    press = canon_taxim.indentation_depth(hm)
    rgb = canon_taxim.render(hm, press, want=("rgb",))["rgb"]
    for i in range(hm.shape[0]):
        rect = kernel_rect(hm[i], press[i])
        assert rect is not None
        o = outside(rect)
        assert np.array_equal(rgb[i][o], flat[o])


def test_push_and_restore_protocol_reproduces_whole_frames(canon_taxim, inputs, flat):
    """tx_obs_push stores the new rectangles into the receiver's buffer, tx_obs_fill restores (old rectangle) - (new rectangle)
    from the flat image; the receiver's previous-rectangle table starts as the whole frame (first fill = complete fill)."""
    hm0 = inputs["config2_sub"].numpy()
    steps = [hm0, np.roll(hm0, (9, -14), axis=(1, 2)), np.roll(hm0, (-31, 40), axis=(1, 2)), hm0.copy(), np.roll(hm0, 3, axis=0)]
    steps[3][::2] = hm0.max()  # every other env loses contact
    n = hm0.shape[0]
    buf = np.full((n, H, W, 3), np.nan, np.float32)  # receiver's copy of the sender's envs, never initialised
    prev = [(0, H - 1, 0, W - 1)] * n
    for hm in steps:
        press = canon_taxim.indentation_depth(hm)
        rgb = canon_taxim.render(hm, press, want=("rgb",))["rgb"]
        for i in range(n):
            new = kernel_rect(hm[i], press[i])
            if new is not None:  # push
                buf[i, new[0]:new[1] + 1, new[2]:new[3] + 1] = rgb[i, new[0]:new[1] + 1, new[2]:new[3] + 1]
            old = prev[i]
            if old is not None:  # fill: restore what the old rectangle covered and the new one does not
                m = np.zeros((H, W), bool)
                m[old[0]:old[1] + 1, old[2]:old[3] + 1] = True
                if new is not None:
                    m[new[0]:new[1] + 1, new[2]:new[3] + 1] = False
                buf[i][m] = flat[m]
            prev[i] = new
        assert np.array_equal(buf, rgb)
