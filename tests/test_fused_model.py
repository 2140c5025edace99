"""CPU: the lane-level model of the fused ridge + Sauvola kernel's column role (oracle/fused_model.py: skewed lanes, rings,
hand-over between warps and bands, bit words) reproduces oracle/restate.py bit for bit.  Heights are chosen around the
124-row band and the 32-row warp blocks so that the 7 replicated bottom rows of cv2.boxFilter fall inside a block, at
the top of a block, and into the next band."""
import numpy as np
import pytest

import _cases
from oracle import fused_model, restate


@pytest.mark.parametrize("w,h,kind", [(64, 8, "grid"), (97, 131, "grid"), (200, 37, "noise"), (72, 124, "noise"), (96, 117, "grid"),
                                      (80, 118, "noise"), (70, 248, "noise"), (65, 152, "grid"), (130, 260, "grid")])
def test_fused_schedule_model_is_bit_exact(w, h, kind):
    img = (_cases.grid_u8 if kind == "grid" else _cases.noise_u8)(w, h, seed=3 * w + h)
    r = restate.frontend(img)
    binary, T = fused_model.fused_column_model(r["b"], h, w)
    assert np.array_equal(T.view(np.uint64), r["T"].view(np.uint64))
    assert np.array_equal(binary, r["binary"])
