"""Full-size parity against the CPU oracle (BASELINE.json configs 2, 3-shard and 5) and the box tolerance.

north_star: "Conv/decode outputs must match within a stated bf16 tolerance (max abs and relative error on logits and
boxes)".  The stated tolerances below are about 2x the errors achieved on a B200 (profiles/r2_parity.md, written by
tools/parity_table.py with the same measurement code, tests/parity_util.py):

  logits, per head : relative L2 error <= 2e-2,  max-abs error <= 3.2e-2 * max|ref|
  boxes            : |centre error| <= 2.2e-3 (image-fraction units), |w/h relative error| <= 0.14, for the
                     boxes whose reference w/h logits are moderate (|t_wh| <= 2)
  objectness <= 2.1e-2, class probabilities <= 4.2e-2 (max-abs)

Reference path: inference.py:109-116 (model -> yolo_decode -> YoloNmsLayer), core/yolo_decode_layer.py:4-12.
"""
import numpy as np
import pytest

import parity_util

pytestmark = pytest.mark.gpu

# stated bf16 tolerances (activations are bf16 between the 75 convs, accumulation fp32) = ~2x the worst error achieved
# over every configuration of profiles/r2_parity.md (in brackets)
REL_L2_TOL = 2e-2        # [9.5e-3]
MAX_ABS_TOL = 3.2e-2     # [1.55e-2]  max-abs error / max |ref|
CENTRE_TOL = 2.2e-3      # [1.08e-3]  image-fraction units
WH_REL_TOL = 0.14        # [6.8e-2]
CONF_TOL = 2.1e-2        # [1.02e-2]  objectness
PROB_TOL = 4.2e-2        # [2.04e-2]  class probabilities
# keras-default init: logits are O(0.1), every error is far smaller in absolute terms
KERAS_MAX_ABS = 1.2e-2   # [6.0e-3]
NMS_MATCH_FLOOR = 0.8    # [0.91 on the near-tied keras-init scores, 0.995 on init-V]


def _check(res, keras=False):
    for k, h in enumerate(res["heads"]):
        assert h["rel_l2"] <= REL_L2_TOL, (k, h)
        if keras:
            assert h["max_abs"] <= KERAS_MAX_ABS, (k, h)
        else:
            assert h["max_abs_over_max_ref"] <= MAX_ABS_TOL, (k, h)
    b = res["boxes"]
    assert b["boxes_compared"] > 0.3 * b["boxes_total"]
    assert b["centre_abs"] <= CENTRE_TOL, b
    assert b["wh_rel"] <= WH_REL_TOL, b
    assert b["conf_abs"] <= CONF_TOL and b["prob_abs"] <= PROB_TOL, b
    # decode kernel on the GPU's own logits vs the numpy oracle on the same logits: float rounding only [1.9e-6]
    assert res["decode_kernel_vs_oracle_abs"] <= 1e-5, res["decode_kernel_vs_oracle_abs"]


@pytest.mark.parametrize("C", [80, 38, 37])
def test_boxes_of_gpu_logits_vs_oracle(cuda, C):
    """VERDICT r1 'missing 2': decode the GPU's own logits (init-V, 416^2) and bound the box error against the oracle's
    decode of the oracle's logits; plus the informational set-overlap of NMS on our own logits."""
    res = parity_util.measure("variance", 416, 2, C, seed=3)
    _check(res)
    n = res["nms"]
    assert n["ref_detections"] > 0
    # informational in SURVEY 8d; a loose floor still catches a broken pipeline (achieved: profiles/r2_parity.md)
    assert n["matched_frac"] >= NMS_MATCH_FLOOR, n


@pytest.mark.parametrize("init,size,B,C", [("variance", 416, 64, 80), ("variance", 608, 32, 80), ("variance", 416, 128, 37),
                                           ("keras", 416, 64, 80)])
def test_full_batches_vs_oracle(cuda, init, size, B, C):
    """BASELINE configs 2 (416^2 B=64), 3 (the 32-image 608^2 shard of the 8-GPU run) and 5 (37 classes, B=128) compared
    with net_oracle.forward directly, every image (the oracle runs at ~16 images/s on 16 cores); the informational NMS
    set overlap is taken over the first 6 images (the oracle NMS of dense init-V boxes takes seconds per image)."""
    res = parity_util.measure(init, size, B, C, seed=17, with_nms=(B <= 64), nms_images=6)
    _check(res, keras=(init == "keras"))
