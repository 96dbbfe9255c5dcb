#!/usr/bin/env python
"""Achieved parity errors of the CUDA path against the CPU oracle, per head and per configuration (GPU only).

Same measurement code as tests/test_parity_full_gpu.py (tests/parity_util.py); writes a markdown report
(default profiles/r2_parity.md via gpurun_out/) so the stated tolerances in the tests can sit ~2x above what is achieved.
usage: python tools/parity_table.py [out.md]"""
import os as _os, sys as _sys
_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
_sys.path.insert(0, _ROOT)
_sys.path.insert(0, _os.path.join(_ROOT, "tests"))
import json
import time

import parity_util

CONFIGS = [
    # (init, size, B, C, with_nms)      BASELINE.json config it stands for
    ("variance", 416, 1, 80, True),     # config 1 shape (B=1), init-V
    ("keras", 416, 1, 80, True),        # config 1, the reference's fresh-model init
    ("variance", 416, 64, 80, True),    # config 2
    ("keras", 416, 64, 80, True),       # config 2, Keras-default init
    ("variance", 608, 32, 80, True),    # config 3: the 32-image shard of the 8-GPU run
    ("variance", 416, 128, 37, False),  # config 5 (37 classes)
    ("variance", 416, 2, 38, True),     # config 5 (38 classes = len(pets_breed.names))
]


def main():
    out = _sys.argv[1] if len(_sys.argv) > 1 else _os.path.join(_ROOT, "gpurun_out", "r2_parity.md")
    rows = []
    for init, size, B, C, with_nms in CONFIGS:
        t0 = time.time()
        r = parity_util.measure(init, size, B, C, seed=17, with_nms=with_nms)
        r["seconds"] = time.time() - t0
        rows.append(r)
        print(json.dumps(r), flush=True)
    with open(out, "w") as f:
        f.write("# Achieved parity errors: CUDA path (bf16 activations, fp32 accumulate) vs the torch-CPU fp32 oracle\n\n")
        f.write("`python tools/parity_table.py` on a B200; same code as `tests/test_parity_full_gpu.py` "
                "(`tests/parity_util.py`). Every image of every batch is compared. Oracle = `oracle/net_oracle.py` "
                "(parity unpinned: TensorFlow is not installable here).\n\n")
        f.write("## Logits, per head (13/26/52 grids at 416, 19/38/76 at 608)\n\n")
        f.write("| init | size | B | C | head | rel L2 | max abs | max abs / max ref | max ref |\n|---|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            for k, h in enumerate(r["heads"]):
                f.write(f"| {r['init']} | {r['size']} | {r['B']} | {r['C']} | {k} | {h['rel_l2']:.3e} | {h['max_abs']:.3e} | "
                        f"{h['max_abs_over_max_ref']:.3e} | {h['max_ref']:.3f} |\n")
        f.write("\n## Decoded boxes of the GPU's own logits vs the oracle's decode of the oracle's logits\n\n")
        f.write("Boxes with |t_wh| <= 2 in the reference logits; image-fraction units.\n\n")
        f.write("| init | size | B | C | boxes compared | centre max abs | w/h max rel | corner max abs | objectness max abs | "
                "class prob max abs | decode kernel vs numpy on same logits |\n|---|---|---|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            b = r["boxes"]
            f.write(f"| {r['init']} | {r['size']} | {r['B']} | {r['C']} | {b['boxes_compared']}/{b['boxes_total']} | "
                    f"{b['centre_abs']:.3e} | {b['wh_rel']:.3e} | {b['box_abs']:.3e} | {b['conf_abs']:.3e} | {b['prob_abs']:.3e} | "
                    f"{r['decode_kernel_vs_oracle_abs']:.2e} |\n")
        f.write("\n## NMS on our own logits vs NMS on the oracle's logits, as sets (informational, SURVEY 8d config 2)\n\n")
        f.write("max 100 boxes, IoU 0.5, score 0.1; a reference detection counts as matched when a GPU detection of the "
                "same class overlaps it with IoU >= 0.9.\n\n")
        f.write("| init | size | B | C | reference detections | GPU detections | matched | matched frac | images with equal count |\n"
                "|---|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            if "nms" in r:
                n = r["nms"]
                f.write(f"| {r['init']} | {r['size']} | {r['B']} | {r['C']} | {n['ref_detections']} | {n['gpu_detections']} | "
                        f"{n['matched']} | {n['matched_frac']:.4f} | {n['images_same_count_frac']:.3f} |\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
