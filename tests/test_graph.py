"""Host logic (no GPU): yaml -> flat graph, reference error behaviour, planner decisions through the C ABI."""
import os

import numpy as np
import pytest
import yaml

REF = "/root/reference"
has_ref = os.path.isdir(os.path.join(REF, "config"))


def _strip(g):
    return [(l.op, l.src0, l.src1, l.ksize, l.stride, l.filters, l.pad, l.batch_normalize, l.activation) for l in g.layers]


def test_builtin_yolov3_structure():
    import yolo_v3_tf2_b200 as y3
    from yolo_v3_tf2_b200 import _lib
    m = y3.ParseModel.builtin_yolov3(80)
    g = m.graph
    assert len(m.conv_shapes) == 75
    assert sum(k * k * ci * co for k, ci, co, _ in m.conv_shapes) == 61_895_776
    assert sum(1 for l in g.layers if l.op == _lib.OP_SHORTCUT) == 23
    assert sum(1 for l in g.layers if l.op == _lib.OP_CONCAT) == 2
    assert sum(1 for l in g.layers if l.op == _lib.OP_UPSAMPLE) == 2
    assert len(g.outputs) == 3
    # only the three head 1x1 convs carry a bias; they have 3*(5+C) filters
    nobn = [(k, co) for k, ci, co, bn in m.conv_shapes if not bn]
    assert nobn == [(1, 255)] * 3
    # concat order is [upsampled, skip] (parse_model.py:116-126): 256+512 and 128+256
    cats = [l for l in g.layers if l.op == _lib.OP_CONCAT]
    assert [(g.channels(c.src0), g.channels(c.src1)) for c in cats] == [(256, 512), (128, 256)]
    # algorithmic FLOPs at 416 (65.864 GFLOP)
    p = m.plan(416, 416, 1)
    flops = 0
    ci = 0
    for l, pl in zip(g.layers, p["layers"]):
        if l.op == _lib.OP_CONV:
            k, cin, cout, _ = m.conv_shapes[ci]
            ci += 1
            flops += 2 * pl["H"] * pl["W"] * cout * k * k * cin
    assert abs(flops / 1e9 - 65.864) < 0.01


def test_builtin_yolov3_tiny_structure():
    """YOLOv3-tiny (reference config/models/yolov3_tiny/*.yaml; parse_model.py:78-99 maxpool): 13 convs, 6 pools,
    heads at H/32 and H/16; the planner pads the 16-filter stem to 32 stored channels and runs every conv on the
    tensor cores."""
    import yolo_v3_tf2_b200 as y3
    from yolo_v3_tf2_b200 import _lib
    m = y3.ParseModel.builtin_yolov3_tiny(80)
    g = m.graph
    assert [co for _, _, co, _ in m.conv_shapes] == [16, 32, 64, 128, 256, 512, 1024, 256, 512, 255, 128, 256, 255]
    pools = [l for l in g.layers if l.op == _lib.OP_MAXPOOL]
    assert [(l.ksize, l.stride, l.pad) for l in pools] == [(2, 2, 1)] * 5 + [(2, 1, 1)]
    cats = [l for l in g.layers if l.op == _lib.OP_CONCAT]
    assert [(g.channels(c.src0), g.channels(c.src1)) for c in cats] == [(128, 256)]
    assert len(g.outputs) == 2
    p = m.plan(416, 416, 4)
    L = p["layers"]
    outs = [L[t - 1] for t in g.outputs]
    assert [(o["H"], o["W"], o["C"]) for o in outs] == [(13, 13, 255), (26, 26, 255)]
    kinds = [l["kernel"] for l in L]
    assert kinds.count(1) == 13 and kinds.count(6) == 6 and kinds.count(2) == 0
    # the stride-1 'same' pool keeps 13x13; the logical channel count of the stem stays 16 (32 are stored)
    pool_layers = [i for i, l in enumerate(g.layers) if l.op == _lib.OP_MAXPOOL]
    assert (L[pool_layers[-1]]["H"], L[pool_layers[-1]]["W"]) == (13, 13)
    assert L[0]["C"] == 16 and L[0]["pix_stride"] == 32 and L[1]["pix_stride"] == 32
    assert sum(1 for l in L if l["fused_upsample"]) == 1
    with pytest.raises(_lib.Y3Unsupported):
        _mini({"a": [_conv(32), {"type": "maxpool", "size_xy": [2, 3], "stride_xy": [2, 2], "padding": "same"}]},
              [{"name": "head", "layers_config_file": "a", "outputs_layers": [-1]}])


@pytest.mark.skipif(not has_ref, reason="reference checkout not present")
def test_reference_tiny_yamls_match_builtin():
    import yolo_v3_tf2_b200 as y3
    from yolo_v3_tf2_b200 import graph
    builtin = y3.ParseModel.builtin_yolov3_tiny(80).graph
    g_ref = graph.load_model_config(os.path.join(REF, "config/models/yolov3_tiny/model.yaml"), 80)
    assert _strip(g_ref) == _strip(builtin)
    assert g_ref.outputs == builtin.outputs


@pytest.mark.skipif(not has_ref, reason="reference checkout not present")
def test_reference_thin_heads_yaml_matches_builtin():
    """config/models/yolov3/model_thin_heads.yaml (multi-output necks, negative entry_index, backbone taps before the
    shortcut adds -- SURVEY.md 8f-3): loads unchanged, equals the built-in thin-heads wiring, differs from model.yaml in
    exactly the four re-wired edges, and plans (the two tapped 3x3 convs feed both an Add and a neck, so their Adds
    cannot be fused into the conv epilogue and run as separate kernels)."""
    import yolo_v3_tf2_b200 as y3
    from yolo_v3_tf2_b200 import graph, _lib
    g_ref = graph.load_model_config(os.path.join(REF, "config/models/yolov3/model_thin_heads.yaml"), 80)
    thin = y3.ParseModel.builtin_yolov3(80, thin_heads=True)
    std = y3.ParseModel.builtin_yolov3(80).graph
    assert _strip(g_ref) == _strip(thin.graph) and g_ref.outputs == thin.graph.outputs == std.outputs
    diff = [i for i, (a, b) in enumerate(zip(_strip(thin.graph), _strip(std))) if a != b]
    assert diff == [83, 85, 94, 96]
    p = thin.plan(416, 416, 2)
    assert [p["layers"][i - 1]["H"] for i in thin.graph.outputs] == [13, 26, 52]
    assert len(thin.conv_shapes) == 75


@pytest.mark.parametrize("nclasses,filters", [(80, 255), (38, 129), (37, 126), (3, 24)])
def test_head_filters_expression(nclasses, filters):
    import yolo_v3_tf2_b200 as y3
    m = y3.ParseModel.builtin_yolov3(nclasses)
    assert [co for k, ci, co, bn in m.conv_shapes if not bn] == [filters] * 3


@pytest.mark.skipif(not has_ref, reason="reference checkout not present")
def test_reference_yamls_load_unchanged_and_match_builtin(monkeypatch):
    """The reference's own files are accepted as-is: current schema, legacy monolithic schema and the built-in
    description all flatten to the same graph."""
    import yolo_v3_tf2_b200 as y3
    from yolo_v3_tf2_b200 import graph
    builtin = y3.ParseModel.builtin_yolov3(80).graph
    g_cur = graph.load_model_config(os.path.join(REF, "config/models/yolov3/model.yaml"), 80)
    g_old = graph.load_model_config(os.path.join(REF, "config/yolov3_model.yaml"), 80)
    assert _strip(g_cur) == _strip(builtin)
    assert _strip(g_old) == _strip(builtin)
    assert g_cur.outputs == g_old.outputs == builtin.outputs
    # same call sequence as inference.py:87-96, from the reference repo root as CWD
    monkeypatch.chdir(REF)
    with open("config/models/yolov3/model.yaml") as f:
        cfg = yaml.safe_load(f)
    model = y3.ParseModel().build_model(None, cfg["sub_models_configs"], cfg["output_stage"], nclasses=80)
    assert _strip(model.graph) == _strip(builtin)
    m2 = y3.ParseModel().create_model(80, "config/yolov3_model.yaml")
    assert _strip(m2.graph) == _strip(builtin)
    # the empty decode-layer cfg is accepted (nothing to parse)
    assert os.path.getsize("config/yolov3_decode_layer.cfg") == 0


@pytest.mark.skipif(not has_ref, reason="reference checkout not present")
def test_reference_anchors_and_names():
    import yolo_v3_tf2_b200 as y3
    from yolo_v3_tf2_b200 import configs
    from yolo_v3_tf2_b200.core.utils import count_file_lines
    a = y3.get_anchors(os.path.join(REF, "datasets/coco2012/anchors.txt"))
    assert a.shape == (3, 3, 2)
    np.testing.assert_allclose(a.astype(np.float32), configs.coco_anchors(), rtol=0, atol=1e-7)
    assert count_file_lines(os.path.join(REF, "datasets/coco2012/coco.names")) == 80
    assert count_file_lines(os.path.join(REF, "datasets/pets_breed.names")) == 38   # SURVEY.md header note 3


def _mini(layers_by_file, subs, nclasses=2, output_stage="head"):
    import yolo_v3_tf2_b200 as y3
    return y3.ParseModel().build_model(None, subs, output_stage, nclasses=nclasses, layer_lists=layers_by_file)


def _conv(f, size=1, stride=1, bn=True, act="leaky"):
    d = {"type": "convolutional", "filters": f, "size": size, "stride": stride, "pad": 1, "activation": act}
    if bn:
        d["batch_normalize"] = 1
    return d


def test_reference_error_behaviour():
    """Same exception types as the reference for the same config mistakes (parse_model.py:48,140,157,227,277)."""
    subs = [{"name": "head", "layers_config_file": "a", "outputs_layers": [-1]}]
    with pytest.raises(ValueError, match="not recognized as layer_conf type"):
        _mini({"a": [{"type": "dropout"}]}, subs)
    with pytest.raises(AssertionError, match="Invalid activation"):
        _mini({"a": [_conv(32, act="relu")]}, subs)
    with pytest.raises(AssertionError, match="Invalid activation"):
        _mini({"a": [_conv(32), _conv(32), _conv(32), {"type": "shortcut", "from": -3, "activation": "leaky"}]}, subs)
    with pytest.raises(ValueError, match="Invalid number of layers"):
        _mini({"a": [_conv(32), _conv(32), _conv(32), {"type": "route", "source": {"layers": [0, 1, 2]}}]}, subs)
    with pytest.raises(IndexError):
        _mini({"a": [_conv(32)]}, [{"name": "head", "inputs": {"source": [{"name": "nope"}]}, "layers_config_file": "a",
                                    "outputs_layers": [-1]}])
    with pytest.raises(KeyError):
        _mini({"a": [{"type": "convolutional", "filters": 8, "activation": "linear"}]}, subs)


def test_planner_through_c_abi():
    """Fusion + arena decisions for the real network, on a planning-only context (no GPU needed)."""
    import yolo_v3_tf2_b200 as y3
    from yolo_v3_tf2_b200 import _lib
    m = y3.ParseModel.builtin_yolov3(80)
    g = m.graph
    p = m.plan(416, 416, 8)
    L = p["layers"]
    assert p["num_convs"] == 75
    kinds = [l["kernel"] for l in L]
    assert kinds.count(1) == 75 and kinds.count(2) == 0          # all 75 convs on the tcgen05 path (stem: gather mode)
    assert (L[0]["block_n"], L[0]["swizzle"]) == (32, 128)
    assert kinds.count(3) == kinds.count(4) == kinds.count(5) == 0   # every add / upsample / concat is fused
    assert sum(1 for l in L if l["fused_add"] >= 0) == 23
    assert sum(1 for l in L if l["fused_upsample"]) == 2
    # the flat-patch 3x3 kernel is opt-in (Y3_FLAT=1, see test_planner_flat_opt_in): by default nothing is haloed
    assert sum(1 for l in L if l["flat"]) == 0 and sum(1 for l in L if l["padded"]) == 0
    # concat operands live inside the concat buffer: same buffer id, channel offsets 0 and Ca, pixel stride Ca+Cb
    for i, l in enumerate(g.layers):
        if l.op == _lib.OP_CONCAT:
            cat, a, b = L[i], L[l.src0 - 1], L[l.src1 - 1]
            assert a["buffer"] == b["buffer"] == cat["buffer"]
            assert (a["chan_offset"], b["chan_offset"]) == (0, a["C"])
            assert a["pix_stride"] == b["pix_stride"] == cat["C"] == a["C"] + b["C"]
    # live buffers never overlap in the arena: rebuild liveness from the graph
    assert 0 < p["arena_bytes"] < 8 * 416 * 416 * 64 * 2 * 4
    # shapes follow H/32, H/16, H/8 for any multiple of 32 (the reference's Reshape breaks at 608; ours must not)
    p608 = m.plan(608, 608, 1)
    outs = [p608["layers"][t - 1] for t in g.outputs]
    assert [(o["H"], o["W"], o["C"]) for o in outs] == [(19, 19, 255), (38, 38, 255), (76, 76, 255)]
    with pytest.raises(ValueError):
        m.plan(400, 416, 1)


def test_planner_flat_opt_in():
    """Y3_FLAT=1 (an experiment knob of the profiling library liby3b200_prof.so, read when the library loads, hence
    the subprocess): all 32 3x3 stride-1 convs behind a 1x1 conv run the flat-patch kernel on a zero-haloed input.
    The release library ignores the environment."""
    import os
    import subprocess
    import sys
    code = (
        "import yolo_v3_tf2_b200 as y3\n"
        "m = y3.ParseModel.builtin_yolov3(80)\n"
        "L = m.plan(416, 416, 8)['layers']\n"
        "assert sum(1 for l in L if l['flat']) == 32 and sum(1 for l in L if l['padded']) == 32\n"
        "for i, l in enumerate(m.graph.layers):\n"
        "    if L[i]['flat']:\n"
        "        assert l.ksize == 3 and l.stride == 1 and L[l.src0 - 1]['padded'] == 1\n"
    )
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, Y3_FLAT="1", Y3_PROF_LIB="1", PYTHONPATH=root)
    subprocess.run([sys.executable, "-c", code], check=True, env=env, cwd=root)
    release = code.replace("== 32 and sum(1 for l in L if l['padded']) == 32", "== 0 and sum(1 for l in L if l['padded']) == 0")
    env = dict(os.environ, Y3_FLAT="1", PYTHONPATH=root)
    env.pop("Y3_PROF_LIB", None)
    subprocess.run([sys.executable, "-c", release], check=True, env=env, cwd=root)


def test_planner_arena_no_live_overlap():
    import yolo_v3_tf2_b200 as y3
    from yolo_v3_tf2_b200 import _lib
    m = y3.ParseModel.builtin_yolov3(80)
    g = m.graph
    B = 4
    p = m.plan(256, 256, B)
    L = p["layers"]
    # materialised tensor -> (buffer, first write layer, last read layer)
    bufs = {}
    for i, (l, pl) in enumerate(zip(g.layers, L)):
        if pl["buffer"] < 0:
            continue
        size = None
        b = bufs.setdefault(pl["buffer"], {"off": pl["arena_offset"], "first": i, "last": i, "bytes": 0})
        b["first"] = min(b["first"], i)
        b["bytes"] = max(b["bytes"], B * pl["H"] * pl["W"] * pl["pix_stride"] * 2)
        for j, l2 in enumerate(g.layers):
            if l2.src0 == i + 1 or l2.src1 == i + 1:
                b["last"] = max(b["last"], j)
    items = list(bufs.values())
    for x in range(len(items)):
        for y in range(x + 1, len(items)):
            a, b = items[x], items[y]
            live_overlap = not (a["last"] < b["first"] or b["last"] < a["first"])
            mem_overlap = not (a["off"] + a["bytes"] <= b["off"] or b["off"] + b["bytes"] <= a["off"])
            assert not (live_overlap and mem_overlap)


def test_unfusable_graphs_fall_back_to_standalone_kernels_or_reject():
    """A shortcut whose conv output is also used elsewhere cannot be fused -> planner schedules an add kernel;
    a maxpool gets its own kernel; pool sizes the kernel does not implement are rejected loudly -- there is no CPU
    fallback."""
    import yolo_v3_tf2_b200 as y3
    from yolo_v3_tf2_b200 import _lib
    layers = [_conv(32, 3), _conv(64, 3, 2), _conv(32, 1), _conv(64, 3),
              {"type": "shortcut", "from": -3, "activation": "linear"},
              {"type": "route", "source": {"layers": [-1, 3]}},      # conv3 output used twice -> add not fusable
              _conv("3*(2+2+1+nclasses)", 1, bn=False, act="linear"), {"type": "yolo", "grid_size": 13}]
    m = _mini({"a": layers}, [{"name": "head", "layers_config_file": "a", "outputs_layers": [-1]}])
    p = m.plan(64, 64, 1)
    kinds = [l["kernel"] for l in p["layers"]]
    assert 3 in kinds                       # stand-alone add
    tiny = [_conv(32, 3), {"type": "maxpool", "size_xy": [2, 2], "stride_xy": [2, 2], "padding": "same"},
            _conv("3*(2+2+1+nclasses)", 1, bn=False, act="linear"), {"type": "yolo", "grid_size": 13}]
    m2 = _mini({"a": tiny}, [{"name": "head", "layers_config_file": "a", "outputs_layers": [-1]}])
    assert [l["kernel"] for l in m2.plan(64, 64, 1)["layers"]][:2] == [1, 6]
    big = [_conv(32, 3), {"type": "maxpool", "size_xy": [5, 5], "stride_xy": [1, 1], "padding": "same"},
           _conv("3*(2+2+1+nclasses)", 1, bn=False, act="linear"), {"type": "yolo", "grid_size": 13}]
    m3 = _mini({"a": big}, [{"name": "head", "layers_config_file": "a", "outputs_layers": [-1]}])
    with pytest.raises(_lib.Y3Unsupported, match="maxpool"):
        m3.plan(64, 64, 1)


def test_planner_first_layer_kernels():
    """Round-2 kernel selection for the first Darknet-53 layers, through the planning-only C ABI (no GPU): the Cin = 32
    3x3 layers leave the 64-byte-row im2col path -- stride 2 on the pixel-pair view (128-byte rows, 64-wide N tile), stride
    1 on the band-resident kernel (64-byte swizzle, no operand ring) when the row fits one TMA box, else on the pixel-pair
    view too -- and the chained runs cover 62 of the 75 layers in 3 launches."""
    from yolo_v3_tf2_b200 import ParseModel
    m = ParseModel.builtin_yolov3(80)
    p = m.plan(416, 416, 64)
    L = p["layers"]
    assert (L[1]["swizzle"], L[1]["block_n"]) == (128, 64)               # 3x3/2 32->64 @208: pixel pairs
    assert (L[3]["swizzle"], L[3]["block_n"], L[3]["stages"]) == (64, 64, 2)   # 3x3 32->64 + Add @208: band resident
    steps = p["steps"]
    assert steps[3]["tiles_n"] == 1 and steps[3]["posts"] == 0 and steps[3]["chained"] == 0
    runs = {}
    for i, s in enumerate(steps):
        if s["run_first"] >= 0:
            runs[s["run_first"]] = runs.get(s["run_first"], 0) + 1
    assert sorted(runs.values(), reverse=True) == [50, 6, 6]
    launches = sum(1 for i, s in enumerate(steps) if s["run_first"] < 0 or s["run_first"] == i)
    assert launches == 16
    # 608 x 608: the 304-pixel rows of that layer do not fit one 256-wide TMA box -> pixel-pair view, N tile = pixel parity
    p6 = m.plan(608, 608, 8)
    assert (p6["layers"][3]["swizzle"], p6["layers"][3]["block_n"]) == (128, 64)
    assert p6["steps"][3]["tiles_n"] == 2
    # YOLOv3-tiny: both Cin = 32 3x3 layers (16 -> 32 @208 behind the first pool, 32 -> 64 @104) run band resident
    t = ParseModel.builtin_yolov3_tiny(80).plan(416, 416, 4)["layers"]
    assert (t[2]["swizzle"], t[2]["block_n"]) == (64, 32) and (t[4]["swizzle"], t[4]["block_n"]) == (64, 64)
