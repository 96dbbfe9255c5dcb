"""Built-in YOLOv3 (Darknet-53 + 3-scale neck/heads) description in the reference's sub-model schema.

The reference ships this topology as yaml files (config/models/yolov3/{model,backbone,neck0-2,head0-2}.yaml).  Those
files are *inputs* of this package -- ``load_model_config`` reads them unchanged -- but they do not exist on a machine
without a reference checkout, so the same structure is generated here from the published Darknet-53 recipe.
tests/test_graph.py checks that the generated description equals the parsed reference yamls whenever
/root/reference is present.
"""

FILTER_EXPR = "3*(2+2+1+nclasses)"


def _conv(filters, size, stride=1, bn=True, act="leaky"):
    d = {"type": "convolutional"}
    if bn:
        d["batch_normalize"] = 1
    d.update({"filters": filters, "size": size, "stride": stride, "pad": 1, "activation": act})
    return d


def _route(layers=None, inputs=None):
    src = {}
    if layers is not None:
        src["layers"] = list(layers)
    if inputs is not None:
        src["inputs"] = list(inputs)
    return {"type": "route", "source": src}


def _residual_stage(out, filters, blocks):
    out.append(_conv(filters, 3, stride=2))
    for _ in range(blocks):
        out.append(_conv(filters // 2, 1))
        out.append(_conv(filters, 3))
        out.append({"type": "shortcut", "from": -3, "activation": "linear"})


def backbone_layers():
    l = [_route(inputs=[0]), _conv(32, 3)]
    for filters, blocks in ((64, 1), (128, 2), (256, 8), (512, 8), (1024, 4)):
        _residual_stage(l, filters, blocks)
    return l   # 76 entries; outputs -39 (52x52x256), -14 (26x26x512), -1 (13x13x1024)


def neck_layers(filters, lateral):
    """filters: the bottleneck width (512/256/128); lateral: None for neck0, else (input index of the coarser neck,
    input index of the backbone skip)."""
    l = []
    if lateral is None:
        l.append(_route(inputs=[0]))
    else:
        coarse, skip = lateral
        l += [_route(inputs=[coarse]), _conv(filters, 1), {"type": "upsample", "stride": 2},
              _route(layers=[-1], inputs=[skip])]
    for i in range(5):
        l.append(_conv(filters if i % 2 == 0 else filters * 2, 1 if i % 2 == 0 else 3))
    return l


def head_layers(filters, grid_size):
    return [_route(inputs=[0]), _conv(filters, 3),
            _conv(FILTER_EXPR, 1, bn=False, act="linear"),
            {"type": "yolo", "grid_size": grid_size, "jitter": 0.3}]


def yolov3_config(thin_heads=False):
    """Returns (model_config dict in the reference's model.yaml schema, {layers_config_file: layers_config list}).
    ``thin_heads``: the wiring of config/models/yolov3/model_thin_heads.yaml instead of model.yaml -- same layer files,
    but the backbone taps layers 36 / 61 (the 3x3 convs BEFORE their shortcut adds), every neck exposes its last two
    layers and the consumers pick them with negative / positive entry_index values."""
    files = {
        "builtin/yolov3/backbone.yaml": backbone_layers(),
        "builtin/yolov3/neck0.yaml": neck_layers(512, None),
        "builtin/yolov3/head0.yaml": head_layers(1024, 13),
        "builtin/yolov3/neck1.yaml": neck_layers(256, (1, 0)),
        "builtin/yolov3/head1.yaml": head_layers(512, 26),
        "builtin/yolov3/neck2.yaml": neck_layers(128, (0, 1)),
        "builtin/yolov3/head2.yaml": head_layers(256, 52),
    }

    def src(*pairs):
        return {"source": [{"name": n, "entry_index": i} for n, i in pairs]}

    subs = [
        {"name": "backbone", "layers_config_file": "builtin/yolov3/backbone.yaml", "outputs_layers": [-39, -14, -1]},
        {"name": "neck0", "inputs": src(("backbone", 2)), "layers_config_file": "builtin/yolov3/neck0.yaml", "outputs_layers": [-1]},
        {"name": "head0", "inputs": src(("neck0", 0)), "layers_config_file": "builtin/yolov3/head0.yaml", "outputs_layers": [-1]},
        {"name": "neck1", "inputs": src(("backbone", 1), ("neck0", 0)), "layers_config_file": "builtin/yolov3/neck1.yaml", "outputs_layers": [-1]},
        {"name": "head1", "inputs": src(("neck1", 0)), "layers_config_file": "builtin/yolov3/head1.yaml", "outputs_layers": [-1]},
        {"name": "neck2", "inputs": src(("neck1", 0), ("backbone", 0)), "layers_config_file": "builtin/yolov3/neck2.yaml", "outputs_layers": [-1]},
        {"name": "head2", "inputs": src(("neck2", 0)), "layers_config_file": "builtin/yolov3/head2.yaml", "outputs_layers": [-1]},
    ]
    if thin_heads:   # config/models/yolov3/model_thin_heads.yaml:5-83
        subs = [
            {"name": "backbone", "layers_config_file": "builtin/yolov3/backbone.yaml", "outputs_layers": [36, 61, -1]},
            {"name": "neck0", "inputs": src(("backbone", 2)), "layers_config_file": "builtin/yolov3/neck0.yaml", "outputs_layers": [-2, -1]},
            {"name": "head0", "inputs": src(("neck0", -1)), "layers_config_file": "builtin/yolov3/head0.yaml", "outputs_layers": [-1]},
            {"name": "neck1", "inputs": src(("backbone", 1), ("neck0", -2)), "layers_config_file": "builtin/yolov3/neck1.yaml", "outputs_layers": [-2, -1]},
            {"name": "head1", "inputs": src(("neck1", 1)), "layers_config_file": "builtin/yolov3/head1.yaml", "outputs_layers": [-1]},
            {"name": "neck2", "inputs": src(("neck1", 0), ("backbone", 0)), "layers_config_file": "builtin/yolov3/neck2.yaml", "outputs_layers": [-2, -1]},
            {"name": "head2", "inputs": src(("neck2", 1)), "layers_config_file": "builtin/yolov3/head2.yaml", "outputs_layers": [-1]},
        ]
    model = {"decay_factor": 0.0005, "output_stage": "head", "grid_sizes": [13, 26, 52], "sub_models_configs": subs}
    return model, files


def _maxpool(size, stride):
    return {"type": "maxpool", "size_xy": [size, size], "stride_xy": [stride, stride], "padding": "same"}


def yolov3_tiny_config():
    """YOLOv3-tiny in the reference's schema (config/models/yolov3_tiny/{model,backbone,neck0-1,head0-1}.yaml):
    7 conv + 6 maxpool backbone (the last pool is 2x2 stride 1 'same'), two heads at H/32 and H/16."""
    backbone = [_route(inputs=[0])]
    for f in (16, 32, 64, 128, 256):
        backbone += [_conv(f, 3), _maxpool(2, 2)]
    backbone += [_conv(512, 3), _maxpool(2, 1), _conv(1024, 3)]
    files = {
        "builtin/yolov3_tiny/backbone.yaml": backbone,
        "builtin/yolov3_tiny/neck0.yaml": [_route(inputs=[0]), _conv(256, 1)],
        "builtin/yolov3_tiny/head0.yaml": [_route(inputs=[0]), _conv(512, 3), _conv(FILTER_EXPR, 1, bn=False, act="linear"),
                                           {"type": "yolo", "grid_size": 13, "jitter": 0.3}],
        "builtin/yolov3_tiny/neck1.yaml": [_route(inputs=[-2]), _conv(128, 1), {"type": "upsample", "stride": 2},
                                           _route(layers=[-1], inputs=[-1])],
        "builtin/yolov3_tiny/head1.yaml": [_route(inputs=[0]), _conv(256, 3), _conv(FILTER_EXPR, 1, bn=False, act="linear"),
                                           {"type": "yolo", "grid_size": 26, "jitter": 0.3}],
    }

    def src(*pairs):
        return {"source": [{"name": n, "entry_index": i} for n, i in pairs]}

    subs = [
        {"name": "backbone", "layers_config_file": "builtin/yolov3_tiny/backbone.yaml", "outputs_layers": [-5, -1]},
        {"name": "neck0", "inputs": src(("backbone", -1)), "layers_config_file": "builtin/yolov3_tiny/neck0.yaml", "outputs_layers": [-1]},
        {"name": "head0", "inputs": src(("neck0", 0)), "layers_config_file": "builtin/yolov3_tiny/head0.yaml", "outputs_layers": [-1]},
        {"name": "neck1", "inputs": src(("neck0", -1), ("backbone", -2)), "layers_config_file": "builtin/yolov3_tiny/neck1.yaml", "outputs_layers": [-1]},
        {"name": "head1", "inputs": src(("neck1", 0)), "layers_config_file": "builtin/yolov3_tiny/head1.yaml", "outputs_layers": [-1]},
    ]
    model = {"decay_factor": 0.0005, "output_stage": "head", "grid_sizes": [13, 26], "sub_models_configs": subs}
    return model, files


# the reference's tiny anchors are not shipped; the Darknet yolov3-tiny anchors / 416, largest first, as a [2,3,2] table
TINY_ANCHORS_PX = [(81, 82), (135, 169), (344, 319), (10, 14), (23, 27), (37, 58)]


def tiny_anchors():
    import numpy as np
    return (np.array(TINY_ANCHORS_PX, dtype=np.float64) / 416.0).astype(np.float32).reshape(2, 3, 2)


# COCO anchors of the reference (datasets/coco2012/anchors.txt:1-9 = the Darknet yolov3 anchors / 416), largest first
COCO_ANCHORS_PX = [(116, 90), (156, 198), (373, 326), (30, 61), (62, 45), (59, 119), (10, 13), (16, 30), (33, 23)]


def coco_anchors():
    import numpy as np
    return (np.array(COCO_ANCHORS_PX, dtype=np.float64) / 416.0).astype(np.float32).reshape(3, 3, 2)
