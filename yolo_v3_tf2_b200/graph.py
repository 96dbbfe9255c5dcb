"""YAML model description -> flat layer list for the C ABI (host logic only, no CUDA).

Mirrors the graph-construction half of reference core/parse_model.py:
  * ``build_model`` loop over ``sub_models_configs``        (parse_model.py:279-314)
  * ``create_sub_model_inputs`` producer lookup              (parse_model.py:216-246)
  * ``create_sub_model_layers`` per-entry dispatch           (parse_model.py:248-278)
  * ``_parse_route`` operand order ``layers`` then ``inputs`` (parse_model.py:102-140)
and additionally accepts the legacy monolithic ``config/yolov3_model.yaml`` (global Darknet indexing), which no code in
the reference reads any more but which describes the same network.

The output is a list of ``Layer`` records; tensor id 0 is the input image and layer i produces tensor i+1
(include/y3b200.h).  Identity routes create no layer: they only alias an existing tensor id.
"""
import ast
import os
from dataclasses import dataclass, field
from typing import List, Optional

import yaml

from . import _lib

OP_NAMES = {_lib.OP_CONV: "conv", _lib.OP_SHORTCUT: "shortcut", _lib.OP_UPSAMPLE: "upsample", _lib.OP_CONCAT: "concat",
            _lib.OP_YOLO: "yolo", _lib.OP_MAXPOOL: "maxpool"}


@dataclass
class Layer:
    op: int
    src0: int
    src1: int = -1
    ksize: int = 0
    stride: int = 0
    filters: int = 0
    pad: int = 0
    batch_normalize: int = 0
    activation: int = 0          # 0 linear, 1 leaky
    sub_model: str = ""
    yaml_index: int = -1          # position inside the sub-model's layers_config


@dataclass
class Graph:
    layers: List[Layer] = field(default_factory=list)
    outputs: List[int] = field(default_factory=list)        # tensor ids, model output order
    conv_layers: List[int] = field(default_factory=list)    # layer indices of convs in creation order
    sub_model_names: List[str] = field(default_factory=list)
    nclasses: int = 0
    # what the Keras checkpoint order needs (keras_weight_slots): per sub-model its producers in ``inputs.source`` order
    # ('' = the model Input), its output tensor ids in ``outputs_layers`` order, and the names of the model's outputs
    sub_model_sources: dict = field(default_factory=dict)
    sub_model_outputs: dict = field(default_factory=dict)
    output_sub_models: List[str] = field(default_factory=list)

    def add(self, layer: Layer) -> int:
        self.layers.append(layer)
        if layer.op == _lib.OP_CONV:
            self.conv_layers.append(len(self.layers) - 1)
        return len(self.layers)   # tensor id produced

    def channels(self, tensor_id: int) -> int:
        """Channel count of a tensor (needed for weight shapes)."""
        if tensor_id == 0:
            return 3
        l = self.layers[tensor_id - 1]
        if l.op == _lib.OP_CONV:
            return l.filters
        if l.op == _lib.OP_CONCAT:
            return self.channels(l.src0) + self.channels(l.src1)
        return self.channels(l.src0)

    def conv_shapes(self):
        """[(k, cin, cout, has_bn)] in conv creation order (= Keras conv2d_<i> / Darknet file order)."""
        out = []
        for li in self.conv_layers:
            l = self.layers[li]
            out.append((l.ksize, self.channels(l.src0), l.filters, bool(l.batch_normalize)))
        return out


def _eval_filters(expr, nclasses):
    """The reference eval()s string ``filters`` with ``nclasses`` in scope (parse_model.py:258-259); we evaluate the same
    arithmetic expressions without eval()."""
    if not isinstance(expr, str):
        return int(expr)

    def ev(node):
        if isinstance(node, ast.Expression):
            return ev(node.body)
        if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)):
            return node.value
        if isinstance(node, ast.Name) and node.id == "nclasses":
            return nclasses
        if isinstance(node, ast.BinOp) and isinstance(node.op, (ast.Add, ast.Sub, ast.Mult, ast.FloorDiv)):
            a, b = ev(node.left), ev(node.right)
            return {ast.Add: a + b, ast.Sub: a - b, ast.Mult: a * b, ast.FloorDiv: a // b if b else 0}[type(node.op)]
        if isinstance(node, ast.UnaryOp) and isinstance(node.op, ast.USub):
            return -ev(node.operand)
        raise ValueError(f"unsupported filters expression: {expr!r}")

    return int(ev(ast.parse(expr, mode="eval")))


def _conv_layer(conf, src, nclasses, sub_model, yaml_index):
    # parse_model.py:26-30 reads these four keys unconditionally -> KeyError when absent, as in the reference
    stride = int(conf["stride"])
    filters = _eval_filters(conf["filters"], nclasses)
    size = int(conf["size"])
    pad = int(conf["pad"])
    assert conf["activation"] in ["linear", "leaky"], "Invalid activation: {}".format(conf["activation"])
    return Layer(op=_lib.OP_CONV, src0=src, ksize=size, stride=stride, filters=filters, pad=pad,
                 batch_normalize=1 if "batch_normalize" in conf else 0,
                 activation=1 if conf["activation"] == "leaky" else 0, sub_model=sub_model, yaml_index=yaml_index)


def _maxpool_layer(conf, src, sub_model, yaml_index):
    """reference parse_model.py:78-99: MaxPooling2D(pool_size=size_xy, strides=stride_xy, padding=padding)."""
    stride_xy = list(map(int, conf["stride_xy"]))
    size_xy = list(map(int, conf["size_xy"]))
    padding = str(conf["padding"]).lower()
    if len(size_xy) != 2 or len(stride_xy) != 2 or size_xy[0] != size_xy[1] or stride_xy[0] != stride_xy[1]:
        raise _lib.Y3Unsupported(f"maxpool with non-square size/stride {size_xy}/{stride_xy}")
    if padding not in ("same", "valid"):
        raise ValueError(f"Invalid maxpool padding: {conf['padding']}")
    return Layer(op=_lib.OP_MAXPOOL, src0=src, ksize=size_xy[0], stride=stride_xy[0], pad=1 if padding == "same" else 0,
                 sub_model=sub_model, yaml_index=yaml_index)


def _resolve(path, search_dirs):
    """``layers_config_file`` paths are relative to the reference repo root (its CWD); try CWD, then search_dirs."""
    if os.path.isabs(path) and os.path.exists(path):
        return path
    for d in [os.getcwd()] + list(search_dirs):
        cand = os.path.join(d, path)
        if os.path.exists(cand):
            return cand
    raise FileNotFoundError(path)


def build_graph(sub_models_configs, output_stage="head", nclasses=0, search_dirs=(), layer_lists=None) -> Graph:
    """Flatten the reference's sub-model schema.  ``layer_lists`` optionally maps layers_config_file -> already parsed
    ``layers_config`` list (used by the built-in configs and by tests)."""
    g = Graph(nclasses=nclasses)
    produced = []   # [{'name', 'outputs': tensor id or list of tensor ids}]
    for sm in sub_models_configs:
        name = sm["name"]
        inputs_config = sm.get("inputs")
        g.sub_model_sources[name] = []
        if inputs_config:
            if "shape" in inputs_config:
                raise _lib.Y3Unsupported("sub-model inputs.shape (fresh Input) is not used by the yolov3 configs")
            entries = []
            for source_entry in inputs_config["source"]:
                sel = [p for p in produced if p["name"] == source_entry["name"]]
                src = sel[0]   # IndexError when the producer does not exist, like parse_model.py:227
                g.sub_model_sources[name].append(src["name"])
                idx = source_entry.get("entry_index", 0)
                out = src["outputs"]
                entries.append(out[idx] if isinstance(out, list) else out)
            inputs_entry = entries[0] if len(entries) == 1 else entries
        else:
            inputs_entry = 0   # the model input (parse_model.py:300)
            g.sub_model_sources[name].append("")

        if layer_lists is not None and sm["layers_config_file"] in layer_lists:
            layers_config = layer_lists[sm["layers_config_file"]]
        else:
            with open(_resolve(sm["layers_config_file"], search_dirs), "r") as stream:
                layers_config = yaml.safe_load(stream)["layers_config"]

        x = inputs_entry if not isinstance(inputs_entry, list) else inputs_entry   # current tensor
        layers = []   # tensor id per yaml entry (parse_model.py:254: every entry appends exactly one tensor)
        for yi, conf in enumerate(layers_config):
            t = conf["type"]
            if t == "convolutional":
                if isinstance(x, list):
                    raise ValueError("convolutional layer fed by a list of inputs")
                x = g.add(_conv_layer(conf, x, nclasses, name, yi))
            elif t == "shortcut":
                frm = layers[int(conf["from"])]
                assert conf["activation"] == "linear", "Invalid activation: {}".format(conf["activation"])
                x = g.add(Layer(op=_lib.OP_SHORTCUT, src0=x, src1=frm, sub_model=name, yaml_index=yi))
            elif t == "yolo":
                x = g.add(Layer(op=_lib.OP_YOLO, src0=x, sub_model=name, yaml_index=yi))
            elif t == "route":
                selected = []
                if "layers" in conf["source"]:
                    selected = [layers[int(l)] for l in conf["source"]["layers"]]
                if "inputs" in conf["source"]:
                    if isinstance(inputs_entry, list):
                        selected += [inputs_entry[i] for i in conf["source"]["inputs"]]
                    else:
                        selected += [inputs_entry]
                if len(selected) == 1:
                    x = selected[0]
                elif len(selected) == 2:
                    x = g.add(Layer(op=_lib.OP_CONCAT, src0=selected[0], src1=selected[1], sub_model=name, yaml_index=yi))
                else:
                    raise ValueError("Invalid number of layers: {}".format(len(selected)))
            elif t == "upsample":
                x = g.add(Layer(op=_lib.OP_UPSAMPLE, src0=x, stride=int(conf["stride"]), sub_model=name, yaml_index=yi))
            elif t == "maxpool":
                x = g.add(_maxpool_layer(conf, x, name, yi))
            else:
                raise ValueError("{} not recognized as layer_conf type".format(t))
            layers.append(x)

        outs = [layers[int(i)] for i in sm["outputs_layers"]]
        # Keras unwraps one-element output lists (parse_model.py:304-307)
        produced.append({"name": name, "outputs": outs[0] if len(outs) == 1 else outs})
        g.sub_model_names.append(name)
        g.sub_model_outputs[name] = list(outs)

    for p in produced:
        if output_stage in p["name"]:
            o = p["outputs"]
            g.outputs += o if isinstance(o, list) else [o]
            g.output_sub_models.append(p["name"])
    return g


def _keras_order(nodes, inbound, outputs):
    """Order of ``Model.layers`` of a Keras functional model (keras/engine/functional.py ``_map_graph_network``):
    decreasing depth (longest path to an output, in layers), ties broken by the order in which a depth-first walk from
    the outputs first completes a layer (inbound layers before the layer itself, outputs in their listed order).
    ``nodes``: hashable ids; ``inbound[n]``: list of ids feeding n (in call-argument order); -> list of ids."""
    index, order = {}, []

    def visit(n):
        stack = [(n, iter(inbound.get(n, ())))]
        seen_local = {n}
        while stack:
            cur, it = stack[-1]
            nxt = next(it, None)
            if nxt is None:
                stack.pop()
                if cur not in index:
                    index[cur] = len(order)
                    order.append(cur)
            elif nxt not in index and nxt not in seen_local:
                seen_local.add(nxt)
                stack.append((nxt, iter(inbound.get(nxt, ()))))

    for o in outputs:
        if o not in index:
            visit(o)
    depth = {n: 0 for n in order}
    for n in reversed(order):            # consumers before producers: order is a topological order
        for p in inbound.get(n, ()):
            if p in depth:
                depth[p] = max(depth[p], depth[n] + 1)
    return sorted(order, key=lambda n: (-depth[n], index[n])), depth


def keras_weight_slots(g: Graph):
    """Per conv (creation order): (i, j_conv, j_bn or None) such that the reference's Keras model stores the conv's
    variables under ``layer_with_weights-<i>/layer_with_weights-<j_conv>/{kernel,bias}`` and its batch-normalization
    under ``layer_with_weights-<i>/layer_with_weights-<j_bn>/...`` (``model.save_weights``, train.py:93-104).

    ``layer_with_weights-<n>`` counts the layers with weights in ``Model.layers`` order, and Keras sorts that list by
    decreasing depth, not by creation order (the reference's convert.py notes the same: ``model.layers`` is not in conv
    creation order).  For yolov3 the sub-models come out as backbone, neck0, neck1, neck2, head0, head1, head2; for
    yolov3-tiny as backbone, neck0, neck1, head0, head1.  The same rule is applied inside every sub-model, one Keras
    layer per ZeroPadding2D / Conv2D / BatchNormalization / LeakyReLU / Add / UpSampling2D / Concatenate / Reshape /
    MaxPooling2D the reference creates (core/parse_model.py:13-213)."""
    names = [n for n in g.sub_model_names]
    if not g.sub_model_sources:          # legacy monolithic schema: sub-models were consecutive slices of one chain
        top = names
    else:
        inbound = {n: [s if s else "__input__" for s in g.sub_model_sources.get(n, [])] for n in names}
        top, _ = _keras_order(names + ["__input__"], inbound, g.output_sub_models)
        top = [n for n in top if n != "__input__"]
    conv_index = {li: ci for ci, li in enumerate(g.conv_layers)}
    slots = {}
    i = 0
    for name in top:
        lids = [k for k, l in enumerate(g.layers) if l.sub_model == name]
        if not any(g.layers[k].op == _lib.OP_CONV for k in lids):
            continue
        mine = set(k + 1 for k in lids)       # tensor ids produced inside this sub-model
        # Keras layers of the sub-model: (layer id, position) with the per-IR-layer chain pad? conv bn? leaky?
        inbound, weighted = {}, {}

        def src_node(t):
            return ("L", t - 1, "end") if t in mine else ("in", t)

        for k in lids:
            l = g.layers[k]
            if l.op == _lib.OP_CONV:
                chain = (["pad"] if l.stride > 1 else []) + ["conv"] + (["bn"] if l.batch_normalize else []) + \
                        (["leaky"] if l.activation == 1 else [])
                prev = src_node(l.src0)
                for pos, kind in enumerate(chain):
                    node = ("L", k, "end") if pos == len(chain) - 1 else ("L", k, kind)
                    inbound[node] = [prev]
                    if kind in ("conv", "bn"):
                        weighted[node] = (k, kind)
                    prev = node
            elif l.op == _lib.OP_SHORTCUT:
                inbound[("L", k, "end")] = [src_node(l.src1), src_node(l.src0)]     # Add()([layers[from], x])
            elif l.op == _lib.OP_CONCAT:
                inbound[("L", k, "end")] = [src_node(l.src0), src_node(l.src1)]
            else:
                inbound[("L", k, "end")] = [src_node(l.src0)]
        outs = [src_node(t) for t in g.sub_model_outputs.get(name, [lids[-1] + 1])]
        order, _ = _keras_order(list(inbound), inbound, outs)
        j = 0
        for node in order:
            if node in weighted:
                k, kind = weighted[node]
                slots.setdefault(conv_index[k], {})[kind] = j
                j += 1
        i_used = i
        for ci in (conv_index[k] for k in lids if g.layers[k].op == _lib.OP_CONV):
            if ci in slots:
                slots[ci]["i"] = i_used
        i += 1
    out = []
    for ci in range(len(g.conv_layers)):
        s = slots.get(ci)
        if s is None or "conv" not in s:
            raise ValueError(f"conv {ci} is not reachable from its sub-model's outputs: Keras would not create it")
        out.append((s["i"], s["conv"], s.get("bn")))
    return out


def build_graph_legacy(model_config, nclasses=0) -> Graph:
    """Legacy monolithic schema (reference config/yolov3_model.yaml): ``sub_models: [{name, layers_config}]`` whose
    entries use global Darknet indexing (route ``layers: [-4]``, ``[-1, 61]``; shortcut ``from: -3``)."""
    g = Graph(nclasses=nclasses)
    tensors = []   # tensor id per global Darknet layer index
    x = 0
    for sm in model_config["sub_models"]:
        name = sm["name"]
        g.sub_model_names.append(name)
        for yi, conf in enumerate(sm["layers_config"]):
            t = conf["type"]
            here = len(tensors)

            def ref(i):
                i = int(i)
                return tensors[here + i] if i < 0 else tensors[i]

            if t == "convolutional":
                x = g.add(_conv_layer(conf, x, nclasses, name, yi))
            elif t == "shortcut":
                assert conf["activation"] == "linear", "Invalid activation: {}".format(conf["activation"])
                x = g.add(Layer(op=_lib.OP_SHORTCUT, src0=x, src1=ref(conf["from"]), sub_model=name, yaml_index=yi))
            elif t == "yolo":
                x = g.add(Layer(op=_lib.OP_YOLO, src0=x, sub_model=name, yaml_index=yi))
                g.outputs.append(x)
            elif t == "route":
                sel = [ref(i) for i in conf["layers"]]
                if len(sel) == 1:
                    x = sel[0]
                elif len(sel) == 2:
                    x = g.add(Layer(op=_lib.OP_CONCAT, src0=sel[0], src1=sel[1], sub_model=name, yaml_index=yi))
                else:
                    raise ValueError("Invalid number of layers: {}".format(len(sel)))
            elif t == "upsample":
                x = g.add(Layer(op=_lib.OP_UPSAMPLE, src0=x, stride=int(conf["stride"]), sub_model=name, yaml_index=yi))
            elif t == "maxpool":
                x = g.add(_maxpool_layer(conf, x, name, yi))
            else:
                raise ValueError("{} not recognized as layer_conf type".format(t))
            tensors.append(x)
    return g


def load_model_config(model_config_file, nclasses, search_dirs=()) -> Graph:
    """Read either schema from a yaml file (``yolov3_model.yaml`` or ``models/yolov3/model.yaml``)."""
    with open(model_config_file, "r") as stream:
        cfg = yaml.safe_load(stream)
    if "sub_models_configs" in cfg:
        here = os.path.dirname(os.path.abspath(model_config_file))
        # model.yaml lives in <root>/config/models/yolov3/, its paths are relative to <root>
        roots = [here, os.path.join(here, ".."), os.path.join(here, "..", ".."), os.path.join(here, "..", "..", "..")]
        return build_graph(cfg["sub_models_configs"], cfg.get("output_stage", "head"), nclasses,
                           search_dirs=list(search_dirs) + roots)
    if "sub_models" in cfg:
        return build_graph_legacy(cfg, nclasses)
    raise ValueError(f"{model_config_file}: neither 'sub_models_configs' nor 'sub_models' found")


def to_descs(g: Graph):
    arr = (_lib.LayerDesc * len(g.layers))()
    for i, l in enumerate(g.layers):
        arr[i] = _lib.LayerDesc(l.op, l.src0, l.src1, l.ksize, l.stride, l.filters, l.pad, l.batch_normalize, l.activation)
    return arr
