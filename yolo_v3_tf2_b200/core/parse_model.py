"""Drop-in for reference core/parse_model.py: ``ParseModel().build_model(...)`` returns a callable model whose forward
pass runs the tcgen05 conv kernels of liby3b200.so instead of Keras layers."""
import ctypes as C

import numpy as np
import torch
import yaml

from .. import _lib
from .. import graph as graph_mod
from .. import weights as weights_mod


class Y3Model:
    """What ``ParseModel.build_model`` returns: callable like the Keras model (``model(x)`` -> list of 3 grids
    ``[B, g, g, 3, 5+nclasses]`` float32 in head0/head1/head2 order, reference parse_model.py:310-313) with the weight
    accessors the reference uses (``load_weights``, ``set_weights``/``get_weights`` in Keras variable order,
    ``predict``)."""

    def __init__(self, g: graph_mod.Graph, name="yolo"):
        self.name = name
        self.graph = g
        self.nclasses = g.nclasses
        self.conv_shapes = g.conv_shapes()
        self._params = None       # list[weights_mod.ConvParams]
        self._nets = {}           # (device, H, W) -> dict(handle, max_batch)
        self._retired = []        # nets replaced by a larger-batch plan: kept alive until close(), because CUDA graphs
                                  # captured on them (Detector.detections_graphed) hold raw pointers into their arena,
                                  # weights and tensor maps
        self._descs = graph_mod.to_descs(g)

    # ---------------- weights ----------------
    def set_params(self, params):
        if len(params) != len(self.conv_shapes):
            raise ValueError(f"expected {len(self.conv_shapes)} conv parameter sets, got {len(params)}")
        for p, (k, cin, cout, bn) in zip(params, self.conv_shapes):
            if p.kernel.shape != (k, k, cin, cout) or p.has_bn != bn:
                raise ValueError(f"conv parameters do not match the model: kernel {p.kernel.shape} vs {(k, k, cin, cout)}")
        self._params = list(params)
        for ent in list(self._nets.values()) + self._retired:
            self._upload(ent["handle"])

    def set_weights(self, arrays):
        """Keras order: per conv layer [kernel, (bias)] followed by its BN layer's [gamma, beta, mean, variance]."""
        self.set_params(weights_mod.params_from_list(arrays, self.conv_shapes))

    def get_weights(self):
        if self._params is None:
            raise _lib.Y3Error("model has no weights yet")
        out = []
        for p in self._params:
            out += p.as_list()
        return out

    def init_weights(self, kind="keras", seed=0):
        """Random-init weights of the architecture: 'keras' = what a freshly built reference model holds,
        'variance' = variance-preserving init used by the tolerance tests."""
        if kind == "keras":
            self.set_params(weights_mod.init_keras_default(self.conv_shapes, seed))
        elif kind == "variance":
            g = self.graph
            feeds_shortcut = {l.src0 for l in g.layers if l.op == _lib.OP_SHORTCUT}
            res = [ci for ci, li in enumerate(g.conv_layers) if (li + 1) in feeds_shortcut]
            self.set_params(weights_mod.init_variance_preserving(self.conv_shapes, seed, nclasses=self.nclasses,
                                                                 residual_convs=res))
        else:
            raise ValueError(kind)
        return self

    def load_weights(self, path):
        """Darknet ``.weights`` (reference convert.py), an ``.npz`` written by ``save_weights``, or a TensorFlow
        checkpoint prefix as written by the reference's ``model.save_weights(prefix)`` (train.py:93-104) and read by
        ``model.load_weights(prefix)`` (inference.py:102): ``<prefix>.index`` + ``<prefix>.data-00000-of-00001``, parsed
        without TensorFlow (yolo_v3_tf2_b200/tf_checkpoint.py)."""
        from .. import tf_checkpoint
        path = str(path)
        if path.endswith(".weights"):
            self.set_params(weights_mod.read_darknet_weights(path, self.conv_shapes))
        elif path.endswith(".npz"):
            z = np.load(path)
            self.set_weights([z[f"arr_{i}"] for i in range(len(z.files))])
        elif tf_checkpoint.is_checkpoint(path):
            self.set_params(tf_checkpoint.params_from_checkpoint(path, self.conv_shapes,
                                                                 graph_mod.keras_weight_slots(self.graph)))
        else:
            raise FileNotFoundError(f"{path}: neither a Darknet .weights file, an .npz, nor a TensorFlow checkpoint prefix "
                                    f"({path}.index not found)")
        return self

    def expect_partial(self):   # Keras load_weights(...).expect_partial() chaining (inference.py:102)
        return self

    def save_weights(self, path):
        """``.npz`` (Keras variable order) or, for any other name, a TensorFlow checkpoint with the object-based
        variable names the reference's Keras model would use (one sub-model per ``sub_models_configs`` entry)."""
        path = str(path)
        if path.endswith(".npz"):
            np.savez(path, *self.get_weights())
            return
        from .. import tf_checkpoint
        if self._params is None:
            raise _lib.Y3Error("model has no weights yet")
        names = tf_checkpoint.keras_variable_names(graph_mod.keras_weight_slots(self.graph), self.conv_shapes)
        tensors = {}
        for p, nm in zip(self._params, names):
            for key, arr in zip(nm, p.as_list()):
                tensors[key] = np.asarray(arr, np.float32)
        tf_checkpoint.write_checkpoint(path, tensors)

    # ---------------- execution ----------------
    def _upload(self, handle):
        lib = _lib.lib()
        for i, p in enumerate(self._params):
            keep = [np.ascontiguousarray(a, np.float32) if a is not None else None
                    for a in (p.kernel, p.bias, p.gamma, p.beta, p.mean, p.var)]
            args = [None if a is None else a.ctypes.data_as(C.c_void_p) for a in keep]
            _lib.check(lib.y3_net_load_conv(handle, i, *args, weights_mod.BN_EPS))

    def _net(self, H, W, B, device=None):
        ctx = _lib.context(device)
        key = (ctx.device, int(H), int(W))
        ent = self._nets.get(key)
        if ent is not None and ent["max_batch"] >= B:
            return ent
        lib = _lib.lib()
        if ent is not None:
            self._retired.append(ent)   # never freed while a captured graph may still replay on it
        h = C.c_void_p()
        _lib.check(lib.y3_net_create(ctx.handle, self._descs, len(self._descs), int(H), int(W), int(B),
                                     int(self.nclasses), C.byref(h)))
        shapes = []
        for k in range(lib.y3_net_num_outputs(h)):
            gh, gw, ch = C.c_int(), C.c_int(), C.c_int()
            _lib.check(lib.y3_net_output_shape(h, k, C.byref(gh), C.byref(gw), C.byref(ch)))
            shapes.append((gh.value, gw.value, ch.value))
        ent = {"handle": h, "max_batch": int(B), "ctx": ctx, "out_shapes": shapes}
        self._nets[key] = ent
        if self._params is not None:
            self._upload(h)
        return ent

    def plan(self, H, W, max_batch=1):
        """Planner output for (H, W) on a planning-only context -- works without a GPU (host-logic tests)."""
        lib = _lib.lib()
        ctx = _lib.Context(-1)
        h = C.c_void_p()
        _lib.check(lib.y3_net_create(ctx.handle, self._descs, len(self._descs), int(H), int(W), int(max_batch),
                                     int(self.nclasses), C.byref(h)))
        plans = (_lib.LayerPlan * len(self._descs))()
        _lib.check(lib.y3_net_get_plan(h, plans, len(self._descs)))
        out = {"arena_bytes": lib.y3_net_arena_bytes(h), "num_convs": lib.y3_net_num_convs(h),
               "layers": [{f: getattr(p, f) for f, _ in _lib.LayerPlan._fields_} for p in plans]}
        # launch-level view: how every kernel of a max_batch forward pass waits for its producers (layer chaining)
        nsteps = lib.y3_net_num_steps(h)
        chain = (_lib.ChainStep * nsteps)()
        _lib.check(lib.y3_net_chain_plan(h, int(max_batch), chain, nsteps))
        out["steps"] = [{f: getattr(c, f) for f, _ in _lib.ChainStep._fields_} for c in chain]
        lib.y3_net_destroy(h)
        ctx.close()
        return out

    def __call__(self, x, training=False, outs=None, padded=False):
        """x: [B, H, W, 3] NHWC (torch CUDA tensor; numpy / CPU tensors are copied to the GPU), either float32 in [0, 1]
        as ``inference.py:157-158`` feeds the Keras model, or **uint8**: the network then runs on ``float32(x) / 255``
        (the reference's ``resize(...) / 255``, core/load_tfrecords.py:46) with the division done inside the stem conv --
        bit-identical results, a quarter of the input bytes.
        ``padded=True`` (used by ``Detector``): each output is [B, gh, gw, P] with P = 3*(5+C) rounded up to a multiple
        of 4 floats and the logits of a pixel at its start -- the head convs then store through TMA and ``yolo_decode``
        reads the pitched layout directly."""
        if self._params is None:
            raise _lib.Y3Error("model has no weights: call load_weights / set_weights / init_weights first")
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x if x.dtype == np.uint8 else x.astype(np.float32, copy=False)))
        if x.dim() != 4 or x.shape[3] != 3:
            raise ValueError(f"input shape {tuple(x.shape)} is not [B, H, W, 3]")
        if not x.is_cuda:
            x = x.to(torch.device("cuda", _lib.context().device), non_blocking=True)
        u8 = x.dtype == torch.uint8
        x = x.contiguous() if u8 else x.contiguous().float()
        B, H, W, _ = x.shape
        ent = self._net(H, W, B, x.device.index)
        lib = _lib.lib()
        if padded:
            if outs is None:
                outs = [torch.empty((B, gh, gw, (ch + 3) // 4 * 4), dtype=torch.float32, device=x.device)
                        for gh, gw, ch in ent["out_shapes"]]
            pitch = (C.c_int * len(outs))(*[int(o.shape[3]) for o in outs])
        else:
            if outs is None:
                outs = [torch.empty((B, gh, gw, 3, ch // 3), dtype=torch.float32, device=x.device)
                        for gh, gw, ch in ent["out_shapes"]]
            pitch = None
        op = (C.c_void_p * len(outs))(*[o.data_ptr() for o in outs])
        if u8:
            _lib.check(lib.y3_net_forward_u8(ent["handle"], _lib.ptr(x), int(B), op, pitch, len(outs), _lib.stream_ptr()))
        elif padded:
            _lib.check(lib.y3_net_forward_pitched(ent["handle"], _lib.ptr(x), int(B), op, pitch, len(outs),
                                                  _lib.stream_ptr()))
        else:
            _lib.check(lib.y3_net_forward(ent["handle"], _lib.ptr(x), int(B), op, len(outs), _lib.stream_ptr()))
        return outs

    def read_layer(self, layer, x_shape, device=None):
        """Activation tensor that graph layer ``layer`` materialised during the last forward pass on an input of shape
        ``x_shape`` = (B, H, W, 3): float32 numpy [B, h, w, C] (bf16 values).  Parity aid -- the reference exposes
        intermediate tensors as Keras sub-model outputs (core/parse_model.py:279-314)."""
        B, H, W, _ = x_shape
        ent = self._net(H, W, B, device)
        plans = (_lib.LayerPlan * len(self._descs))()
        lib = _lib.lib()
        _lib.check(lib.y3_net_get_plan(ent["handle"], plans, len(self._descs)))
        pl = plans[layer]
        raw = np.empty((B, pl.H, pl.W, pl.C), np.uint16)
        _lib.check(lib.y3_net_read_layer(ent["handle"], int(layer), int(B), raw.ctypes.data_as(C.c_void_p)))
        return (raw.astype(np.uint32) << 16).view(np.float32)

    def profile_layers(self, x):
        """Per-kernel device times of one forward pass: list of (layer index, ms).  Profiling aid, not part of the
        reference surface."""
        B, H, W, _ = x.shape
        ent = self._net(H, W, B, x.device.index)
        lib = _lib.lib()
        n = lib.y3_net_num_steps(ent["handle"])
        outs = [torch.empty((B, gh, gw, 3, ch // 3), dtype=torch.float32, device=x.device)
                for gh, gw, ch in ent["out_shapes"]]
        op = (C.c_void_p * len(outs))(*[o.data_ptr() for o in outs])
        ms = np.zeros(n, np.float32)
        layer = np.zeros(n, np.int32)
        _lib.check(lib.y3_net_forward_timed(ent["handle"], _lib.ptr(x.contiguous().float()), int(B), op, len(outs),
                                            _lib.stream_ptr(), ms.ctypes.data_as(C.c_void_p),
                                            layer.ctypes.data_as(C.c_void_p), n))
        return list(zip(layer.tolist(), ms.tolist()))

    def predict(self, x, batch_size=32, **kwargs):
        """Keras ``model.predict``: batches of ``batch_size`` (Keras default 32), numpy arrays out."""
        if isinstance(x, torch.Tensor):
            x = x.detach().cpu().numpy()
        x = np.ascontiguousarray(x) if x.dtype == np.uint8 else np.ascontiguousarray(x, dtype=np.float32)
        chunks = []
        for i in range(0, x.shape[0], batch_size):
            chunks.append([o.cpu().numpy() for o in self(x[i:i + batch_size])])
        return [np.concatenate([c[k] for c in chunks], axis=0) for k in range(len(chunks[0]))]

    def summary(self, print_fn=print):
        print_fn(f'Model: "{self.name}"  ({len(self.graph.layers)} layers, {len(self.conv_shapes)} convs, '
                 f'{sum(k * k * ci * co for k, ci, co, _ in self.conv_shapes):,} conv weights)')
        for i, l in enumerate(self.graph.layers):
            print_fn(f"{i:4d} {l.sub_model:9s} {graph_mod.OP_NAMES[l.op]:9s} src={l.src0},{l.src1} "
                     f"k={l.ksize} s={l.stride} filters={l.filters} bn={l.batch_normalize} act={l.activation}")

    def close(self):
        for ent in list(self._nets.values()) + self._retired:
            _lib.lib().y3_net_destroy(ent["handle"])
        self._nets = {}
        self._retired = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ParseModel:
    """Same entry points as reference core/parse_model.py:10 (``build_model`` :279-314, ``create_model`` :316-322)."""

    def build_model(self, model_inputs, sub_models_configs, output_stage='head', decay_factor=0, nclasses=0,
                    search_dirs=(), layer_lists=None, **kwargs):
        """``model_inputs`` (a Keras ``Input`` in the reference) is accepted and ignored: H and W come from the tensor
        the model is called on.  ``decay_factor`` only matters for training and is ignored."""
        g = graph_mod.build_graph(sub_models_configs, output_stage, nclasses, search_dirs=search_dirs,
                                  layer_lists=layer_lists)
        if not g.outputs:
            raise ValueError(f"no sub-model name contains output_stage={output_stage!r}")
        return Y3Model(g)

    def create_model(self, nclasses, model_config_file):
        with open(model_config_file, 'r') as _stream:
            model_config = yaml.safe_load(_stream)
        if "sub_models" in model_config:   # legacy monolithic config/yolov3_model.yaml
            return Y3Model(graph_mod.build_graph_legacy(model_config, nclasses))
        return Y3Model(graph_mod.load_model_config(model_config_file, nclasses))

    @staticmethod
    def builtin_yolov3_tiny(nclasses):
        """YOLOv3-tiny from the built-in description (identical graph to config/models/yolov3_tiny/model.yaml)."""
        from .. import configs
        model, files = configs.yolov3_tiny_config()
        return ParseModel().build_model(None, model["sub_models_configs"], model["output_stage"], nclasses=nclasses,
                                        layer_lists=files)

    @staticmethod
    def builtin_yolov3(nclasses, thin_heads=False):
        """Darknet-53 YOLOv3 from the built-in description (identical graph to config/models/yolov3/model.yaml, or to
        config/models/yolov3/model_thin_heads.yaml with ``thin_heads=True``)."""
        from .. import configs
        model, files = configs.yolov3_config(thin_heads)
        return ParseModel().build_model(None, model["sub_models_configs"], model["output_stage"], nclasses=nclasses,
                                        layer_lists=files)
