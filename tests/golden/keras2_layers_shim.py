"""Eager stand-ins for the `keras.layers` classes (and keras_applications' `_depthwise_conv_block`) that the graph
BUILDERS of the reference's myolo/model.py call -- TEST INFRASTRUCTURE for make_reference_graph_fixtures.py.

With these, conv_block (42-52), mobilenet_graph (55-79), yolo_branch_graph (249-278) and build_mask_graph (668-715) run
unmodified from /root/reference and what they wire -- which layer follows which, filters and strides per block, paddings,
biases, which BatchNormalization follows the learning phase and which is called with training=False, activations, the
TimeDistributed wrapping -- becomes golden vectors.  The layers themselves are third-party (Keras 2.x on TensorFlow 1.x)
and restated from their documented behaviour:
  Conv2D / DepthwiseConv2D  NHWC, kernels HWIO / [kh,kw,C,1]; 'same' = TensorFlow SAME (total pad = max((ceil(n/s)-1)*s+k-n, 0),
                            the smaller half first); 'valid' = no padding
  Conv2DTranspose           kernel [kh,kw,Cout,Cin], 'valid', output = (n-1)*s + k
  BatchNormalization        epsilon 1e-3; learning phase 1 (or training=True): statistics of the batch over every axis but the
                            last, biased variance; otherwise moving statistics.  `training=None` follows the learning phase.
  TimeDistributed           batch size unknown at graph-build time (KL.Input): [B, T, ...] is reshaped to [B*T, ...], the
                            wrapped layer applied once, and the result reshaped back; the wrapper's name scopes the variables
  _depthwise_conv_block     keras_applications 1.0.4-1.0.6 as pinned by the reference's GraphDef (tests/golden/graph_fixture.json):
                            ZeroPadding2D((1,1)) for every stride, depthwise 3x3 VALID, BN, relu6, pointwise 1x1 SAME, BN, relu6
Weights come from WEIGHTS[<layer name>/<variable>] (numpy float32), the Keras variable names of the graph fixture."""
import numpy as np
import torch
import torch.nn.functional as F

import tf1_numpy_shim as tfs

WEIGHTS = {}
STATE = {"learning_phase": 1, "dtype": np.float64}     # layers compute in float64: the fixture then pins the fp64 oracle to 1e-9
BN_EPS = 1e-3


def _np(x):
    return x.a if isinstance(x, tfs.T) else np.asarray(x)


def _w(name):
    return torch.from_numpy(np.ascontiguousarray(WEIGHTS[name], dtype=STATE["dtype"]))


def _same_pad(n, k, s):
    total = max((-(-n // s) - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def _act(x, activation):
    if activation is None:
        return x
    if callable(activation):
        return _np(activation(tfs.T(x)))
    return {"relu": lambda v: np.maximum(v, 0).astype(v.dtype),
            "sigmoid": lambda v: (1 / (1 + np.exp(-v))).astype(v.dtype)}[activation](x)


class Layer(object):
    def __init__(self, name=None, **kwargs):
        self.name = name

    def __call__(self, inputs, **kwargs):
        return self.call(inputs, **kwargs)


class ZeroPadding2D(Layer):
    def __init__(self, padding=(1, 1), name=None):
        Layer.__init__(self, name)
        self.p = padding

    def call(self, x):
        ph, pw = self.p
        return tfs.T(np.pad(_np(x), [(0, 0), (ph, ph), (pw, pw), (0, 0)]))


class Conv2D(Layer):
    def __init__(self, filters, kernel_size, strides=(1, 1), padding="valid", use_bias=True, activation=None, name=None):
        Layer.__init__(self, name)
        self.filters, self.k = filters, kernel_size if isinstance(kernel_size, (tuple, list)) else (kernel_size, kernel_size)
        self.s = strides if isinstance(strides, (tuple, list)) else (strides, strides)
        self.padding, self.use_bias, self.activation = padding.lower(), use_bias, activation      # Keras normalises the case

    def call(self, x):
        x = torch.from_numpy(np.ascontiguousarray(_np(x), dtype=STATE["dtype"])).permute(0, 3, 1, 2)
        w = _w(self.name + "/kernel")
        assert tuple(w.shape[:2]) == tuple(self.k) and w.shape[3] == self.filters and w.shape[2] == x.shape[1], (self.name, w.shape)
        if self.padding == "same":
            (t, b), (l, r) = _same_pad(x.shape[2], self.k[0], self.s[0]), _same_pad(x.shape[3], self.k[1], self.s[1])
            x = F.pad(x, (l, r, t, b))
        y = F.conv2d(x, w.permute(3, 2, 0, 1).contiguous(), stride=tuple(self.s)).permute(0, 2, 3, 1)
        if self.use_bias:
            y = y + _w(self.name + "/bias")
        return tfs.T(_act(y.contiguous().numpy(), self.activation))


class DepthwiseConv2D(Layer):
    def __init__(self, kernel_size, strides=(1, 1), padding="valid", depth_multiplier=1, use_bias=True, name=None):
        Layer.__init__(self, name)
        assert depth_multiplier == 1 and padding == "valid" and not use_bias
        self.k, self.s = kernel_size, strides

    def call(self, x):
        x = torch.from_numpy(np.ascontiguousarray(_np(x), dtype=STATE["dtype"])).permute(0, 3, 1, 2)
        w = _w(self.name + "/depthwise_kernel")                       # [kh, kw, C, 1]
        assert tuple(w.shape) == (self.k[0], self.k[1], x.shape[1], 1), (self.name, w.shape)
        y = F.conv2d(x, w.permute(2, 3, 0, 1).contiguous(), stride=tuple(self.s), groups=x.shape[1])
        return tfs.T(y.permute(0, 2, 3, 1).contiguous().numpy())


class Conv2DTranspose(Layer):
    def __init__(self, filters, kernel_size, strides=1, padding="valid", activation=None, name=None):
        Layer.__init__(self, name)
        assert padding == "valid"
        self.filters, self.k, self.s, self.activation = filters, kernel_size, strides, activation

    def call(self, x):
        x = torch.from_numpy(np.ascontiguousarray(_np(x), dtype=STATE["dtype"])).permute(0, 3, 1, 2)
        w = _w(self.name + "/kernel")                                 # [kh, kw, Cout, Cin]
        assert tuple(w.shape) == (self.k[0], self.k[1], self.filters, x.shape[1]), (self.name, w.shape)
        y = F.conv_transpose2d(x, w.permute(3, 2, 0, 1).contiguous(), stride=self.s).permute(0, 2, 3, 1)
        y = y + _w(self.name + "/bias")
        return tfs.T(_act(y.contiguous().numpy(), self.activation))


class BatchNormalization(Layer):
    def __init__(self, axis=-1, name=None):
        Layer.__init__(self, name)
        assert axis in (-1, 3)

    def call(self, x, training=None):
        x = _np(x)
        g, b = WEIGHTS[self.name + "/gamma"], WEIGHTS[self.name + "/beta"]
        if training is None:
            training = STATE["learning_phase"]
        if training:
            red = tuple(np.arange(x.ndim - 1))
            mean = x.mean(axis=red, dtype=np.float64)
            var = ((x.astype(np.float64) - mean) ** 2).mean(axis=red)
        else:
            mean, var = WEIGHTS[self.name + "/moving_mean"].astype(np.float64), WEIGHTS[self.name + "/moving_variance"].astype(np.float64)
        USED.append((self.name, "batch" if training else "moving"))
        y = (x.astype(np.float64) - mean) / np.sqrt(var + BN_EPS) * g + b
        return tfs.T(y.astype(STATE["dtype"]))


USED = []          # (BatchNormalization name, "batch" | "moving") in call order: which statistics each layer used


class Activation(Layer):
    def __init__(self, activation, name=None):
        Layer.__init__(self, name)
        self.activation = activation

    def call(self, x):
        return tfs.T(_act(_np(x), self.activation))


class Reshape(Layer):
    def __init__(self, target_shape, name=None):
        Layer.__init__(self, name)
        self.shape = tuple(target_shape)

    def call(self, x):
        x = _np(x)
        return tfs.T(x.reshape((x.shape[0],) + self.shape))


class TimeDistributed(Layer):
    def __init__(self, layer, name=None):
        Layer.__init__(self, name)
        self.layer = layer
        layer.name = name                                             # the wrapper's name scopes the variables

    def call(self, x, training=None):
        x = _np(x)
        flat = x.reshape((-1,) + x.shape[2:])
        y = _np(self.layer.call(flat, training=training) if isinstance(self.layer, BatchNormalization) else self.layer.call(flat))
        return tfs.T(y.reshape(x.shape[:2] + y.shape[1:]))


class _Backend(object):
    @staticmethod
    def relu(x, max_value=None):
        a = _np(x)
        a = np.maximum(a, 0).astype(a.dtype)
        return tfs.T(a if max_value is None else np.minimum(a, max_value).astype(a.dtype))

    @staticmethod
    def image_data_format():
        return "channels_last"


backend = _Backend()


def depthwise_conv_block(inputs, pointwise_conv_filters, alpha, depth_multiplier=1, strides=(1, 1), block_id=1):
    """keras_applications.mobilenet._depthwise_conv_block, the variant the reference's shipped graph was built with."""
    relu6 = lambda x: backend.relu(x, max_value=6)                    # noqa: E731
    x = ZeroPadding2D((1, 1), name="conv_pad_%d" % block_id)(inputs)
    x = DepthwiseConv2D((3, 3), padding="valid", depth_multiplier=depth_multiplier, strides=strides, use_bias=False,
                        name="conv_dw_%d" % block_id)(x)
    x = BatchNormalization(axis=-1, name="conv_dw_%d_bn" % block_id)(x)
    x = Activation(relu6, name="conv_dw_%d_relu" % block_id)(x)
    x = Conv2D(int(pointwise_conv_filters * alpha), (1, 1), padding="same", use_bias=False, strides=(1, 1),
               name="conv_pw_%d" % block_id)(x)
    x = BatchNormalization(axis=-1, name="conv_pw_%d_bn" % block_id)(x)
    return Activation(relu6, name="conv_pw_%d_relu" % block_id)(x)


# ---- the functional-API names MaskYOLO.build (787-941) uses, evaluated eagerly
FEEDS = {}         # Input name -> numpy array


def Input(shape=None, name=None, dtype=None, **kwargs):
    """A placeholder is bound to its value at once (the graph is evaluated while it is being built)."""
    return tfs.T(FEEDS[name])


class Lambda(Layer):
    def __init__(self, function, name=None, **kwargs):
        Layer.__init__(self, name)
        self.function = function

    def call(self, x):
        return self.function(x)


class Model(object):
    """KM.Model(inputs, outputs): keeps the evaluated outputs.  Calling the model on tensors (the nested yolo_model,
    model.py:851-852) returns them, after checking that the tensors are the ones its placeholders were bound to."""

    def __init__(self, inputs, outputs, name=None):
        self.inputs, self.outputs, self.name, self.layers = list(inputs), outputs, name, []

    def __call__(self, inputs):
        assert len(inputs) == len(self.inputs)
        for a, b in zip(inputs, self.inputs):
            assert np.array_equal(_np(a), _np(b)), "nested model called on a tensor other than its bound placeholder"
        return self.outputs

    def summary(self):
        return "%s: %d outputs" % (self.name, len(self.outputs) if isinstance(self.outputs, (list, tuple)) else 1)
