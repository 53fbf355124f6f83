"""A minimal EAGER stand-in for the `tensorflow` 1.x / `keras.backend` names that the loss / decode / target
functions of the reference's myolo/model.py use -- TEST INFRASTRUCTURE for tests/golden/make_reference_graph_fixtures.py.

Purpose: TensorFlow 1.x and Keras 2.x cannot be installed here, but the *formulas* of the reference (which tensor is
multiplied with which, the masks, the normalisers, the ordering of positives and negatives, the x/y swap of the ROI boxes,
the padding) live in the reference's own Python source.  With the primitive ops below supplied over numpy, that source
runs unmodified, line by line, and its outputs become golden vectors.  What this pins is therefore the reference's
COMPOSITION; the primitives themselves (sigmoid, reductions, `crop_and_resize`, `top_k`, Keras `binary_crossentropy`)
are restated from TensorFlow's / Keras' documented behaviour, in float32 like the graph the reference builds.

Semantics kept from TF: tensors are float32 / int32 / int64 / bool; a Python or numpy operand combined with a tensor is
converted to the TENSOR's dtype (tf.convert_to_tensor with a dtype hint), so `tensor * np.reshape(ANCHORS, ...)` stays
float32; `tf.round` is half-to-even; `tf.where(cond)` returns int64 coordinates in row-major order; `tf.nn.top_k` is
stable (lower index first among equals)."""
import builtins
import contextlib
import types

import numpy as np

float32, int32, int64, int8, bool_ = np.float32, np.int32, np.int64, np.int8, np.bool_


class T(object):
    """Eager tensor: a numpy array with TensorFlow's operand-conversion rule."""
    __array_ufunc__ = None                      # ndarray <op> T defers to T.__r<op>__

    def __init__(self, a):
        self.a = a.a if isinstance(a, T) else np.asarray(a)

    # -- conversion of the other operand
    def _o(self, other):
        if isinstance(other, T):
            return other.a
        return np.asarray(other).astype(self.a.dtype)

    shape = property(lambda self: tuple(int(s) for s in self.a.shape))
    dtype = property(lambda self: self.a.dtype)

    def __add__(self, o): return T(self.a + self._o(o))
    def __radd__(self, o): return T(self._o(o) + self.a)
    def __sub__(self, o): return T(self.a - self._o(o))
    def __rsub__(self, o): return T(self._o(o) - self.a)
    def __mul__(self, o): return T(self.a * self._o(o))
    def __rmul__(self, o): return T(self._o(o) * self.a)
    def __truediv__(self, o): return T(self.a / self._o(o))
    def __rtruediv__(self, o): return T(self._o(o) / self.a)
    def __neg__(self): return T(-self.a)
    def __lt__(self, o): return T(self.a < self._o(o))
    def __le__(self, o): return T(self.a <= self._o(o))
    def __gt__(self, o): return T(self.a > self._o(o))
    def __ge__(self, o): return T(self.a >= self._o(o))

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        idx = tuple(int(i) if isinstance(i, T) else i for i in idx)
        return T(self.a[idx])

    def __len__(self): return self.a.shape[0]
    def __iter__(self): return (T(x) for x in self.a)
    def __int__(self): return int(self.a)
    def __index__(self): return int(self.a)
    def __float__(self): return float(self.a)
    def __bool__(self): return builtins.bool(self.a)
    def __repr__(self): return "T(%r)" % (self.a,)


def _a(x, dtype=None):
    a = x.a if isinstance(x, T) else np.asarray(x)
    return a if dtype is None else a.astype(dtype)


def _ints(seq):
    if isinstance(seq, T):
        seq = seq.a.tolist()
    return [int(v) for v in (seq if isinstance(seq, (list, tuple)) else [seq])]


def _like(x, y):
    """binary op operands: a non-tensor takes the dtype of the tensor it meets"""
    if isinstance(x, T) and not isinstance(y, T):
        return x.a, np.asarray(y).astype(x.a.dtype)
    if isinstance(y, T) and not isinstance(x, T):
        return np.asarray(x).astype(y.a.dtype), y.a
    return _a(x), _a(y)


def constant(v, dtype=None, name=None):
    a = np.asarray(v)
    if dtype is None:
        dtype = np.float32 if a.dtype.kind == "f" else (np.int32 if a.dtype.kind in "iu" else a.dtype)
    return T(a.astype(dtype))


def shape(x, name=None): return T(np.asarray(_a(x).shape, dtype=np.int32))
def size(x): return T(np.int32(_a(x).size))
def zeros(shp, dtype=np.float32): return T(np.zeros(_ints(shp), dtype=dtype))
def ones_like(x): return T(np.ones_like(_a(x)))
def to_float(x): return T(_a(x).astype(np.float32))
def cast(x, dtype, name=None): return T(_a(x).astype(dtype))
def identity(x, name=None): return x if isinstance(x, T) else T(x)
def stop_gradient(x): return x


def range(start, limit=None, delta=1):       # noqa: A001 - the tf name
    if limit is None:
        start, limit = 0, start
    return T(np.arange(int(start), int(limit), int(delta), dtype=np.int32))


def reshape(x, shp, name=None): return T(_a(x).reshape(_ints(shp)))
def tile(x, multiples): return T(np.tile(_a(x), _ints(multiples)))
def transpose(x, perm=None): return T(np.transpose(_a(x), perm))
def concat(values, axis, name=None): return T(np.concatenate([_a(v) for v in values], axis=axis))
def stack(values, axis=0, name=None): return T(np.stack([_a(v) for v in values], axis=axis))
def expand_dims(x, axis=None, dim=None): return T(np.expand_dims(_a(x), axis if axis is not None else dim))
def squeeze(x, axis=None): return T(np.squeeze(_a(x), axis=axis))
def split(x, num, axis=0): return [T(p) for p in np.split(_a(x), num, axis=axis)]


def pad(x, paddings, name=None):
    return T(np.pad(_a(x), [tuple(_ints(list(p))) for p in paddings], mode="constant"))


def sigmoid(x):
    a = _a(x)
    return T((np.float32(1) / (np.float32(1) + np.exp(-a))).astype(a.dtype))


def exp(x): return T(np.exp(_a(x)))
def log(x): return T(np.log(_a(x, np.float32) if not isinstance(x, T) else _a(x)))
def sqrt(x): return T(np.sqrt(_a(x)))
def square(x): return T(np.square(_a(x)))
def abs(x): return T(np.abs(_a(x)))          # noqa: A001
def round(x): return T(np.round(_a(x)))      # noqa: A001  (numpy rounds half to even, like tf.round)
def maximum(x, y): return T(np.maximum(*_like(x, y)))
def minimum(x, y): return T(np.minimum(*_like(x, y)))
def truediv(x, y): return T(np.true_divide(*_like(x, y)))
def divide(x, y): return T(np.true_divide(*_like(x, y)))
def less(x, y): return T(np.less(*_like(x, y)))
def greater(x, y): return T(np.greater(*_like(x, y)))
def equal(x, y): return T(np.equal(*_like(x, y)))
def reduce_sum(x, axis=None): return T(np.sum(_a(x), axis=axis, dtype=_a(x).dtype if _a(x).dtype != np.bool_ else None))
def reduce_max(x, axis=None):
    a = _a(x)
    if a.size == 0 and axis is not None:      # TF reduces an empty axis to the lowest finite value of the dtype
        shp = [s for k, s in enumerate(a.shape) if k != (axis % a.ndim)]
        return T(np.full(shp, np.finfo(a.dtype).min if a.dtype.kind == "f" else np.iinfo(a.dtype).min, dtype=a.dtype))
    return T(np.max(a, axis=axis))
def argmax(x, axis=None): return T(np.argmax(_a(x), axis=axis).astype(np.int64))


def gather(params, indices, axis=0, name=None):
    return T(np.take(_a(params), _a(indices).astype(np.int64), axis=axis))


def gather_nd(params, indices):
    idx = _a(indices).astype(np.int64)
    return T(_a(params)[tuple(idx[:, k] for k in np.arange(idx.shape[1]))])


def boolean_mask(x, mask, name=None): return T(_a(x)[_a(mask).astype(np.bool_)])
def where(cond): return T(np.argwhere(_a(cond)).astype(np.int64))


class Variable(T):
    def __init__(self, v):
        T.__init__(self, np.float32(v))


def assign_add(var, v):
    var.a = (var.a + np.asarray(_a(v)).astype(var.a.dtype))
    return var


def cond(pred, true_fn=None, false_fn=None, **kw):
    return true_fn() if builtins.bool(_a(pred)) else false_fn()


def Print(x, data, message=None, summarize=None): return x
def Assert(condition, data, name=None): return None
def control_dependencies(deps): return contextlib.nullcontext()


def _sparse_softmax_cross_entropy_with_logits(labels=None, logits=None):
    z = _a(logits)
    m = z.max(axis=-1, keepdims=True)
    lse = np.log(np.exp(z - m).sum(axis=-1, keepdims=True, dtype=z.dtype)) + m
    picked = np.take_along_axis(z, _a(labels).astype(np.int64)[..., None], axis=-1)
    return T((lse - picked)[..., 0])


def _top_k(x, k=1, sorted=True):             # noqa: A002
    a = _a(x)
    order = np.argsort(-a.astype(np.int64) if a.dtype.kind in "iu" else -a, kind="stable")[:int(k)]
    return types.SimpleNamespace(values=T(a[order]), indices=T(order.astype(np.int32)))


def _crop_and_resize(image, boxes, box_ind, crop_size, method="bilinear", extrapolation_value=0, name=None):
    """tf.image.crop_and_resize (CropAndResize CPU kernel), float32.  boxes are [y1, x1, y2, x2] normalised to
    [0, 1] <-> [0, size-1]; one bilinear sample per output element; samples outside [0, size-1] take the
    extrapolation value; value = top + (bottom - top) * y_lerp with top = tl + (tr - tl) * x_lerp."""
    img, bx, bi = _a(image).astype(np.float32), _a(boxes).astype(np.float32), _a(box_ind).astype(np.int64)
    ch, cw = _ints(crop_size)
    H, W, C = img.shape[1], img.shape[2], img.shape[3]
    f = np.float32
    out = np.full((bx.shape[0], ch, cw, C), f(extrapolation_value), dtype=np.float32)
    for n in np.arange(bx.shape[0]):
        y1, x1, y2, x2 = bx[n]
        hs = (y2 - y1) * f(H - 1) / f(ch - 1) if ch > 1 else f(0)
        ws = (x2 - x1) * f(W - 1) / f(cw - 1) if cw > 1 else f(0)
        for y in np.arange(ch):
            in_y = y1 * f(H - 1) + f(y) * hs if ch > 1 else f(0.5) * (y1 + y2) * f(H - 1)
            if in_y < 0 or in_y > H - 1:
                continue
            top, bot = int(np.floor(in_y)), int(np.ceil(in_y))
            yl = f(in_y - f(top))
            xs = (x1 * f(W - 1) + np.arange(cw, dtype=np.float32) * ws) if cw > 1 else np.full(1, f(0.5) * (x1 + x2) * f(W - 1))
            ok = (xs >= 0) & (xs <= W - 1)
            xl_i = np.floor(xs[ok]).astype(np.int64)
            xr_i = np.ceil(xs[ok]).astype(np.int64)
            xlerp = (xs[ok] - xl_i.astype(np.float32))[:, None]
            tl, tr = img[bi[n], top, xl_i], img[bi[n], top, xr_i]
            bl, br = img[bi[n], bot, xl_i], img[bi[n], bot, xr_i]
            t = tl + (tr - tl) * xlerp
            b = bl + (br - bl) * xlerp
            out[n, y, ok] = t + (b - t) * yl
    return T(out)


nn = types.SimpleNamespace(sparse_softmax_cross_entropy_with_logits=_sparse_softmax_cross_entropy_with_logits, top_k=_top_k)
image = types.SimpleNamespace(crop_and_resize=_crop_and_resize)
__version__ = "1.12.0"
bool = bool_                                  # noqa: A001  tf.bool


# ---- keras.backend names (Keras 2.1/2.2, TensorFlow backend)
def k_reshape(x, shp): return reshape(x, shp)
def k_switch(condition, then_expression, else_expression): return then_expression if builtins.bool(_a(condition)) else else_expression
def k_mean(x, axis=None): return T(np.mean(_a(x), axis=axis, dtype=np.float32))


def k_binary_crossentropy(target, output, from_logits=False):
    """clip to [1e-7, 1-1e-7], logit, tf.nn.sigmoid_cross_entropy_with_logits = max(x,0) - x*z + log1p(exp(-|x|))"""
    z, p = _a(target).astype(np.float32), _a(output).astype(np.float32)
    if not from_logits:
        eps = np.float32(1e-7)
        p = np.clip(p, eps, np.float32(1) - eps)
        p = np.log(p / (np.float32(1) - p))
    return T(np.maximum(p, 0) - p * z + np.log1p(np.exp(-np.abs(p))))
