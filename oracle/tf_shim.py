"""TEST INFRASTRUCTURE — a torch-backed stand-in for the ~70 TensorFlow 2.4 symbols that the reference's transformer files call
(viewformer/models/migt.py, models/branching_attention.py, models/utils.py, utils/tensorflow.py, utils/geometry_tf.py,
utils/metrics.py, utils/schedules.py, evaluate/evaluate_transformer*.py).

TensorFlow 2.4.1 has no cp312 wheel and there is no network, so the reference transformer cannot run as shipped.  With this module
installed as ``tensorflow`` in ``sys.modules`` (``install()``), the reference's OWN source files are executed unmodified from
/root/reference (oracle/ref_loader.py::load_reference_migt): every line of model wiring — embeddings, the three streams, the branching
attention masks, the pose head, the losses, train_step's GradientTape — is the reference's; only the leaf tensor ops are answered here,
each one a few lines of torch restating the documented TensorFlow semantics (tf.split with an int = number of parts, tf.repeat =
element-wise repeat, tf.shape = int32 vector, Keras LayerNormalization over the last axis, exact-erf gelu, clip_by_norm, ...).
That pins the restatement in oracle/migt_oracle.py to the reference's code rather than to my reading of it; what is NOT pinned is
TensorFlow's own floating-point evaluation order (torch's CPU kernels are used instead) — irrelevant at the 1e-5 bars of the tests.

Never imported by the product (viewformer_b200/); container-only like /root/reference itself.
"""
import contextlib
import re
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

_DT = {"float32": torch.float32, "float64": torch.float64, "float16": torch.float16, "int32": torch.int32, "int64": torch.int64,
       "bool": torch.bool, "uint8": torch.uint8}


class _VarDType:
    """dtype of a shim tensor: compares / hashes like the torch dtype and carries tf.DType's ``base_dtype`` (the reference's optimizer
    keys its per-dtype state by ``var.dtype.base_dtype``, models/utils.py:515) and ``as_numpy_dtype`` (utils/geometry_tf.py:38)."""

    def __init__(self, d):
        self.base_dtype = d
        self.is_floating_point = d.is_floating_point

    def __getattr__(self, name):                       # is_complex, itemsize, is_signed ... answered by the torch dtype
        return getattr(self.base_dtype, name)

    def as_numpy_dtype(self, *a):
        return np.dtype(str(self.base_dtype).replace("torch.", "")).type(*a)

    @property
    def max(self):
        return torch.finfo(self.base_dtype).max if self.is_floating_point else torch.iinfo(self.base_dtype).max

    def __eq__(self, other):
        return self.base_dtype == (other.base_dtype if isinstance(other, _VarDType) else other)

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self.base_dtype)

    def __repr__(self):
        return repr(self.base_dtype)


def _dtype(d):
    if isinstance(d, _VarDType):
        return d.base_dtype
    if d is None or isinstance(d, torch.dtype):
        return d
    if isinstance(d, str):
        return _DT[d]
    raise TypeError(f"dtype {d!r}")


class TensorShape(tuple):
    """x.shape of a shim tensor: a tuple with TensorFlow's ``as_list`` / ``rank``; ``TensorShape(None)`` is the unknown shape."""
    _unknown = False

    def __new__(cls, dims=None):
        self = super().__new__(cls, () if dims is None else tuple(int(d) for d in dims))
        self._unknown = dims is None
        return self

    def as_list(self):
        return list(self)

    @property
    def rank(self):
        return None if self._unknown else len(self)

    def __eq__(self, other):
        if isinstance(other, TensorShape) and (self._unknown or other._unknown):
            return self._unknown and other._unknown
        return tuple(self) == tuple(other)

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = tuple.__hash__


class Tensor(torch.Tensor):
    """torch.Tensor whose ``.shape`` answers like a tf.Tensor's.  ``__module__`` starts with 'tensorflow' because the reference's
    schedules dispatch on it (utils/schedules.py:33-36)."""

    @property
    def shape(self):
        return TensorShape(torch.Tensor.size(self))

    @property
    def dtype(self):
        return _VarDType(torch.Tensor.dtype.__get__(self))

    def numpy(self):
        return torch.Tensor.numpy(self.detach().as_subclass(torch.Tensor))


Tensor.__module__ = "tensorflow.python.framework.ops"


def _t(x, dtype=None):
    """anything -> shim Tensor"""
    dtype = _dtype(dtype)
    if isinstance(x, torch.Tensor):
        out = x if dtype is None or _dtype(x.dtype) == dtype else x.to(dtype)
    else:
        if isinstance(x, (list, tuple)) and any(isinstance(e, torch.Tensor) for e in x):
            x = [e.item() if isinstance(e, torch.Tensor) else e for e in x]
        if isinstance(x, np.ndarray) and dtype is None:
            out = torch.from_numpy(np.ascontiguousarray(x))          # numpy arrays keep their dtype
            return out.as_subclass(Tensor)
        if dtype is None:                                            # python numbers / lists: TensorFlow's defaults float32 / int32
            a = np.asarray(x)
            dtype = torch.float32 if a.dtype.kind == "f" else (torch.int32 if a.dtype.kind in "iu" else (torch.bool if a.dtype.kind == "b" else None))
        out = torch.as_tensor(x, dtype=dtype)
    return out if isinstance(out, Tensor) else out.as_subclass(Tensor)


def _wrap(x):
    if isinstance(x, torch.Tensor):
        return x if isinstance(x, Tensor) else x.as_subclass(Tensor)
    if isinstance(x, np.ndarray):
        return _t(torch.from_numpy(x))
    if isinstance(x, dict):
        return {k: _wrap(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_wrap(v) for v in x)
    return x


def _ints(shape):
    """A TensorFlow shape argument (python ints, 0-d tensors, a 1-D tensor, a scalar meaning a 1-D shape) -> list of ints"""
    if isinstance(shape, torch.Tensor):
        return [int(v) for v in shape.reshape(-1).tolist()]
    if isinstance(shape, (int, np.integer)):
        return [int(shape)]
    return [int(v) for v in shape]


def _axes(axis):
    if axis is None:
        return None
    if isinstance(axis, (list, tuple)):
        return [int(a) for a in axis]
    return int(axis)


class Variable(Tensor):
    """tf.Variable: a leaf tensor with a name and assign ops."""

    @staticmethod
    def make(value, name="Variable", trainable=True, dtype=None):
        v = _t(value, dtype).detach().clone().as_subclass(Variable)
        v.vname = name
        v.trainable = trainable
        if trainable and _dtype(v.dtype).is_floating_point:
            v.requires_grad_(True)
        return v

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        # ops on a variable produce ordinary tensors (not nameless Variables)
        with torch._C.DisableTorchFunctionSubclass():
            ret = func(*args, **(kwargs or {}))
        if func in torch.overrides.get_default_nowrap_functions():
            return ret
        return torch._tensor._convert(ret, Tensor)

    @property
    def name(self):
        return self.vname

    def assign(self, value, **kw):
        with torch.no_grad():
            self.copy_(torch.as_tensor(value))
        return self

    def assign_sub(self, value, **kw):
        with torch.no_grad():
            self.sub_(torch.as_tensor(value))
        return self

    def assign_add(self, value, **kw):
        with torch.no_grad():
            self.add_(torch.as_tensor(value))
        return self


Variable.__module__ = "tensorflow.python.ops.resource_variable_ops"


# ------------------------------------------------------------------------------------------------------------ name scopes
_SCOPE = []


@contextlib.contextmanager
def name_scope(name):
    _SCOPE.append(str(name))
    try:
        yield "/".join(_SCOPE) + "/"
    finally:
        _SCOPE.pop()


# ------------------------------------------------------------------------------------------------------------ initializers
class _Init:
    def __call__(self, shape, dtype=torch.float32):
        raise NotImplementedError


class TruncatedNormal(_Init):
    def __init__(self, mean=0.0, stddev=0.05, seed=None):       # NB positional argument 1 is the MEAN (Keras signature)
        self.mean, self.stddev = mean, stddev

    def __call__(self, shape, dtype=torch.float32):
        x = torch.empty(_ints(shape), dtype=dtype)
        torch.nn.init.trunc_normal_(x, mean=self.mean, std=self.stddev, a=self.mean - 2 * self.stddev, b=self.mean + 2 * self.stddev)
        return x


class _Zeros(_Init):
    def __call__(self, shape, dtype=torch.float32):
        return torch.zeros(_ints(shape), dtype=dtype)


class _Ones(_Init):
    def __call__(self, shape, dtype=torch.float32):
        return torch.ones(_ints(shape), dtype=dtype)


class _Constant(_Init):
    def __init__(self, value):
        self.value = value

    def __call__(self, shape, dtype=torch.float32):
        return torch.as_tensor(self.value, dtype=dtype).expand(_ints(shape)).clone()


def _get_initializer(i):
    if i is None:
        return TruncatedNormal(0.0, 0.05)
    if isinstance(i, str):
        return {"zeros": _Zeros(), "ones": _Ones()}[i]
    return i


# ------------------------------------------------------------------------------------------------------------ Keras layers / model
def _snake(name):
    s = re.sub("(.)([A-Z][a-z0-9]+)", r"\1_\2", name)
    return re.sub("([a-z0-9])([A-Z])", r"\1_\2", s).lower()


class Layer:
    """tf.keras.layers.Layer: lazily built on first call, variables named by the call-time scope path (``migt/h.0/attn/c_attn/weight:0``)."""

    def __init__(self, name=None, dtype=None, trainable=True, autocast=None, **kwargs):
        if kwargs:
            raise TypeError(f"unexpected Layer kwargs {sorted(kwargs)}")
        object.__setattr__(self, "_name", name or _snake(type(self).__name__))
        object.__setattr__(self, "_layer_dtype", _dtype(dtype) or torch.float32)
        object.__setattr__(self, "built", False)
        object.__setattr__(self, "_own_weights", [])

    @property
    def name(self):
        return self._name

    @property
    def dtype(self):
        return self._layer_dtype

    def add_weight(self, name=None, shape=None, dtype=None, initializer=None, trainable=True, **kwargs):
        init = _get_initializer(initializer)
        v = Variable.make(init(shape, _dtype(dtype) or torch.float32), name="/".join(_SCOPE + [name]) + ":0", trainable=trainable)
        self._own_weights.append(v)
        return v

    def build(self, input_shape):
        self.built = True

    def get_config(self):
        return {"name": self._name}

    def __call__(self, *args, **kwargs):
        args, kwargs = _wrap(args), _wrap(kwargs)                    # Keras converts array-like inputs to tf.Tensor
        with name_scope(self._name):
            if not self.built:
                first = args[0] if args else None
                self.build(first.shape if isinstance(first, torch.Tensor) else None)
                self.built = True
            return self.call(*args, **kwargs)

    # --- variable tracking: own weights first, then attributes in creation order (layers, lists of layers)
    def _children(self):
        for v in self.__dict__.values():
            if isinstance(v, Layer):
                yield v
            elif isinstance(v, (list, tuple)):
                for e in v:
                    if isinstance(e, Layer):
                        yield e

    @property
    def variables(self):
        out, seen = [], set()

        def walk(layer):
            if id(layer) in seen:
                return
            seen.add(id(layer))
            out.extend(layer._own_weights)
            for c in layer._children():
                walk(c)
        walk(self)
        return out

    weights = variables

    @property
    def trainable_variables(self):
        return [v for v in self.variables if v.trainable]

    trainable_weights = trainable_variables


class Model(Layer):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        object.__setattr__(self, "optimizer", None)
        object.__setattr__(self, "_train_counter", Variable.make(0, name="train_counter", trainable=False, dtype=torch.int64))

    @property
    def metrics(self):
        return []

    def compile(self, optimizer=None, **kwargs):
        self.optimizer = optimizer


class Dropout(Layer):
    def __init__(self, rate, **kwargs):
        super().__init__(**kwargs)
        self.rate = rate

    def call(self, x, training=False):
        if training and self.rate > 0:
            raise NotImplementedError("tf_shim: stochastic dropout (TensorFlow's RNG stream cannot be reproduced) — run with dropout = 0")
        return x


class LayerNormalization(Layer):
    def __init__(self, axis=-1, epsilon=1e-3, **kwargs):
        super().__init__(**kwargs)
        assert axis == -1
        self.epsilon = epsilon

    def build(self, input_shape):
        self.gamma = self.add_weight("gamma", shape=[input_shape[-1]], initializer=_Ones())
        self.beta = self.add_weight("beta", shape=[input_shape[-1]], initializer=_Zeros())

    def call(self, x):
        return F.layer_norm(x, (x.shape[-1],), self.gamma, self.beta, self.epsilon)


class Activation(Layer):
    def __init__(self, activation, **kwargs):
        super().__init__(**kwargs)
        self.activation = activation

    def call(self, x):
        return self.activation(x)


class Mean:
    """tf.keras.metrics.Mean"""

    def __init__(self, name="mean", dtype=None, **kwargs):
        self.name = name
        self.dtype = _dtype(dtype) or torch.float32
        self.reset_states()

    def reset_states(self):
        self.total, self.count = 0.0, 0.0

    def update_state(self, values, sample_weight=None):
        v = torch.as_tensor(values).detach().to(torch.float64).reshape(-1)
        if sample_weight is None:
            self.total += float(v.sum())
            self.count += v.numel()
        else:
            w = torch.as_tensor(sample_weight).detach().to(torch.float64).reshape(-1)
            self.total += float((v * w).sum())
            self.count += float(w.sum())

    def result(self):
        return _t(self.total / self.count if self.count else 0.0, torch.float32)


class Metric(Mean):
    pass


class _MeanMetricWrapper(Mean):
    """tf.keras.metrics.MeanMetricWrapper: both arguments are CAST to the metric's dtype (float32) — uint8 images are not rescaled —
    the wrapped function reduces the last axis, and the running mean is taken over what is left."""
    _fn = None

    def update_state(self, y_true, y_pred, sample_weight=None):
        a, b = _t(y_true).to(torch.float32), _t(y_pred).to(torch.float32)
        super().update_state(type(self)._fn(a, b), sample_weight)


class MeanSquaredError(_MeanMetricWrapper):
    _fn = staticmethod(lambda a, b: ((b - a) ** 2).mean(-1))


class MeanAbsoluteError(_MeanMetricWrapper):
    _fn = staticmethod(lambda a, b: (b - a).abs().mean(-1))


class GradientTape:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def gradient(self, target, sources):
        g = torch.autograd.grad(target, list(sources), allow_unused=True)
        return [None if x is None else _t(x) for x in g]


class LearningRateSchedule:
    pass


class CosineDecay(LearningRateSchedule):
    """tf.keras.experimental.CosineDecay"""

    def __init__(self, initial_learning_rate, decay_steps, alpha=0.0, name=None):
        self.initial_learning_rate, self.decay_steps, self.alpha = initial_learning_rate, decay_steps, alpha

    def __call__(self, step):
        step = torch.minimum(torch.as_tensor(step).to(torch.float32), torch.tensor(float(self.decay_steps)))
        cosine = 0.5 * (1 + torch.cos(torch.pi * step / self.decay_steps))
        return _t(self.initial_learning_rate * ((1 - self.alpha) * cosine + self.alpha))


class Adam:
    """tf.keras.optimizers.Adam (OptimizerV2): the hooks the reference's AdamWeightDecay overrides (_prepare_local, _resource_apply_dense,
    apply_gradients) with Keras' update  var -= lr_t * sqrt(1 - b2^t) / (1 - b1^t) * m / (sqrt(v) + eps)."""
    _use_locking = False

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, amsgrad=False, name="Adam", **kwargs):
        assert not amsgrad
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = learning_rate, beta_1, beta_2, epsilon
        self.iterations = Variable.make(0, name="iter", trainable=False, dtype=torch.int64)
        self._slots = {}

    def _decayed_lr(self, var_dtype):
        lr = self.learning_rate
        if callable(lr):
            lr = lr(self.iterations)
        return _t(lr, var_dtype)

    def _prepare_local(self, var_device, var_dtype, apply_state):
        lr_t = self._decayed_lr(var_dtype)
        local_step = float(self.iterations) + 1
        b1p, b2p = self.beta_1 ** local_step, self.beta_2 ** local_step
        apply_state[(var_device, var_dtype)] = dict(lr_t=lr_t, lr=lr_t * (np.sqrt(1 - b2p) / (1 - b1p)), epsilon=self.epsilon,
                                                    beta_1_t=self.beta_1, beta_2_t=self.beta_2)

    def _fallback_apply_state(self, var_device, var_dtype):
        st = {}
        self._prepare_local(var_device, var_dtype, st)
        return st[(var_device, var_dtype)]

    def _resource_apply_dense(self, grad, var, apply_state=None):
        c = (apply_state or {}).get((var.device, _dtype(var.dtype))) or self._fallback_apply_state(var.device, _dtype(var.dtype))
        m, v = self._slots.setdefault(id(var), (torch.zeros_like(var.detach()), torch.zeros_like(var.detach())))
        with torch.no_grad():
            g = grad.detach()
            m.mul_(c["beta_1_t"]).add_(g, alpha=1 - c["beta_1_t"])
            v.mul_(c["beta_2_t"]).addcmul_(g, g, value=1 - c["beta_2_t"])
            var.sub_(c["lr"] * m / (v.sqrt() + c["epsilon"]))

    def apply_gradients(self, grads_and_vars, name=None, **kwargs):
        gv = [(g, v) for g, v in grads_and_vars if g is not None]
        apply_state = {}
        for _, v in gv:
            if (v.device, _dtype(v.dtype)) not in apply_state:
                self._prepare_local(v.device, _dtype(v.dtype), apply_state)
        for g, v in gv:
            self._resource_apply_dense(g, v, apply_state=apply_state)
        self.iterations.assign_add(1)

    def get_config(self):
        return {}


class _Policy:
    compute_dtype = "float32"


# ------------------------------------------------------------------------------------------------------------ ops
def _reduce(fn):
    def red(x, axis=None, keepdims=False, name=None):
        x = _t(x)
        a = _axes(axis)
        if a is None:
            return fn(x)
        return fn(x, dim=a, keepdim=keepdims)
    return red


def _split(x, num_or_size_splits, axis=0, name=None):
    if isinstance(num_or_size_splits, (int, np.integer)):                # an int is the NUMBER of equal parts
        n = x.shape[axis]
        assert n % num_or_size_splits == 0
        return list(torch.split(x, n // int(num_or_size_splits), dim=axis))
    return list(torch.split(x, _ints(num_or_size_splits), dim=axis))


def _matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return torch.matmul(a, b)


def _constant(value, dtype=None, shape=None, name=None):
    out = _t(value, dtype).detach().clone()
    return out.reshape(_ints(shape)) if shape is not None else out


def _cast(x, dtype, name=None):
    return _t(x).to(_dtype(dtype))


def _range(start, limit=None, delta=1, dtype=None, name=None):
    if limit is None:
        start, limit = 0, start
    return _t(torch.arange(int(start), int(limit), int(delta), dtype=_dtype(dtype) or torch.int32))


def _shape(x, out_type=None, name=None):
    return _t(list(torch.Tensor.size(_t(x))), _dtype(out_type) or torch.int32)


def _one_hot(indices, depth, on_value=None, off_value=None, axis=None, dtype=None, name=None):
    assert on_value is None and off_value is None and axis in (None, -1)
    return F.one_hot(_t(indices).long(), int(depth)).to(_dtype(dtype) or torch.float32)


def _softmax_xent(labels=None, logits=None, axis=-1, name=None):
    return -(labels * F.log_softmax(logits, dim=axis)).sum(axis)


def _sparse_softmax_xent(labels=None, logits=None, name=None):
    lp = F.log_softmax(logits, dim=-1)
    return -torch.gather(lp, -1, _t(labels).long().unsqueeze(-1)).squeeze(-1)


def _clip_by_norm(t, clip_norm, axes=None, name=None):
    assert axes is None
    n = torch.sqrt((t * t).sum())
    return t * clip_norm / torch.maximum(n, torch.as_tensor(float(clip_norm)))


def _l2_normalize(x, axis=None, epsilon=1e-12, name=None):
    sq = (x * x).sum(dim=_axes(axis), keepdim=True)
    return x * torch.rsqrt(torch.clamp(sq, min=epsilon))


def _norm(tensor, ord="euclidean", axis=None, keepdims=None, name=None):
    p = 2 if ord in ("euclidean", 2) else ord
    return torch.linalg.vector_norm(tensor, ord=p, dim=_axes(axis), keepdim=bool(keepdims))


def _mse(y_true, y_pred):
    return ((_t(y_pred) - _t(y_true)) ** 2).mean(-1)


def _random_uniform(shape, minval=0, maxval=None, dtype=torch.float32, seed=None, name=None):
    dtype = _dtype(dtype)
    if dtype.is_floating_point:
        maxval = 1.0 if maxval is None else maxval
        return _t(torch.rand(_ints(shape), dtype=dtype) * (maxval - minval) + minval)
    return _t(torch.randint(int(minval), int(maxval), _ints(shape), dtype=dtype))


def _assert_type(tensor, tf_type, message=None, name=None):
    if _t(tensor).dtype != _dtype(tf_type):
        raise TypeError(f"tensor has dtype {tensor.dtype}, expected {tf_type}")


def _assert_equal(x, y, message=None, name=None):
    if not bool((_t(x) == _t(y)).all()):
        raise AssertionError(f"assert_equal failed: {x} vs {y}")


def _convert_image_dtype(image, dtype, saturate=False, name=None):
    """tf.image.convert_image_dtype: uint8 -> float multiplies by 1/255; float -> uint8 multiplies by 255.5 and casts (truncation)."""
    image, dtype = _t(image), _dtype(dtype)
    src = _dtype(image.dtype)
    if src == dtype:
        return image
    if not src.is_floating_point and dtype.is_floating_point:
        return image.to(dtype) * torch.tensor(1.0 / torch.iinfo(src).max, dtype=dtype)
    if src.is_floating_point and not dtype.is_floating_point:
        scaled = image * (torch.iinfo(dtype).max + 0.5)
        if saturate:
            scaled = scaled.clamp(torch.iinfo(dtype).min, torch.iinfo(dtype).max)
        return scaled.to(dtype)
    raise NotImplementedError((src, dtype))


def _assert_near(x, y, rtol=None, atol=None, message=None, summarize=None, name=None):
    x, y = _t(x), _t(y)
    eps = torch.finfo(_dtype(x.dtype)).eps
    rtol = 10 * eps if rtol is None else float(rtol)
    atol = 10 * eps if atol is None else float(atol)
    if not bool(((x - y).abs() <= atol + rtol * y.abs()).all()):
        raise AssertionError(f"assert_near failed: max |x - y| = {float((x - y).abs().max())}")


def _cond(pred, true_fn, false_fn, name=None):
    return true_fn() if bool(pred) else false_fn()


def _argmax(x, axis=None, output_type=torch.int64, name=None):
    return torch.argmax(x, dim=axis).to(_dtype(output_type))


def _binary(fn):
    def op(a, b, name=None):
        a = _t(a)
        return fn(a, _t(b, a.dtype) if not isinstance(b, torch.Tensor) else b)
    return op


def _unary(fn):
    return lambda x, name=None: fn(_t(x))


def _top_k(x, k=1, sorted=True, name=None):
    r = torch.topk(x, int(k), dim=-1, sorted=sorted)
    return r.values, r.indices.to(torch.int32)


def _depthwise_conv2d(input, filter, strides, padding, data_format=None, dilations=None, name=None):
    """tf.nn.depthwise_conv2d: input NHWC, filter [fh, fw, in_channels, channel_multiplier]; output channel k*mult + q is
    in-channel k filtered with filter[:, :, k, q] (cross-correlation, like every TF conv).  VALID padding only (utils/metrics.py:37)."""
    x, w = _t(input), _t(filter)
    if padding != "VALID" or (data_format not in (None, "NHWC")) or dilations not in (None, [1, 1], (1, 1)):
        raise NotImplementedError("shim depthwise_conv2d: VALID / NHWC only")
    fh, fw, cin, mult = w.shape
    wt = w.permute(2, 3, 0, 1).reshape(cin * mult, 1, fh, fw)          # torch grouped conv: out channel g*mult + q <- group g
    y = F.conv2d(x.permute(0, 3, 1, 2), wt.to(_dtype(x.dtype)), stride=(int(strides[1]), int(strides[2])), groups=cin)
    return _t(y.permute(0, 2, 3, 1))


def _psnr(a, b, max_val, name=None):
    """tf.image.psnr: 20*log10(max_val) - 10*log10(mean((a-b)^2 over the last three axes)), computed in float32."""
    a, b = _t(a).to(torch.float32), _t(b).to(torch.float32)
    mse = ((a - b) ** 2).mean(dim=(-3, -2, -1))
    return _t(20.0 * torch.log(torch.as_tensor(float(max_val))) / np.log(10.0) - 10.0 / np.log(10.0) * torch.log(mse))


def _mod(name):
    import importlib.machinery
    m = types.ModuleType(name)
    m.__path__ = []
    m.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)      # importlib.util.find_spec("tensorflow") must not raise
    return m


def build():
    tf = _mod("tensorflow")
    tf.__version__ = "2.4.1-viewformer-b200-shim"
    tf.Tensor, tf.Variable, tf.TensorShape, tf.GradientTape, tf.DType = Tensor, Variable.make, TensorShape, GradientTape, torch.dtype
    for k, v in _DT.items():
        setattr(tf, k, v)
    tf.newaxis = None
    tf.name_scope = name_scope
    tf.is_tensor = lambda x: isinstance(x, torch.Tensor)
    tf.convert_to_tensor = lambda value, dtype=None, name=None: _t(value, dtype)
    tf.constant, tf.cast, tf.range, tf.shape = _constant, _cast, _range, _shape
    tf.rank = lambda x, name=None: _t(x).dim()
    tf.reshape = lambda x, shape, name=None: _t(x).reshape(_ints(shape))
    tf.transpose = lambda x, perm=None, name=None: x.permute(*_ints(perm)) if perm is not None else x.permute(*reversed(range(x.dim())))
    tf.expand_dims = lambda x, axis, name=None: _t(x).unsqueeze(int(axis))
    tf.squeeze = lambda x, axis=None, name=None: x.squeeze() if axis is None else x.squeeze(int(axis))
    tf.concat = lambda values, axis, name=None: _t(torch.cat([_t(v) for v in values], dim=int(axis)))
    tf.stack = lambda values, axis=0, name=None: _t(torch.stack([_t(v) for v in values], dim=int(axis)))
    tf.unstack = lambda x, num=None, axis=0, name=None: list(torch.unbind(x, dim=int(axis)))
    tf.split = _split
    tf.repeat = lambda x, repeats, axis=None, name=None: torch.repeat_interleave(_t(x), int(repeats) if not isinstance(repeats, (list, tuple)) else torch.as_tensor(repeats), dim=axis)
    tf.tile = lambda x, multiples, name=None: x.repeat(*_ints(multiples))
    tf.broadcast_to = lambda x, shape, name=None: _t(x).expand(_ints(shape))
    tf.gather = lambda params, indices, axis=0, batch_dims=0, name=None: _t(params)[_t(indices).long()] if (axis == 0 and batch_dims == 0) else _gather_nd_axis(params, indices, axis, batch_dims)
    tf.where = lambda c, x=None, y=None, name=None: torch.where(c, _t(x), _t(y, _t(x).dtype) if not isinstance(y, torch.Tensor) else y)
    tf.matmul, tf.einsum = _matmul, lambda eq, *ops, **kw: torch.einsum(eq, *ops)
    tf.ones = lambda shape, dtype=torch.float32, name=None: _t(torch.ones(_ints(shape), dtype=_dtype(dtype)))
    tf.zeros = lambda shape, dtype=torch.float32, name=None: _t(torch.zeros(_ints(shape), dtype=_dtype(dtype)))
    tf.fill = lambda dims, value, name=None: _t(torch.full(_ints(dims), value.item() if isinstance(value, torch.Tensor) else value,
                                                           dtype=_dtype(value.dtype) if isinstance(value, torch.Tensor) else None))
    tf.ones_like = lambda x, dtype=None, name=None: torch.ones_like(_t(x), dtype=_dtype(dtype))
    tf.zeros_like = lambda x, dtype=None, name=None: torch.zeros_like(_t(x), dtype=_dtype(dtype))
    tf.one_hot, tf.argmax = _one_hot, _argmax
    for nm, f in (("exp", torch.exp), ("cos", torch.cos), ("sin", torch.sin), ("sqrt", torch.sqrt), ("abs", torch.abs), ("sign", torch.sign),
                  ("asin", torch.asin)):
        setattr(tf, nm, _unary(f))
    tf.minimum, tf.maximum = _binary(torch.minimum), _binary(torch.maximum)
    tf.reduce_mean, tf.reduce_sum = _reduce(torch.mean), _reduce(torch.sum)
    tf.reduce_max = lambda x, axis=None, keepdims=False, name=None: x.max() if axis is None else torch.amax(x, dim=_axes(axis), keepdim=keepdims)
    tf.reduce_min = lambda x, axis=None, keepdims=False, name=None: x.min() if axis is None else torch.amin(x, dim=_axes(axis), keepdim=keepdims)
    tf.reduce_all = lambda x, axis=None, name=None: x.all() if axis is None else x.all(dim=_axes(axis))
    tf.norm = _norm
    tf.clip_by_value = lambda t, lo, hi, name=None: torch.clamp(t, lo, hi)
    tf.clip_by_norm = _clip_by_norm
    tf.cond = _cond
    tf.no_op = lambda name=None: None
    tf.control_dependencies = lambda deps: contextlib.nullcontext()
    tf.sort = lambda x, axis=-1, direction="ASCENDING", name=None: torch.sort(x, dim=axis, descending=direction != "ASCENDING").values
    tf.identity = lambda x, name=None: x
    tf.stop_gradient = lambda x, name=None: x.detach()
    tf.function = lambda fn=None, **kw: fn if fn is not None else (lambda f: f)
    tf.zeros_initializer, tf.ones_initializer, tf.constant_initializer = _Zeros, _Ones, _Constant

    tf.math = _mod("tensorflow.math")
    tf.math.pow = lambda x, y, name=None: torch.pow(_t(x), y)
    tf.math.is_nan = lambda x, name=None: torch.isnan(x)
    tf.math.sqrt, tf.math.squared_difference = tf.sqrt, lambda a, b, name=None: (a - b) ** 2
    tf.math.atan2, tf.math.asin = (lambda y, x, name=None: torch.atan2(y, x)), tf.asin
    tf.math.top_k = _top_k
    tf.math.reduce_mean, tf.math.reduce_sum, tf.math.minimum, tf.math.maximum = tf.reduce_mean, tf.reduce_sum, tf.minimum, tf.maximum
    tf.math.log, tf.math.exp = (lambda x, name=None: torch.log(_t(x))), tf.exp

    tf.linalg = _mod("tensorflow.linalg")
    tf.linalg.norm, tf.linalg.l2_normalize, tf.linalg.matmul = _norm, _l2_normalize, _matmul

    tf.nn = _mod("tensorflow.nn")
    tf.nn.softmax = lambda logits, axis=-1, name=None: torch.softmax(logits, dim=axis)
    tf.nn.log_softmax = lambda logits, axis=-1, name=None: torch.log_softmax(logits, dim=axis)
    tf.nn.gelu = lambda x, approximate=False, name=None: F.gelu(x, approximate="tanh" if approximate else "none")
    tf.nn.softmax_cross_entropy_with_logits = _softmax_xent
    tf.nn.sparse_softmax_cross_entropy_with_logits = _sparse_softmax_xent
    tf.nn.compute_average_loss = lambda per_example_loss, sample_weight=None, global_batch_size=None: per_example_loss.sum() / (global_batch_size or per_example_loss.shape[0])
    tf.nn.l2_normalize = _l2_normalize
    tf.nn.top_k = _top_k
    tf.nn.depthwise_conv2d = _depthwise_conv2d

    tf.random = _mod("tensorflow.random")
    tf.random.uniform = _random_uniform
    tf.random.normal = lambda shape, mean=0.0, stddev=1.0, dtype=torch.float32, seed=None, name=None: _t(torch.randn(_ints(shape), dtype=_dtype(dtype)) * stddev + mean)
    tf.random.set_seed = lambda s: torch.manual_seed(s)

    tf.debugging = _mod("tensorflow.debugging")
    tf.debugging.assert_type, tf.debugging.assert_equal = _assert_type, _assert_equal
    tf.debugging.assert_near = _assert_near

    tf.losses = _mod("tensorflow.losses")
    tf.losses.mse = _mse

    tf.metrics = _mod("tensorflow.metrics")
    tf.metrics.Mean, tf.metrics.Metric = Mean, Metric

    tf.image = _mod("tensorflow.image")
    tf.image.convert_image_dtype = _convert_image_dtype
    tf.image.psnr = _psnr

    tf.io = _mod("tensorflow.io")
    tf.io.gfile = _mod("tensorflow.io.gfile")
    import os
    tf.io.gfile.exists = os.path.exists

    keras = _mod("tensorflow.keras")
    keras.Model = Model
    keras.layers = _mod("tensorflow.keras.layers")
    keras.layers.Layer, keras.layers.Dropout, keras.layers.LayerNormalization, keras.layers.Activation = Layer, Dropout, LayerNormalization, Activation
    keras.initializers = _mod("tensorflow.keras.initializers")
    keras.initializers.TruncatedNormal = TruncatedNormal
    keras.metrics = _mod("tensorflow.keras.metrics")
    keras.metrics.Mean, keras.metrics.Metric = Mean, Metric
    keras.metrics.MeanSquaredError, keras.metrics.MeanAbsoluteError = MeanSquaredError, MeanAbsoluteError
    keras.optimizers = _mod("tensorflow.keras.optimizers")
    keras.optimizers.Adam = Adam
    keras.optimizers.schedules = _mod("tensorflow.keras.optimizers.schedules")
    keras.optimizers.schedules.LearningRateSchedule = LearningRateSchedule
    keras.experimental = _mod("tensorflow.keras.experimental")
    keras.experimental.CosineDecay = CosineDecay
    keras.mixed_precision = _mod("tensorflow.keras.mixed_precision")
    keras.mixed_precision.global_policy = lambda: _Policy()
    keras.losses = tf.losses
    tf.keras = keras
    tf.initializers = keras.initializers
    tf.optimizers = keras.optimizers

    py = _mod("tensorflow.python")
    py.util = _mod("tensorflow.python.util")
    py.util.nest = _mod("tensorflow.python.util.nest")
    py.util.nest.map_structure = lambda fn, s: {k: fn(v) for k, v in s.items()} if isinstance(s, dict) else (type(s)(fn(v) for v in s) if isinstance(s, (list, tuple)) else fn(s))
    tf.python = py

    mods = {"tensorflow": tf, "tensorflow.math": tf.math, "tensorflow.linalg": tf.linalg, "tensorflow.nn": tf.nn, "tensorflow.random": tf.random,
            "tensorflow.debugging": tf.debugging, "tensorflow.losses": tf.losses, "tensorflow.metrics": tf.metrics, "tensorflow.io": tf.io, "tensorflow.image": tf.image,
            "tensorflow.keras": keras, "tensorflow.keras.layers": keras.layers, "tensorflow.keras.initializers": keras.initializers,
            "tensorflow.keras.metrics": keras.metrics, "tensorflow.keras.optimizers": keras.optimizers,
            "tensorflow.keras.optimizers.schedules": keras.optimizers.schedules, "tensorflow.keras.experimental": keras.experimental,
            "tensorflow.keras.mixed_precision": keras.mixed_precision, "tensorflow.python": py, "tensorflow.python.util": py.util,
            "tensorflow.python.util.nest": py.util.nest}
    return mods


def _gather_nd_axis(params, indices, axis, batch_dims):
    """tf.gather with batch_dims = axis = 1 style arguments (evaluate scripts): indices [B, n] pick along axis 1 of params [B, T, ...]"""
    assert axis == batch_dims or (axis in (1,) and batch_dims in (0, 1))
    params, idx = _t(params), _t(indices).long()
    if batch_dims == 0:
        return torch.index_select(params, axis, idx.reshape(-1)).reshape(params.shape[:axis] + tuple(idx.shape) + params.shape[axis + 1:])
    b = torch.arange(params.shape[0]).reshape(-1, *([1] * (idx.dim() - 1))).expand_as(idx)
    return params[b, idx]


_BUILT = None


def install():
    """Put the shim into sys.modules as ``tensorflow`` (refuses to shadow a real TensorFlow).  Idempotent: the same module objects are
    re-inserted after ``uninstall()``, so reference modules loaded earlier keep working."""
    global _BUILT
    cur = sys.modules.get("tensorflow")
    if cur is not None and not str(getattr(cur, "__version__", "")).endswith("shim"):
        raise RuntimeError("a real tensorflow is already imported")
    if _BUILT is None:
        _BUILT = build()
    sys.modules.update(_BUILT)
    return sys.modules["tensorflow"]


def uninstall():
    """Remove the shim's entries from sys.modules (the reference modules already loaded keep their reference to it)."""
    for name in list(_BUILT or {}):
        if sys.modules.get(name) is _BUILT[name]:
            del sys.modules[name]
