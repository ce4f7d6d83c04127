"""NumPy/torch-CPU stand-in for the ~25 TensorFlow names the reference touches
(SURVEY.md Appendix A.7).  TEST INFRASTRUCTURE ONLY: it exists so that the
reference's own agents/, training/, environments/ sources can be executed
verbatim from /root/reference to (a) validate oracle/rpbcac_oracle.py and
(b) generate the golden vectors in tests/golden/ (oracle/make_golden.py).

Keras-internal arithmetic follows SURVEY.md Appendix A ([TF-semantics]); the
gradients come from torch autograd (CPU, float32), i.e. a code path that is
independent of the hand-written backward pass in rpbcac_oracle.py."""
import sys
import types
import numpy as np

float32 = np.float32


class Tensor(np.ndarray):
    """ndarray with .numpy(); scalars produced by indexing stay Tensors so that
    `critic(x)[0][0].numpy()` (training/train_agents.py:62) works."""

    def numpy(self):
        a = np.asarray(self)
        return a[()] if a.ndim == 0 else a

    def __getitem__(self, idx):
        out = super().__getitem__(idx)
        if not isinstance(out, np.ndarray):
            out = np.asarray(out).view(Tensor)
        return out


def _t(a):
    return np.asarray(a).view(Tensor)


def convert_to_tensor(value, dtype=None):
    a = np.asarray(value if not isinstance(value, (list, tuple)) else [np.asarray(v) for v in value])
    if dtype is not None:
        a = a.astype(dtype)
    elif a.dtype == np.float64 and isinstance(value, (list, tuple)) and len(value) and \
            np.asarray(value[0]).dtype == np.float32:
        a = a.astype(np.float32)
    return _t(a)


def concat(values, axis):
    return _t(np.concatenate([np.asarray(v) for v in values], axis=axis))


def zeros(shape, dtype=np.float32):
    return _t(np.zeros(shape, dtype))


def sort(values, axis=-1):
    return _t(np.sort(np.asarray(values), axis=axis))


def clip_by_value(t, lo, hi):
    return _t(np.maximum(np.minimum(np.asarray(t), np.asarray(hi)), np.asarray(lo)))


def reduce_mean(t, axis=None):
    a = np.asarray(t)
    return _t(a.mean(axis=axis, dtype=a.dtype))


math = types.ModuleType("tensorflow.math")
math.minimum = lambda a, b: _t(np.minimum(np.asarray(a), np.asarray(b)))
math.maximum = lambda a, b: _t(np.maximum(np.asarray(a), np.asarray(b)))
math.square = lambda a: _t(np.square(np.asarray(a)))
math.reduce_sum = lambda a, axis=None: _t(np.asarray(a).sum(axis=axis, dtype=np.asarray(a).dtype))
sys.modules["tensorflow.math"] = math


class _Random(types.ModuleType):
    """tf.random.set_seed seeds the facade's *own* stream (weight init, fit
    shuffling); it never touches NumPy's global RNG, like real TF."""
    seed = 0
    rs = np.random.RandomState(0)

    def set_seed(self, seed):
        _Random.seed = seed
        _Random.rs = np.random.RandomState(seed)


random = _Random("tensorflow.random")
sys.modules["tensorflow.random"] = random

# Hook: callable(B) -> permutation used by Model.fit(shuffle=True).  Tests and
# make_golden.py install their own so the oracle can consume identical perms.
perm_hook = None


def _fit_permutation(B):
    if perm_hook is not None:
        return np.asarray(perm_hook(B))
    return _Random.rs.permutation(B)


class _Logger:
    def setLevel(self, *_a, **_k):
        pass


def get_logger():
    return _Logger()


from . import keras  # noqa: E402,F401
