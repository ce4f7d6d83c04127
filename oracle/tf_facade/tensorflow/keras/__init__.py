"""Keras names used by the reference on torch-CPU
autograd.  TEST INFRASTRUCTURE ONLY -- see tensorflow/__init__.py."""
import sys
import types
import numpy as np
import torch

import tensorflow as _tf

torch.set_grad_enabled(True)


# ------------------------------------------------------------------ layers
class _Sym:
    """Symbolic handle (`layer.output`, `model.inputs`)."""

    def __init__(self, model, idx):
        self.model, self.idx = model, idx


class InputSpec:
    def __init__(self, shape):
        self.shape = tuple(shape)


def Input(shape=None, **_k):
    return InputSpec(shape)


class LeakyReLU:
    def __init__(self, alpha=0.3):
        self.alpha = alpha

    def __call__(self, z):
        return torch.where(z > 0, z, self.alpha * z)


class Flatten:
    trainable = True
    n_weights = 0

    def forward(self, x, logits=False):
        return x.reshape(x.shape[0], -1)

    def get_weights(self):
        return []

    def set_weights(self, w):
        assert len(w) == 0


class Dense:
    n_weights = 2

    def __init__(self, units, activation=None):
        self.units, self.activation = units, activation
        self.trainable = True
        self.W = self.b = None
        self._model = None
        self._idx = None

    def build(self, d_in):
        lim = np.sqrt(6.0 / (d_in + self.units))                 # glorot_uniform
        W = _tf._Random.rs.uniform(-lim, lim, size=(d_in, self.units)).astype(np.float32)
        self.W = torch.tensor(W, requires_grad=True)
        self.b = torch.zeros(self.units, dtype=torch.float32, requires_grad=True)

    def forward(self, x, logits=False):
        z = x @ self.W + self.b
        if self.activation is None:
            return z
        if self.activation == 'softmax':
            return z if logits else torch.softmax(z, dim=-1)
        return self.activation(z)

    def get_weights(self):
        return [self.W.detach().numpy().copy(), self.b.detach().numpy().copy()]

    def set_weights(self, w):
        assert len(w) == 2 and tuple(w[0].shape) == tuple(self.W.shape) and tuple(w[1].shape) == tuple(self.b.shape)
        with torch.no_grad():
            self.W.copy_(torch.as_tensor(np.asarray(w[0], np.float32)))
            self.b.copy_(torch.as_tensor(np.asarray(w[1], np.float32)))

    @property
    def output(self):
        return _Sym(self._model, self._idx)


layers = types.ModuleType("tensorflow.keras.layers")
layers.Flatten, layers.Dense, layers.LeakyReLU, layers.Input = Flatten, Dense, LeakyReLU, Input
sys.modules["tensorflow.keras.layers"] = layers


# ------------------------------------------------------------ optimizers
class SGD:
    def __init__(self, learning_rate=0.01):
        self.lr = learning_rate

    def apply(self, params, grads):
        with torch.no_grad():
            for p, g in zip(params, grads):
                p -= self.lr * g


class Adam:
    """TF-2.x Keras Adam: eps outside the bias correction."""

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta_1, beta_2, epsilon
        self.t, self.m, self.v = 0, None, None

    def apply(self, params, grads):
        if self.m is None:
            self.m = [torch.zeros_like(p) for p in params]
            self.v = [torch.zeros_like(p) for p in params]
        self.t += 1
        lr_t = np.float32(self.lr * np.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t))
        with torch.no_grad():
            for p, g, m, v in zip(params, grads, self.m, self.v):
                m.mul_(self.b1).add_(g, alpha=1.0 - self.b1)
                v.mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
                p -= float(lr_t) * m / (v.sqrt() + self.eps)


optimizers = types.ModuleType("tensorflow.keras.optimizers")
optimizers.SGD, optimizers.Adam = SGD, Adam
sys.modules["tensorflow.keras.optimizers"] = optimizers


# ----------------------------------------------------------------- losses
class MeanSquaredError:
    needs_logits = False

    def per_sample(self, y_true, y_pred):
        return ((y_pred - y_true.reshape(y_pred.shape)) ** 2).mean(dim=-1)


class SparseCategoricalCrossentropy:
    needs_logits = True          # traced train step back-tracks softmax to its logits

    def per_sample(self, y_true, logits):
        lsm = torch.log_softmax(logits, dim=-1)
        idx = y_true.reshape(-1).long()
        return -lsm[torch.arange(lsm.shape[0]), idx]


losses = types.ModuleType("tensorflow.keras.losses")
losses.MeanSquaredError, losses.SparseCategoricalCrossentropy = MeanSquaredError, SparseCategoricalCrossentropy
sys.modules["tensorflow.keras.losses"] = losses


# ------------------------------------------------------------------ models
class History:
    def __init__(self):
        self.history = {'loss': []}


class Model:
    def __init__(self, inputs=None, outputs=None):
        sym = outputs
        src = sym.model
        self._input_spec = src._input_spec
        self.layers = src.layers[:sym.idx + 1]            # shares layer objects (res..py:39-40)
        self._train_layers = None
        self.optimizer = self.loss = None

    # -- structure
    @property
    def inputs(self):
        return [_Sym(self, -1)]

    @property
    def output_shape(self):
        d = None
        for l in self.layers:
            if isinstance(l, Dense):
                d = l.units
        return (None, d)

    @property
    def trainable(self):
        return all(l.trainable for l in self.layers)

    @trainable.setter
    def trainable(self, val):
        for l in self.layers:
            l.trainable = bool(val)

    def get_weights(self):
        return [a for l in self.layers for a in l.get_weights()]

    def set_weights(self, w):
        w = list(w)
        assert len(w) == sum(l.n_weights for l in self.layers)
        i = 0
        for l in self.layers:
            l.set_weights(w[i:i + l.n_weights])
            i += l.n_weights

    # -- execution
    def _forward(self, x, logits=False):
        h = torch.as_tensor(np.asarray(x, dtype=np.float32))
        for l in self.layers:
            h = l.forward(h, logits=logits)
        return h

    def __call__(self, x):
        with torch.no_grad():
            return _tf._t(self._forward(x).numpy())

    def predict(self, x):
        with torch.no_grad():
            return self._forward(x).numpy()

    def compile(self, optimizer=None, loss=None):
        self.optimizer, self.loss = optimizer, loss
        # Keras collects trainable weights when the train function is built
        self._train_layers = [l for l in self.layers if isinstance(l, Dense) and l.trainable]

    def _train_step(self, x, y, sw):
        params = [p for l in self._train_layers for p in (l.W, l.b)]
        out = self._forward(x, logits=self.loss.needs_logits)
        yt = torch.as_tensor(np.asarray(y, dtype=np.float32))
        per = self.loss.per_sample(yt, out)
        if sw is not None:
            per = per * torch.as_tensor(np.asarray(sw, dtype=np.float32)).reshape(-1)
        loss = per.sum() / per.shape[0]                     # SUM_OVER_BATCH_SIZE (Appendix A.2)
        grads = torch.autograd.grad(loss, params)
        self.optimizer.apply(params, grads)
        return float(loss.detach())

    def train_on_batch(self, x, y, sample_weight=None):
        return np.float32(self._train_step(x, y, sample_weight))

    def fit(self, x, y, batch_size=None, epochs=1, verbose=0, sample_weight=None, shuffle=True):
        x = np.asarray(x, np.float32)
        y = np.asarray(y, np.float32)
        sw = None if sample_weight is None else np.asarray(sample_weight, np.float32)
        B = x.shape[0]
        bs = 32 if batch_size is None else int(batch_size)
        hist = History()
        for _ in range(epochs):
            # a single full batch is order-independent: no permutation is drawn (keeps the
            # injected-permutation stream aligned with oracle/rpbcac_oracle.py)
            perm = _tf._fit_permutation(B) if (shuffle and bs < B) else np.arange(B)
            tot, cnt = 0.0, 0
            for k in range(0, B, bs):
                idx = perm[k:k + bs]
                l = self._train_step(x[idx], y[idx], None if sw is None else sw[idx])
                tot += l * len(idx)
                cnt += len(idx)
            hist.history['loss'].append(np.float32(tot / cnt))
        return hist


class Sequential(Model):
    def __init__(self, layer_list):
        spec = layer_list[0]
        assert isinstance(spec, InputSpec)
        self._input_spec = spec
        self.layers = list(layer_list[1:])
        d = int(np.prod(spec.shape))
        for i, l in enumerate(self.layers):
            if isinstance(l, Dense):
                l.build(d)
                d = l.units
                l._model, l._idx = self, i
        self._train_layers = None
        self.optimizer = self.loss = None


sys.modules["tensorflow.keras"] = sys.modules[__name__]
