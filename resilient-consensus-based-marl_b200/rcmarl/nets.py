"""Packed network parameters <-> Keras weight lists (main.py:60-82 layout:
[W1(d_in,20), b1, W2(20,20), b2, W3(20,n_out), b3], y = x @ W + b)."""
import numpy as np

from ._lib import HIDDEN, param_count


def shapes(d_in, n_out):
    return [(d_in, HIDDEN), (HIDDEN,), (HIDDEN, HIDDEN), (HIDDEN,), (HIDDEN, n_out), (n_out,)]


def pack(weights):
    """list of 6 arrays -> flat float32 vector (C order)."""
    return np.concatenate([np.asarray(a, np.float32).reshape(-1) for a in weights])


def unpack(flat, d_in, n_out):
    flat = np.asarray(flat, np.float32).reshape(-1)
    assert flat.size == param_count(d_in, n_out), (flat.size, d_in, n_out)
    out, o = [], 0
    for shp in shapes(d_in, n_out):
        n = int(np.prod(shp))
        out.append(flat[o:o + n].reshape(shp).copy())
        o += n
    return out


def n_hidden_params(d_in):
    """W1, b1, W2, b2 -- the arrays the hidden-layer consensus overwrites
    (agents/resilient_CAC_agents.py:153 `weights_agg[:-2]`)."""
    return d_in * HIDDEN + HIDDEN + HIDDEN * HIDDEN + HIDDEN


def glorot_uniform(d_in, n_out, rng):
    """Keras Dense default init: glorot_uniform kernels, zero biases."""
    ws = []
    for shp in shapes(d_in, n_out):
        if len(shp) == 2:
            lim = np.sqrt(6.0 / (shp[0] + shp[1]))
            ws.append(rng.uniform(-lim, lim, size=shp).astype(np.float32))
        else:
            ws.append(np.zeros(shp, np.float32))
    return ws
