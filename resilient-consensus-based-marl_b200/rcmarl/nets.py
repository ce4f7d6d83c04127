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


# ---- agent counts other than the two kernel instantiations (main.py:26 takes any --n_agents) ------------------------------
# The kernels are instantiated for 5 and 16 agents.  A team of n agents runs on the next instantiation with the extra
# agent slots ZERO: zero input columns and zero rows of W1.  Forward values are unchanged (0 * w adds exactly 0), the
# gradient of a padded W1 row is sum(x_pad * delta) = 0, so SGD / Adam / consensus keep those rows at zero: the padded
# problem is the n-agent problem bit for bit, not an approximation.
KERNEL_AGENTS = (5, 16)


def kernel_agents(n_agents):
    """Smallest kernel instantiation that holds `n_agents` agents."""
    for k in KERNEL_AGENTS:
        if n_agents <= k:
            return k
    raise ValueError(f"n_agents = {n_agents} exceeds the largest kernel instantiation ({KERNEL_AGENTS[-1]})")


def pack_padded(weights, d_in_k):
    """list of 6 arrays with W1 of shape (d_in, 20), d_in <= d_in_k -> packed vector of the d_in_k-input network."""
    w = [np.asarray(a, np.float32) for a in weights]
    d_in = w[0].shape[0]
    if d_in == d_in_k:
        return pack(w)
    assert d_in < d_in_k
    W1 = np.zeros((d_in_k, HIDDEN), np.float32)
    W1[:d_in] = w[0]
    return pack([W1] + w[1:])


def unpack_padded(flat, d_in, d_in_k, n_out):
    """Inverse of pack_padded: the first d_in rows of W1 and the other five arrays."""
    w = unpack(flat, d_in_k, n_out)
    w[0] = w[0][:d_in].copy()
    return w


def pad_agent_slots(x, n_agents, n_kernel):
    """(B, n_agents, f) or (B, n_agents * f) torch tensor -> (B, n_kernel * f) with zero slots for the extra agents."""
    import torch
    B = x.shape[0]
    x = x.reshape(B, n_agents, -1)
    if n_agents == n_kernel:
        return x.reshape(B, -1).contiguous()
    out = torch.zeros(B, n_kernel, x.shape[2], dtype=x.dtype, device=x.device)
    out[:, :n_agents] = x
    return out.reshape(B, -1)


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
