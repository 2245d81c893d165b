"""Flat parameter layout of the actor / critic residual MLPs.

The CUDA kernels (csrc/navppo_kernels.cu) see each network as ONE contiguous fp32 vector;
the Python containers (nets.py) expose views of it under the reference's state_dict keys
(net_actor.py:56-93, net_critic.py:50-81) so `torch.save(actor.state_dict())` files are
interchangeable (main.py:52-89).  Order = the order `nn.Module.parameters()` yields the
LIVE parameters in the reference, i.e. what its two Adams iterate over (ppo.py:116-117)
minus the BatchNorm affine terms, which are constructed but never used in forward
(net_actor.py:44,48 commented out) and therefore never receive a gradient.

    offset   count   key                 shape
    0        8192    rb1.fc1.weight      [512, 16]
    8192     512     rb1.fc1.bias        [512]
    8704     8192    rb1.fc2.weight      [16, 512]
    16896    16      rb1.fc2.bias        [16]
    16912    16384   rb2.fc1.weight      [512, 32]
    33296    512     rb2.fc1.bias        [512]
    33808    16384   rb2.fc2.weight      [32, 512]
    50192    32      rb2.fc2.bias        [32]
    50224    ...     heads: actor  out1.weight[1,32] out1.bias[1] out2.weight[1,32] out2.bias[1]  (66)
                            critic out.weight[1,32]  out.bias[1]                                  (33)
"""
from __future__ import annotations

import numpy as np

OBS_DIM = 16
ACT_DIM = 2
HIDDEN = 512
X1_DIM = 2 * OBS_DIM

_TRUNK = [
    ("rb1.fc1.weight", (HIDDEN, OBS_DIM)),
    ("rb1.fc1.bias", (HIDDEN,)),
    ("rb1.fc2.weight", (OBS_DIM, HIDDEN)),
    ("rb1.fc2.bias", (OBS_DIM,)),
    ("rb2.fc1.weight", (HIDDEN, X1_DIM)),
    ("rb2.fc1.bias", (HIDDEN,)),
    ("rb2.fc2.weight", (X1_DIM, HIDDEN)),
    ("rb2.fc2.bias", (X1_DIM,)),
]
ACTOR_SPEC = _TRUNK + [("out1.weight", (1, X1_DIM)), ("out1.bias", (1,)), ("out2.weight", (1, X1_DIM)),
                       ("out2.bias", (1,))]
CRITIC_SPEC = _TRUNK + [("out.weight", (1, X1_DIM)), ("out.bias", (1,))]

# state_dict entries the reference carries but never uses (BatchNorm1d buffers + affine):
# kept so checkpoints round-trip key for key.  (name, num_features)
DEAD_BATCHNORMS = [("bn1", OBS_DIM), ("rb1.bn1", HIDDEN), ("rb1.bn2", OBS_DIM), ("rb2.bn1", HIDDEN),
                   ("rb2.bn2", X1_DIM)]


def _count(spec):
    return int(sum(int(np.prod(s)) for _, s in spec))


ACTOR_PARAMS = _count(ACTOR_SPEC)      # 50290
CRITIC_PARAMS = _count(CRITIC_SPEC)    # 50257
TOTAL_PARAMS = ACTOR_PARAMS + CRITIC_PARAMS
HEAD_OFFSET = _count(_TRUNK)           # 50224


def spec(kind: str):
    if kind == "actor":
        return ACTOR_SPEC
    if kind == "critic":
        return CRITIC_SPEC
    raise ValueError(kind)


def offsets(kind: str):
    """{key: (offset, shape)} inside the network's flat vector."""
    out, off = {}, 0
    for name, shape in spec(kind):
        out[name] = (off, shape)
        off += int(np.prod(shape))
    return out


def flatten_state_dict(sd, kind: str) -> np.ndarray:
    """Reference state_dict (torch tensors or arrays) -> flat float32 vector."""
    parts = []
    for name, shape in spec(kind):
        a = sd[name]
        a = a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
        assert tuple(a.shape) == tuple(shape), (name, a.shape, shape)
        parts.append(a.astype(np.float32).ravel())
    return np.concatenate(parts)


def unflatten(flat, kind: str):
    """Flat vector -> {key: array view of the right shape}."""
    out = {}
    for name, (off, shape) in offsets(kind).items():
        n = int(np.prod(shape))
        out[name] = flat[off:off + n].reshape(shape)
    return out
