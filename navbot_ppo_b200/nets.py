"""NetActor / NetCritic: weight containers with the reference's constructor, state_dict keys
and initialisation (project_ppo/src/net_actor.py:56-144, net_critic.py:50-130), whose
forward runs the sm_100a inference kernel (csrc/navppo_kernels.cu) through the C-ABI.

The modules are built from the same torch layers in the same order as the reference, so under
the same `torch.manual_seed` they start from bit-identical weights, and
`torch.save(net.state_dict())` files are interchangeable with the reference's checkpoints
(main.py:52-89), dead BatchNorm entries included.  The live parameters are views into one
flat fp32 vector (layout.py) — that vector is what the kernels read and Adam updates.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np
import torch
from torch import nn

from . import _capi, layout


class _Handles:
    """One navppo handle per (device, max_samples bucket), shared by nets and trainers."""

    _cache: dict = {}

    @classmethod
    def get(cls, device: torch.device, max_samples: int = 1 << 20, clip: float = 0.2, lr: float = 3e-4,
            precision: int = _capi.PREC_FP32):
        key = (device.index, int(max_samples), float(clip), float(lr), int(precision))
        if key not in cls._cache:
            cfg = _capi.default_ppo_cfg()
            cfg.device = device.index
            cfg.max_samples = int(max_samples)
            cfg.clip = float(clip)
            cfg.lr = float(lr)
            cfg.precision = int(precision)
            h = ctypes.c_void_p()
            _capi.check(_capi.lib().navppo_create(ctypes.byref(h), ctypes.byref(cfg)))
            cls._cache[key] = h
        return cls._cache[key]


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class ResBlock(nn.Module):
    """Parameter container for the reference's ResBlock (net_actor.py:16-53).  bn1 / bn2 are
    constructed because the reference constructs them (their keys are in every checkpoint) and
    never applied, exactly like the reference (:44,48 commented out)."""

    def __init__(self, Fin, Fout, n_neurons=512, actor_init=False):
        super().__init__()
        self.Fin, self.Fout = Fin, Fout
        self.fc1 = nn.Linear(Fin, n_neurons)
        if actor_init:  # net_actor.py:28 (the critic keeps torch's default init, net_critic.py:26)
            nn.init.uniform_(self.fc1.weight, -1 / math.sqrt(Fin), 1 / math.sqrt(Fin))
        self.bn1 = nn.BatchNorm1d(n_neurons)
        self.fc2 = nn.Linear(n_neurons, Fout)
        if actor_init:  # net_actor.py:32
            nn.init.uniform_(self.fc2.weight, -1 / math.sqrt(n_neurons), 1 / math.sqrt(n_neurons))
        self.bn2 = nn.BatchNorm1d(Fout)


class _FlatNet(nn.Module):
    kind = "actor"

    def _bind(self, flat: torch.Tensor):
        """Move the live parameters into `flat` (this network's slice of the trainer's flat
        vector) and re-point them at views of it."""
        named = dict(self.named_parameters())
        with torch.no_grad():
            for name, (off, shape) in layout.offsets(self.kind).items():
                n = int(np.prod(shape))
                view = flat[off:off + n].view(shape)
                view.copy_(named[name].data.to(flat.device))
                named[name].data = view
        for name, p in self.named_parameters():   # dead BatchNorm affines follow the device
            if p.device != flat.device:
                p.data = p.data.to(flat.device)
        for b in self.buffers():
            b.data = b.data.to(flat.device)

    def _ensure_bound(self, device=None):
        if getattr(self, "_owner_flat", None) is None:
            dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
            buf = torch.zeros(_capi.PPO_FLAT, dtype=torch.float32, device=dev)
            off = 0 if self.kind == "actor" else _capi.PPO_CRITIC_OFFSET
            n = layout.ACTOR_PARAMS if self.kind == "actor" else layout.CRITIC_PARAMS
            self._bind(buf[off:off + n])
            self._owner_flat = buf
        return self._owner_flat

    def attach(self, flat_all: torch.Tensor):
        """Called by the trainer: live in the trainer's flat [actor | critic] vector."""
        off = 0 if self.kind == "actor" else _capi.PPO_CRITIC_OFFSET
        n = layout.ACTOR_PARAMS if self.kind == "actor" else layout.CRITIC_PARAMS
        self._bind(flat_all[off:off + n])
        self._owner_flat = flat_all

    def to(self, *args, **kwargs):  # PPO does `.to(device)` (ppo.py:112-113); placement is the kernels' device
        return self

    def _prep(self, obs):
        flat = self._ensure_bound()
        if isinstance(obs, np.ndarray):                      # net_actor.py:105-106
            obs = torch.tensor(obs, dtype=torch.float)
        obs = obs.to(device=flat.device, dtype=torch.float32)
        if obs.dim() == 1:                                   # net_actor.py:110-112
            obs = obs.unsqueeze(0)
        if obs.shape[-1] != layout.OBS_DIM:
            raise ValueError(f"observation must be {layout.OBS_DIM}-D, got {tuple(obs.shape)}")
        return flat, obs.contiguous()


class NetActor(_FlatNet):
    kind = "actor"

    def __init__(self, in_dim, out_dim, n_neurons=512, use_vision=False, vision_feat_dim=1280, vision_proj_dim=64,
                 **kwargs):
        super().__init__()
        if use_vision:
            raise NotImplementedError("the camera path is outside the LiDAR hot path (SURVEY.md section 2, row 9)")
        if in_dim != layout.OBS_DIM or out_dim != layout.ACT_DIM or n_neurons != layout.HIDDEN:
            raise ValueError("the kernels are built for the reference's 16 -> (512) -> 2 network")
        self.use_vision, self.base_state_dim, self.vision_proj = False, in_dim, None
        d = in_dim
        self.bn1 = nn.BatchNorm1d(d)                                    # net_actor.py:85 (never applied, :136)
        self.rb1 = ResBlock(d, d, n_neurons, actor_init=True)
        self.rb2 = ResBlock(2 * d, 2 * d, n_neurons, actor_init=True)
        self.out1 = nn.Linear(2 * d, out_dim - 1)
        nn.init.uniform_(self.out1.weight, -1 / math.sqrt(d), 1 / math.sqrt(d))              # :89
        self.out2 = nn.Linear(2 * d, out_dim - 1)
        nn.init.uniform_(self.out2.weight, -1 / math.sqrt(2 * d), 1 / math.sqrt(2 * d))      # :91
        self.do = nn.Dropout(p=.1, inplace=False)                       # :92, unused

    def forward(self, obs, vision_feat=None):
        flat, obs = self._prep(obs)
        mu = torch.empty((obs.shape[0], 2), dtype=torch.float32, device=flat.device)
        h = _Handles.get(flat.device)
        _capi.check(_capi.lib().navppo_forward(h, flat.data_ptr(), obs.data_ptr(), obs.shape[0], mu.data_ptr(), None,
                                               _stream(flat.device)))
        return mu


class NetCritic(_FlatNet):
    kind = "critic"

    def __init__(self, in_dim, out_dim, n_neurons=512, use_vision=False, vision_feat_dim=1280, vision_proj_dim=64,
                 **kwargs):
        super().__init__()
        if use_vision:
            raise NotImplementedError("the camera path is outside the LiDAR hot path (SURVEY.md section 2, row 9)")
        if in_dim != layout.OBS_DIM or out_dim != 1 or n_neurons != layout.HIDDEN:
            raise ValueError("the kernels are built for the reference's 16 -> (512) -> 1 network")
        self.use_vision, self.base_state_dim, self.vision_proj = False, in_dim, None
        d = in_dim
        self.bn1 = nn.BatchNorm1d(d)
        self.rb1 = ResBlock(d, d, n_neurons)
        self.rb2 = ResBlock(2 * d, 2 * d, n_neurons)
        self.out = nn.Linear(2 * d, out_dim)                            # net_critic.py:80
        self.do = nn.Dropout(p=.1, inplace=False)

    def forward(self, obs, vision_feat=None):
        flat, obs = self._prep(obs)
        v = torch.empty((obs.shape[0], 1), dtype=torch.float32, device=flat.device)
        h = _Handles.get(flat.device)
        _capi.check(_capi.lib().navppo_forward(h, flat.data_ptr(), obs.data_ptr(), obs.shape[0], None, v.data_ptr(),
                                               _stream(flat.device)))
        return v
