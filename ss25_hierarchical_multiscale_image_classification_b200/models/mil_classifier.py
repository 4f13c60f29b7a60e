"""ABMIL head over a bag of patch features: ``MILAttentionPooling`` / ``MILClassifier`` with the reference's names, constructor
arguments and state-dict keys (``src/models/mil_classifier.py:5-45``: ``aggregator.attn_V``, ``aggregator.attn_U``,
``classifier.0``, ``classifier.2``), so a reference checkpoint loads unchanged.  ``forward`` on a CUDA bag runs the fused
kernels of ``csrc/mil.cu`` (projection + tanh + score, online softmax over the instances, weighted sum, two Linear layers:
two launches, fp32); there is no PyTorch fallback for CUDA inputs.  CPU tensors take the plain ``nn.Module`` arithmetic of
the reference, which is what parameter initialisation, ``state_dict`` round trips and the CPU tests use.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from .. import _lib

_POOLING = {"attention": 0, "mean": 1, "max": 2}


def _f32(t):
    return t.detach().to("cpu", torch.float32).contiguous()


def pack_mil(attn_V: nn.Linear | None, attn_U: nn.Linear | None, fc1: nn.Linear, fc2: nn.Linear, device) -> torch.Tensor:
    """The head's parameters as one device blob in the layout ``hipac_mil_forward`` reads."""
    if fc1.in_features != 512 or fc1.out_features != 128 or fc2.in_features != 128:
        raise ValueError("the fused head is built for Linear(512,128) -> ReLU -> Linear(128,k) (the reference's layout)")
    if attn_V is not None and (attn_V.in_features != 512 or attn_V.out_features != 128 or attn_U.out_features != 1):
        raise ValueError("the fused head is built for attn_V = Linear(512,128), attn_U = Linear(128,1) (the reference's defaults)")
    l = _lib.lib()
    k = int(fc2.out_features)
    host = np.zeros(int(l.hipac_mil_packed_floats(k)), dtype=np.float32)
    keep = [_f32(fc1.weight), _f32(fc1.bias), _f32(fc2.weight), _f32(fc2.bias)]
    att = [_f32(attn_V.weight), _f32(attn_V.bias), _f32(attn_U.weight), _f32(attn_U.bias)] if attn_V is not None else [None] * 4
    ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None   # noqa: E731
    _lib.check(l.hipac_mil_pack(*[ptr(t) for t in att], *[ptr(t) for t in keep], k, host.ctypes.data_as(C.c_void_p)), "hipac_mil_pack")
    return torch.from_numpy(host).to(device)


def mil_forward(bag: torch.Tensor, packed: torch.Tensor, num_classes: int, pooling: str = "attention", count: torch.Tensor | None = None,
                want_attention: bool = True, stream=None):
    """(logits float32 ``[k]``, attention float32 ``[N, 1]`` | None, pooled float32 ``[512]``) of a CUDA bag ``[N, 512]``.
    ``count``: optional int32 device tensor (element 0 = number of valid instances, e.g. the tile scan's survivor counter)."""
    if not (bag.is_cuda and bag.dtype == torch.float32 and bag.dim() == 2 and bag.shape[1] == 512):
        raise ValueError("bag must be a CUDA float32 tensor of shape [N, 512]")
    bag = bag.contiguous()
    n, dev = int(bag.shape[0]), bag.device
    l = _lib.lib()
    st = stream or torch.cuda.current_stream(dev)
    with torch.cuda.device(dev), torch.cuda.stream(st):
        logits = torch.empty((num_classes,), dtype=torch.float32, device=dev)
        pooled = torch.empty((512,), dtype=torch.float32, device=dev)
        attn = torch.empty((max(n, 1),), dtype=torch.float32, device=dev) if (want_attention and pooling == "attention") else None
        ws = torch.empty((int(l.hipac_mil_workspace_bytes(n)),), dtype=torch.uint8, device=dev)
        _lib.check(l.hipac_mil_forward(bag.data_ptr() if n else None, n, count.data_ptr() if count is not None else None, packed.data_ptr(),
                                       num_classes, _POOLING[pooling], logits.data_ptr(), attn.data_ptr() if attn is not None else None,
                                       pooled.data_ptr(), ws.data_ptr(), int(ws.numel()), st.cuda_stream), "hipac_mil_forward")
        ws.record_stream(st)
    return logits, (attn[:n].unsqueeze(1) if attn is not None else None), pooled


class MILAttentionPooling(nn.Module):
    """Attention-based pooling as in Ilse et al. (ABMIL); reference ``src/models/mil_classifier.py:5-18``."""

    def __init__(self, in_dim, attn_dim=128):
        super().__init__()
        self.attn_V = nn.Linear(in_dim, attn_dim)
        self.attn_U = nn.Linear(attn_dim, 1)

    def forward(self, x):
        A = torch.tanh(self.attn_V(x))
        A = self.attn_U(A)
        A = torch.softmax(A, dim=0)
        M = torch.sum(A * x, dim=0)
        return M, A


class MILClassifier(nn.Module):
    """Reference ``src/models/mil_classifier.py:20-45``: pooling ('attention' | 'mean' | 'max') + ``Linear(feature_dim,128) -> ReLU
    -> Linear(128, num_classes)``; ``forward(bag) -> (logits [num_classes], attention weights [N,1] | None)``."""

    def __init__(self, feature_dim, num_classes=2, pooling='attention'):
        super().__init__()
        self.pooling = pooling
        if pooling == 'attention':
            self.aggregator = MILAttentionPooling(feature_dim)
        elif pooling == 'mean':
            self.aggregator = lambda x: (x.mean(dim=0), None)
        elif pooling == 'max':
            self.aggregator = lambda x: (x.max(dim=0)[0], None)
        else:
            raise ValueError("Unknown pooling: choose from 'attention', 'mean', 'max'")
        self.classifier = nn.Sequential(nn.Linear(feature_dim, 128), nn.ReLU(), nn.Linear(128, num_classes))
        self._packed = None

    def _packed_for(self, device):
        key = (str(device), tuple(int(p._version) for p in self.parameters()))
        if self._packed is None or self._packed[0] != key:
            agg = self.aggregator if self.pooling == 'attention' else None
            self._packed = (key, pack_mil(agg.attn_V if agg is not None else None, agg.attn_U if agg is not None else None,
                                          self.classifier[0], self.classifier[2], device))
        return self._packed[1]

    def forward(self, bag, count=None):
        """bag: ``[num_patches, feature_dim]``.  CUDA bags run the fused kernels; CPU bags the reference arithmetic."""
        if bag.is_cuda:
            logits, attn, _ = mil_forward(bag.float(), self._packed_for(bag.device), int(self.classifier[2].out_features), self.pooling, count)
            return logits, attn
        pooled, attn = self.aggregator(bag)
        return self.classifier(pooled), attn
