"""Tensor-level ResNet18 feature extraction / classification on the CUDA path.

Replaces the model forward of the reference's ``extract_features`` hot loop
(``src/main.py:867-871``: ``feats = model(imgs.to(device))``) and the classifier forward of
``evaluate_resnet_classifier`` (``src/main.py:1004-1009``).  Weights come from a torchvision
``resnet18`` module / state dict (the reference's wrappers in ``src/models/resnet.py`` are thin
shells around exactly that); eval-mode BatchNorm is folded in ``hipac_resnet18_pack``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

_BLOCKS = [("layer1", 0, False), ("layer1", 1, False), ("layer2", 0, True), ("layer2", 1, False),
           ("layer3", 0, True), ("layer3", 1, False), ("layer4", 0, True), ("layer4", 1, False)]


def conv_bn_names():
    """(conv, bn) module-name pairs in the C ABI's layer order (torchvision state-dict order)."""
    names = [("conv1", "bn1")]
    for layer, blk, ds in _BLOCKS:
        p = f"{layer}.{blk}"
        names += [(f"{p}.conv1", f"{p}.bn1"), (f"{p}.conv2", f"{p}.bn2")]
        if ds:
            names.append((f"{p}.downsample.0", f"{p}.downsample.1"))
    return names


class PackedResNet18:
    """Folded + re-laid-out weights resident on one device."""

    def __init__(self, blob: torch.Tensor, num_classes: int):
        self.blob = blob
        self.num_classes = num_classes

    @property
    def device(self):
        return self.blob.device


def pack_resnet18(state_dict: dict, device, prefix: str = "", bn_eps: float = 1e-5) -> PackedResNet18:
    """Pack a torchvision-style resnet18 state dict (keys ``{prefix}conv1.weight`` ...).

    A present ``{prefix}fc.weight`` of shape ``[k,512]`` becomes the classifier head."""
    l = _lib.lib()
    tensors = []
    for conv, bn in conv_bn_names():
        for key in (f"{conv}.weight", f"{bn}.weight", f"{bn}.bias", f"{bn}.running_mean", f"{bn}.running_var"):
            tensors.append(state_dict[prefix + key].detach().to("cpu", torch.float32).contiguous())
    fcw = state_dict.get(prefix + "fc.weight")
    num_classes = 0
    if fcw is not None and fcw.dim() == 2 and fcw.shape[1] == 512:
        num_classes = int(fcw.shape[0])
        tensors.append(fcw.detach().to("cpu", torch.float32).contiguous())
        tensors.append(state_dict[prefix + "fc.bias"].detach().to("cpu", torch.float32).contiguous())
    ptrs = (C.c_void_p * 102)(*([t.data_ptr() for t in tensors] + [None] * (102 - len(tensors))))
    nbytes = l.hipac_resnet18_packed_bytes(num_classes)
    host = np.zeros(nbytes, dtype=np.uint8)
    _lib.check(l.hipac_resnet18_pack(ptrs, 102, num_classes, bn_eps, host.ctypes.data_as(C.c_void_p), nbytes),
               "hipac_resnet18_pack")
    return PackedResNet18(torch.from_numpy(host).to(device), num_classes)


_LAYOUT_BY_SHAPE = {(224, 224, 3): _lib.LAYOUT_NHWC3_BF16, (112, 115, 16): _lib.LAYOUT_S2D16_BF16}


def _forward(batch: torch.Tensor, packed: PackedResNet18, want_logits: bool, chunk: int, stream=None,
             count: torch.Tensor | None = None):
    """``count``: int32 device tensor whose element 0 is the number of valid patches in ``batch`` (device-count mode:
    ``batch.shape[0]`` is only the capacity; nothing here waits for the value)."""
    l = _lib.lib()
    if not (batch.is_cuda and batch.dtype == torch.bfloat16 and batch.dim() == 4 and tuple(batch.shape[1:]) in _LAYOUT_BY_SHAPE):
        raise ValueError("batch must be a CUDA bf16 tensor [N,224,224,3] or [N,112,115,16]")
    if batch.device != packed.device:
        raise ValueError("batch and packed weights live on different devices")
    if want_logits and packed.num_classes == 0:
        raise ValueError("packed weights have no classifier head")
    batch = batch.contiguous()
    n = int(batch.shape[0])
    dev = batch.device
    st = stream or torch.cuda.current_stream(dev)
    with torch.cuda.device(dev), torch.cuda.stream(st):
        feats = torch.empty((n, 512), dtype=torch.float32, device=dev)
        logits = torch.empty((n, packed.num_classes), dtype=torch.float32, device=dev) if want_logits else None
        if n == 0:
            return feats, logits
        ws_bytes = l.hipac_resnet18_workspace_bytes(n, chunk)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        if count is None:
            rc = l.hipac_resnet18_forward(packed.blob.data_ptr(), packed.num_classes, batch.data_ptr(),
                                          _LAYOUT_BY_SHAPE[tuple(batch.shape[1:])], n, feats.data_ptr(),
                                          logits.data_ptr() if logits is not None else None, ws.data_ptr(), ws_bytes,
                                          chunk, st.cuda_stream)
        else:
            if not (count.is_cuda and count.dtype == torch.int32 and count.numel() >= 1):
                raise ValueError("count must be a CUDA int32 tensor")
            rc = l.hipac_resnet18_forward_dcount(packed.blob.data_ptr(), packed.num_classes, batch.data_ptr(),
                                                 _LAYOUT_BY_SHAPE[tuple(batch.shape[1:])], n, count.data_ptr(),
                                                 feats.data_ptr(), logits.data_ptr() if logits is not None else None,
                                                 ws.data_ptr(), ws_bytes, chunk, st.cuda_stream)
        _lib.check(rc, "hipac_resnet18_forward")
        ws.record_stream(st)
    return feats, logits


def extract_features_tensor(batch: torch.Tensor, packed: PackedResNet18, chunk: int = 4096, stream=None,
                            count: torch.Tensor | None = None) -> torch.Tensor:
    """float32 ``[N,512]`` pooled trunk features of a bf16 patch batch (asynchronous)."""
    return _forward(batch, packed, False, chunk, stream, count)[0]


def classify_tensor(batch: torch.Tensor, packed: PackedResNet18, chunk: int = 4096, stream=None,
                    count: torch.Tensor | None = None):
    """(features float32 ``[N,512]``, logits float32 ``[N,k]``) of a bf16 patch batch (asynchronous)."""
    return _forward(batch, packed, True, chunk, stream, count)


def conv_layer(packed: PackedResNet18, layer: int, x: torch.Tensor, residual: torch.Tensor | None, relu: bool):
    """Test hook: run one conv(+folded BN)(+residual)(+ReLU) layer on bf16 NHWC input."""
    l = _lib.lib()
    cout_hout = {0: (64, 112), **{i: (64, 56) for i in range(1, 5)}, **{i: (128, 28) for i in range(5, 10)},
                 **{i: (256, 14) for i in range(10, 15)}, **{i: (512, 7) for i in range(15, 20)}}
    cout, hout = cout_hout[layer]
    n = int(x.shape[0])
    out = torch.empty((n, hout, hout, cout), dtype=torch.bfloat16, device=x.device)
    st = torch.cuda.current_stream(x.device)
    rc = l.hipac_resnet18_conv_layer(packed.blob.data_ptr(), packed.num_classes, layer, x.contiguous().data_ptr(),
                                     residual.contiguous().data_ptr() if residual is not None else None,
                                     out.data_ptr(), n, int(relu), st.cuda_stream)
    _lib.check(rc, "hipac_resnet18_conv_layer")
    return out


def stem(packed: PackedResNet18, x: torch.Tensor) -> torch.Tensor:
    """Test hook: fused conv1 + BN + ReLU + max pool on an S2D16 batch -> bf16 ``[n,56,56,64]``."""
    l = _lib.lib()
    n = int(x.shape[0])
    out = torch.empty((n, 56, 56, 64), dtype=torch.bfloat16, device=x.device)
    rc = l.hipac_resnet18_stem(packed.blob.data_ptr(), packed.num_classes, x.contiguous().data_ptr(), out.data_ptr(), n,
                               torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(rc, "hipac_resnet18_stem")
    return out


def conv_ds_fused(packed: PackedResNet18, stage: int, x: torch.Tensor, block_in: torch.Tensor) -> torch.Tensor:
    """Test hook: conv2 of layer{2,3,4}.0 + fused 1x1/s2 projection shortcut + ReLU (stage 0/1/2)."""
    l = _lib.lib()
    n, cout, hout = int(x.shape[0]), (128, 256, 512)[stage], (28, 14, 7)[stage]
    out = torch.empty((n, hout, hout, cout), dtype=torch.bfloat16, device=x.device)
    rc = l.hipac_resnet18_conv_ds_fused(packed.blob.data_ptr(), packed.num_classes, stage, x.contiguous().data_ptr(),
                                        block_in.contiguous().data_ptr(), out.data_ptr(), n,
                                        torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(rc, "hipac_resnet18_conv_ds_fused")
    return out
