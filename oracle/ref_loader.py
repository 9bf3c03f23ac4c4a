"""Load pieces of the read-only reference (/root/reference) on the CPU -- TEST INFRASTRUCTURE ONLY.

Used by oracle/gen_golden.py (to make tests/golden/) and by container-only tests that are skipped
when /root/reference is absent (it never exists on the GPU box).  Nothing here is imported by the
product package.

The reference's model files import third-party packages that are not installed in this image
(mamba_ssm, monai, timm, dynamic_network_architectures, ...).  We satisfy those imports with empty
stub modules, give real behaviour only to the handful of symbols the SS2D / SSND modules actually
execute (DropPath, trunc_normal_, monai's conv-only Convolution), and bind
``mamba_ssm.ops.selective_scan_interface.selective_scan_fn`` to the reference's own pure-PyTorch
``selective_scan_ref`` so that the reference modules run their own arithmetic end to end
(recipe: SURVEY.md appendix).
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types
import warnings

REF_ROOT = os.environ.get("NNUZOO_REFERENCE_ROOT", "/root/reference")
_STUB_ROOTS = ("mamba_ssm", "monai", "timm", "dynamic_network_architectures", "batchgenerators",
               "batchgeneratorsv2", "acvl_utils", "torchinfo", "nnunetv2", "causal_conv1d",
               "causal_conv1d_cuda", "selective_scan_cuda", "prettytable", "deep_utils")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "nnunetv2", "nets"))


class _Anything:
    """Placeholder for any attribute of a stubbed package (usable in annotations like ``A | str``)."""

    def __init__(self, name="stub"):
        self._name = name

    def __call__(self, *a, **k):
        raise RuntimeError(f"stubbed third-party symbol {self._name} was actually called")

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return _Anything(f"{self._name}.{item}")

    def __or__(self, other):
        return self

    __ror__ = __or__

    def __mro_entries__(self, bases):  # allow ``class X(stub):``
        return (object,)


class _StubModule(types.ModuleType):
    __path__: list = []

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return _Anything(f"{self.__name__}.{item}")


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


_installed = False
_cache: dict = {}


def _install_stubs():
    global _installed
    if _installed:
        return
    import torch
    import torch.nn as nn

    sys.meta_path.insert(0, _StubFinder())
    _installed = True

    # --- timm.layers: the two symbols SS2D-family files execute ---------------------------------
    import timm.layers as tl  # noqa: E402  (stub)

    class DropPath(nn.Module):  # stochastic depth; identity in eval() and at drop_prob 0
        def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
            super().__init__()
            self.drop_prob, self.scale_by_keep = drop_prob, scale_by_keep

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1 - self.drop_prob
            mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
            if keep > 0.0 and self.scale_by_keep:
                mask.div_(keep)
            return x * mask

    tl.DropPath = DropPath
    tl.trunc_normal_ = lambda t, std=1.0, **kw: nn.init.trunc_normal_(t, std=std, **kw)

    # --- monai conv-only Convolution (ssnd2net.py:110-120): Sequential with one "conv" child ------
    import monai.networks.blocks as mb  # noqa: E402  (stub)

    class Convolution(nn.Sequential):
        def __init__(self, spatial_dims, in_channels, out_channels, strides=1, kernel_size=3,
                     groups=1, bias=True, conv_only=False, dilation=1, padding=None, **kw):
            super().__init__()
            assert conv_only, "only the conv_only form is stood in for"
            conv_t = {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d}[spatial_dims]
            if padding is None:  # monai: same_padding(kernel_size, dilation)
                ks = (kernel_size,) * spatial_dims if isinstance(kernel_size, int) else tuple(kernel_size)
                dl = (dilation,) * spatial_dims if isinstance(dilation, int) else tuple(dilation)
                padding = tuple((k - 1) // 2 * d for k, d in zip(ks, dl))
            self.add_module("conv", conv_t(in_channels, out_channels, kernel_size, stride=strides,
                                           padding=padding, dilation=dilation, groups=groups, bias=bias))

    mb.Convolution = Convolution
    import monai.networks.blocks.convolutions as mbc  # noqa: E402

    mbc.Convolution = Convolution

    # --- monai factories the LightM-UNet blocks execute (lm2net.py:17-18, :130-132) ----------------
    import monai.networks.layers.utils as mlu  # noqa: E402  (stub)

    def get_norm_layer(name, spatial_dims=1, channels=1):
        nm, args = (name, {}) if isinstance(name, str) else (name[0], dict(name[1]))
        nm = nm.lower()
        if nm == "group":
            return nn.GroupNorm(num_channels=channels, **args)
        if nm == "instance":
            return {2: nn.InstanceNorm2d, 3: nn.InstanceNorm3d}[spatial_dims](channels, **args)
        if nm == "batch":
            return {2: nn.BatchNorm2d, 3: nn.BatchNorm3d}[spatial_dims](channels, **args)
        raise NotImplementedError(name)

    def get_act_layer(name):
        nm, args = (name, {}) if isinstance(name, str) else (name[0], dict(name[1]))
        return {"relu": nn.ReLU, "leakyrelu": nn.LeakyReLU, "gelu": nn.GELU}[nm.lower()](**args)

    mlu.get_norm_layer, mlu.get_act_layer = get_norm_layer, get_act_layer

    import dynamic_network_architectures.initialization.weight_init as wi  # noqa: E402

    wi.init_last_bn_before_add_to_0 = lambda m: None

    # --- the scan itself: the reference's own pure-PyTorch statement -----------------------------
    ssi = load_file("ref_selective_scan_interface",
                    "nnunetv2/nets/seg_mamba/selective_scan_interface.py")
    import mamba_ssm.ops.selective_scan_interface as mssi  # noqa: E402  (stub)

    mssi.selective_scan_fn = ssi.selective_scan_ref
    mssi.selective_scan_ref = ssi.selective_scan_ref
    _ = torch


def load_file(modname: str, relpath: str):
    """Import one reference source file by path under a private module name."""
    if modname in _cache:
        return _cache[modname]
    if not available():
        raise FileNotFoundError(f"reference not mounted at {REF_ROOT}")
    if modname != "ref_selective_scan_interface":
        _install_stubs()
    else:
        for m in ("causal_conv1d", "causal_conv1d_cuda", "selective_scan_cuda"):
            if m not in sys.modules:
                sys.modules[m] = types.ModuleType(m)
                sys.modules[m].causal_conv1d_fn = None
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(mod)
    _cache[modname] = mod
    return mod


def selective_scan_ref():
    """The reference's oracle, selective_scan_interface.py:86-152, verbatim."""
    return load_file("ref_selective_scan_interface",
                     "nnunetv2/nets/seg_mamba/selective_scan_interface.py").selective_scan_ref


def mamba_simple():
    """The reference's vendored Mamba block (seg_mamba/mamba_simple.py) wired to run on the CPU:
    ``selective_scan_fn`` := the reference's own selective_scan_ref, ``causal_conv1d_fn`` := None (the
    file's own F.conv1d fallback, :316-317), and the CUDA-only fused ``mamba_inner_fn*`` replaced by a
    stand-in that follows MambaInnerFnNoOutProj.forward (selective_scan_interface.py:159-226) op for op
    with those two substitutions."""
    import torch.nn.functional as F
    from einops import rearrange

    ref = selective_scan_ref()
    # mamba_simple.py:13-16 needs both names importable (its except branch cannot run: `a, b = None`)
    cc = sys.modules.get("causal_conv1d") or types.ModuleType("causal_conv1d")
    cc.causal_conv1d_fn = None
    cc.causal_conv1d_update = None
    sys.modules["causal_conv1d"] = cc
    mod = load_file("ref_mamba_simple", "nnunetv2/nets/seg_mamba/mamba_simple.py")
    mod.selective_scan_fn = ref
    mod.causal_conv1d_fn = None

    def inner_no_out_proj(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, A, B=None, C=None,
                          D=None, delta_bias=None, B_proj_bias=None, C_proj_bias=None, delta_softplus=True):
        L = xz.shape[-1]
        delta_rank = delta_proj_weight.shape[1]
        d_state = A.shape[-1]
        x, z = xz.chunk(2, dim=1)
        w = conv1d_weight.shape[-1]
        x = F.silu(F.conv1d(x, conv1d_weight, conv1d_bias, padding=w - 1, groups=x.shape[1])[..., :L])
        x_dbl = F.linear(rearrange(x, "b d l -> (b l) d"), x_proj_weight)
        delta = rearrange(delta_proj_weight @ x_dbl[:, :delta_rank].t(), "d (b l) -> b d l", l=L)
        Bv = rearrange(x_dbl[:, delta_rank:delta_rank + d_state], "(b l) dstate -> b 1 dstate l", l=L).contiguous()
        Cv = rearrange(x_dbl[:, -d_state:], "(b l) dstate -> b 1 dstate l", l=L).contiguous()
        return ref(x, delta, A, Bv, Cv, D, z=z, delta_bias=delta_bias, delta_softplus=delta_softplus)

    def inner(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, out_proj_weight, out_proj_bias, A,
              B=None, C=None, D=None, delta_bias=None, B_proj_bias=None, C_proj_bias=None, delta_softplus=True):
        y = inner_no_out_proj(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, A, B, C, D,
                              delta_bias, B_proj_bias, C_proj_bias, delta_softplus)
        return F.linear(rearrange(y, "b d l -> b l d"), out_proj_weight, out_proj_bias)

    mod.mamba_inner_fn_no_out_proj = inner_no_out_proj
    mod.mamba_inner_fn = inner
    return mod


def m2net():
    return load_file("ref_m2net", "nnunetv2/nets/m2net.py")


def _bind_mamba():
    """``from mamba_ssm import Mamba`` (lm2net.py:14, mamba_nd2net.py:26) := the reference's vendored block wired for the
    CPU (mamba_simple() above); with bimamba_type "none" it is the same contract as upstream's."""
    ms = mamba_simple()
    import mamba_ssm  # noqa: E402  (stub)
    mamba_ssm.Mamba = ms.Mamba
    return ms


def lm2net():
    """LightM-UNet blocks (MambaLayer, GSC, ResMambaBlock): nnunetv2/nets/lm2net.py."""
    _bind_mamba()
    return load_file("ref_lm2net", "nnunetv2/nets/lm2net.py")


def mamba_nd2net():
    """MambaND blocks (Block, create_block, MambaNDCore): nnunetv2/nets/mamba_nd2net.py."""
    _bind_mamba()
    return load_file("ref_mamba_nd2net", "nnunetv2/nets/mamba_nd2net.py")


def ssnd2net():
    return load_file("ref_ssnd2net", "nnunetv2/nets/ssnd2net.py")
