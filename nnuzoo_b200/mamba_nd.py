"""The callers of the 1-D scan path: LightM-UNet's Mamba layers (Alt1DM2Net / LM2Net) and MambaND's ordered blocks.

Reference classes mirrored here (same constructor arguments, parameter names and shapes, so reference state_dicts load):

  MambaLayer        nnunetv2/nets/lm2net.py:64-92       LN -> Mamba -> + skip_scale * x -> LN -> Linear over flattened tokens
  GSC               lm2net.py:417-460                    gated spatial convolution in front of the Mamba layers
  ResMambaBlock     lm2net.py:107-176                    GSC, then twice (norm, act, nd_mamba_order(order, x, MambaLayer)), + identity
  nd_mamba_order    lm2net.py:161-176                    axis order of the token walk: "d h w" / "d w h" / "w h d" (2-D: "h w" / "w h")
  Block             nnunetv2/nets/mamba_nd2net.py:565-666  LN -> Mamba (+ skip) over tokens re-ordered per `order`, reversed layers
  MambaNDCore       mamba_nd2net.py:725-1001             patch embedding + the stack of Blocks: orders cycle every 2 layers,
                                                         odd layers run reversed

What is different from the reference is layout work only: the token re-orderings are single permuted copies (no einops
round trips), and a reversed Block does not flip anything -- LayerNorm and the residual add are per token, so
flip(LN(flip x) + Mamba(LN(flip x))) == LN(x) + flip(Mamba(flip(LN x))), and the inner flip pair is the reversed walk of
``Mamba.forward(..., reverse=True)`` (anti-causal convolution + reversed-walk scan; nnuzoo_b200/mamba_inner.py).
``mamba_ssm.Mamba`` -- what the reference imports for these nets (lm2net.py:14, mamba_nd2net.py:26) -- has one direction
only, so the Mamba blocks here are built with ``extra_directions=False`` (no ``*_b`` / ``*_s`` parameters).
"""
from __future__ import annotations

from functools import partial
from typing import Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from .mamba import Mamba


def _conv_only(spatial_dims, cin, cout, kernel_size=3, stride=1, padding=None, groups=1, bias=True):
    """monai ``Convolution(..., conv_only=True)``: a Sequential with one child named ``conv`` (state_dict naming) and
    'same' padding when none is given."""
    conv_t = {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d}[spatial_dims]
    if padding is None:
        ks = (kernel_size,) * spatial_dims if isinstance(kernel_size, int) else tuple(kernel_size)
        padding = tuple((k - 1) // 2 for k in ks)
    seq = nn.Sequential()
    seq.add_module("conv", conv_t(cin, cout, kernel_size, stride=stride, padding=padding, groups=groups, bias=bias))
    return seq


def get_dwconv_layer(spatial_dims, in_channels, out_channels, kernel_size=3, stride=1, bias=False, padding=None):
    """depthwise conv followed by a 1x1 conv (lm2net.py:44-61, mamba_nd2net.py:171-186)."""
    return nn.Sequential(
        _conv_only(spatial_dims, in_channels, in_channels, kernel_size, stride, padding, groups=in_channels, bias=bias),
        _conv_only(spatial_dims, in_channels, out_channels, 1, 1, padding if padding is not None else None, bias=bias))


class MambaLayer(nn.Module):
    """lm2net.py:64-92.  x: (B, C, *spatial) -> (B, output_dim, *spatial); tokens are the flattened spatial axes."""

    def __init__(self, input_dim, output_dim, d_state=16, d_conv=4, expand=2):
        super().__init__()
        self.input_dim, self.output_dim = input_dim, output_dim
        self.norm = nn.LayerNorm(input_dim)
        self.mamba = Mamba(d_model=input_dim, d_state=d_state, d_conv=d_conv, expand=expand, extra_directions=False)
        self.proj = nn.Linear(input_dim, output_dim)
        self.skip_scale = nn.Parameter(torch.ones(1))

    def forward(self, x):
        if x.dtype == torch.float16:  # :79-80
            x = x.type(torch.float32)
        B, C = x.shape[:2]
        assert C == self.input_dim
        img_dims = x.shape[2:]
        x_flat = x.reshape(B, C, -1).transpose(-1, -2)           # (B, L, C) view
        x_norm = self.norm(x_flat)
        x_mamba = self.mamba(x_norm) + self.skip_scale * x_flat
        x_mamba = self.proj(self.norm(x_mamba))
        return x_mamba.transpose(-1, -2).reshape(B, self.output_dim, *img_dims)


_AXES3 = {"d": 2, "h": 3, "w": 4}
_AXES2 = {"h": 2, "w": 3}


def nd_mamba_order(order: str, x: torch.Tensor, mamba_module: nn.Module, spatial_dims: Optional[int] = None):
    """Run ``mamba_module`` over the tokens of x walked in axis order ``order`` and return the result in x's own axis order
    (ResMambaBlock.nd_mamba_order, lm2net.py:161-176: rearrange, module, rearrange back)."""
    spatial_dims = x.dim() - 2 if spatial_dims is None else spatial_dims
    axes = _AXES3 if spatial_dims == 3 else _AXES2
    names = order.split()
    if sorted(names) != sorted(axes):
        raise ValueError(f"order {order!r} must be a permutation of {' '.join(axes)}")
    perm = [0, 1] + [axes[n] for n in names]
    if perm == list(range(x.dim())):
        return mamba_module(x)
    inv = [0] * x.dim()
    for i, p in enumerate(perm):
        inv[p] = i
    return mamba_module(x.permute(perm)).permute(inv)


class InstanceNorm(nn.Module):
    def __init__(self, spatial_dims, in_channels):
        super().__init__()
        self.layer = (nn.InstanceNorm2d if spatial_dims == 2 else nn.InstanceNorm3d)(in_channels)

    def forward(self, x):
        return self.layer(x)


class GSC(nn.Module):
    """lm2net.py:417-460."""

    def __init__(self, spatial_dims, in_channels):
        super().__init__()
        self.proj = get_dwconv_layer(spatial_dims, in_channels, in_channels, stride=1, bias=True)
        self.norm = InstanceNorm(spatial_dims, in_channels)
        self.nonliner = nn.ReLU()
        self.proj2 = _conv_only(spatial_dims, in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.norm2 = InstanceNorm(spatial_dims, in_channels)
        self.nonliner2 = nn.ReLU()
        self.proj3 = get_dwconv_layer(spatial_dims, in_channels, in_channels, stride=1, bias=True)
        self.norm3 = InstanceNorm(spatial_dims, in_channels)
        self.nonliner3 = nn.ReLU()

    def forward(self, x):
        x1 = self.nonliner(self.proj(self.norm(x)))
        x2 = self.nonliner2(self.proj2(self.norm2(x)))
        y = x1 + x2
        return self.nonliner3(self.proj3(self.norm3(y))) + x


def _norm_layer(norm, spatial_dims, channels):
    """The subset of monai ``get_norm_layer`` LM2Net uses (lm2net.py:234: ("GROUP", {"num_groups": 8}); also instance / batch)."""
    name, args = (norm, {}) if isinstance(norm, str) else (norm[0], dict(norm[1]))
    name = name.lower()
    if name == "group":
        return nn.GroupNorm(num_channels=channels, **args)
    if name == "instance":
        return (nn.InstanceNorm2d if spatial_dims == 2 else nn.InstanceNorm3d)(channels, **args)
    if name == "batch":
        return (nn.BatchNorm2d if spatial_dims == 2 else nn.BatchNorm3d)(channels, **args)
    raise NotImplementedError(f"norm {norm!r}")


def _act_layer(act):
    name, args = (act, {}) if isinstance(act, str) else (act[0], dict(act[1]))
    table = {"relu": nn.ReLU, "leakyrelu": nn.LeakyReLU, "gelu": nn.GELU, "silu": nn.SiLU, "swish": nn.SiLU}
    if name.lower() not in table:
        raise NotImplementedError(f"act {act!r}")
    return table[name.lower()](**args)


class ResMambaBlock(nn.Module):
    """lm2net.py:107-176."""

    def __init__(self, spatial_dims, in_channels, norm, kernel_size=3, act=("RELU", {"inplace": True}), order="d h w"):
        super().__init__()
        if kernel_size % 2 != 1:
            raise AssertionError("kernel_size should be an odd number.")
        self.order, self.spatial_dims = order, spatial_dims
        self.gsc = GSC(spatial_dims, in_channels)
        self.norm1 = _norm_layer(norm, spatial_dims, in_channels)
        self.norm2 = _norm_layer(norm, spatial_dims, in_channels)
        self.act = _act_layer(act)
        self.mamba1 = MambaLayer(input_dim=in_channels, output_dim=in_channels)
        self.mamba2 = MambaLayer(input_dim=in_channels, output_dim=in_channels)

    def nd_mamba_order(self, order, x, mamba_module):
        return nd_mamba_order(order, x, mamba_module, self.spatial_dims)

    def forward(self, x):
        x = self.gsc(x)
        identity = x
        x = self.act(self.norm1(x))
        x = self.nd_mamba_order(self.order, x, self.mamba1)
        x = self.act(self.norm2(x))
        x = self.nd_mamba_order(self.order, x, self.mamba2)
        return x + identity


# ------------------------------------------------------------------------------------------------------------------
# MambaND (mamba_nd2net.py)
# ------------------------------------------------------------------------------------------------------------------
class DropPath(nn.Module):
    """Stochastic depth per sample (mamba_nd2net.py:510-544)."""

    def __init__(self, drop_prob: float = 0.1):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = (keep + torch.rand((x.shape[0],) + (1,) * (x.ndim - 1), dtype=x.dtype, device=x.device)).floor()
        return x.div(keep) * mask


class Dropout(nn.Dropout):
    def __init__(self, drop_prob: float = 0.5, inplace: bool = False):
        super().__init__(p=drop_prob, inplace=inplace)


def _token_perm(order: str, n_dim_pos: int):
    """Axis permutation of the (n, t, h, w, c) view for Block.forward's ``order`` (mamba_nd2net.py:623-631)."""
    names = order.split()
    if n_dim_pos != 4:
        if len(names) != 4:
            raise AssertionError("with n_dim_pos != 4 the order must name 4 axes")
        raise NotImplementedError("n_dim_pos != 4 (batch-folded leading axes) is not used by nnUZoo's MambaND2Net")
    ax = {"t": 1, "h": 2, "w": 3}
    if sorted(names) != ["h", "t", "w"]:
        raise ValueError(f"order {order!r} must be a permutation of 't h w'")
    return [0] + [ax[n] for n in names] + [4]


class Block(nn.Module):
    """mamba_nd2net.py:565-666: LN -> mixer (+ skip onto the normed tokens), tokens walked in ``order``; a ``reverse``
    block walks them backwards.  hidden_states: (n, t*h*w, c) in "t h w" order, returned in the same order."""

    def __init__(self, spatial_dims, dim, mixer_cls, norm_cls=nn.LayerNorm, fused_add_norm=False, residual_in_fp32=False,
                 reverse=False, drop_path_rate=0.0, drop_rate=0.0):
        super().__init__()
        self.spatial_dims = spatial_dims
        self.residual_in_fp32, self.fused_add_norm = residual_in_fp32, fused_add_norm
        self.mixer = mixer_cls(dim)
        self.norm = norm_cls(dim)
        self.reverse = reverse
        self.drop_path = DropPath(drop_prob=drop_path_rate)
        self.dropout = Dropout(drop_prob=drop_rate)
        self.ffn = None
        if fused_add_norm and not isinstance(self.norm, nn.LayerNorm):
            raise NotImplementedError("fused_add_norm is implemented for LayerNorm only")

    def _mix(self, h, inference_params):
        if isinstance(self.mixer, Mamba):  # the reversed walk by addressing: no flipped copies
            return self.mixer(h, inference_params=inference_params, reverse=self.reverse)
        if self.reverse:
            return self.mixer(h.flip(1), inference_params=inference_params).flip(1)
        return self.mixer(h, inference_params=inference_params)

    def forward(self, hidden_states, residual=None, inference_params=None, order="t h w", shape=None, skip=True,
                n_dim_pos=4):
        assert shape is not None
        if self.spatial_dims == 3:
            t, h, w = shape
        else:
            (h, w), t = shape, 1
        n, L, c = hidden_states.shape
        perm = _token_perm(order, n_dim_pos)
        ident = perm == [0, 1, 2, 3, 4]

        def to_order(x):
            return x if ident else x.reshape(n, t, h, w, c).permute(perm).reshape(n, L, c)

        hidden_states = to_order(hidden_states)
        if not self.fused_add_norm:  # :642-649
            hidden_states = self.norm(hidden_states)
            mixed = self.drop_path(self.dropout(self._mix(hidden_states, inference_params)))
            hidden_states = hidden_states + mixed if skip else mixed
        else:  # :650-661 -- layer_norm_fn(..., residual, prenorm=True): LN(x + residual); the residual stream is dropped
            if residual is not None:
                hidden_states = hidden_states + to_order(residual)
            if self.residual_in_fp32:
                hidden_states = hidden_states.float()
            hidden_states = self.norm(hidden_states.to(self.norm.weight.dtype))
            hidden_states = self.drop_path(self._mix(hidden_states, inference_params))
        if ident:
            return hidden_states
        inv = [0] * 5
        for i, p in enumerate(perm):
            inv[p] = i
        dims = [(n, t, h, w, c)[p] for p in perm]
        return hidden_states.reshape(dims).permute(inv).reshape(n, L, c)


def create_block(spatial_dims, d_model, ssm_cfg=None, norm_epsilon=1e-5, rms_norm=False, residual_in_fp32=False,
                 fused_add_norm=False, layer_idx=None, device=None, dtype=None, reverse=None, drop_rate=0.1,
                 drop_path_rate=0.1):
    """mamba_nd2net.py:672-722."""
    if rms_norm:
        raise NotImplementedError("RMSNorm blocks are not used by nnUZoo's MambaND2Net (rms_norm=False, :852)")
    ssm_cfg = dict(ssm_cfg or {})
    factory_kwargs = {"device": device, "dtype": dtype}
    mixer_cls = partial(Mamba, layer_idx=layer_idx, extra_directions=False, **ssm_cfg, **factory_kwargs)
    norm_cls = partial(nn.LayerNorm, eps=norm_epsilon, **factory_kwargs)
    block = Block(spatial_dims, d_model, mixer_cls, norm_cls=norm_cls, fused_add_norm=fused_add_norm,
                  residual_in_fp32=residual_in_fp32, reverse=bool(reverse), drop_rate=drop_rate,
                  drop_path_rate=drop_path_rate)
    block.layer_idx = layer_idx
    return block


class PatchEmbed(nn.Module):
    """Strided depthwise + 1x1 convolution to tokens (mamba_nd2net.py:189-311, the zero-padding form MambaNDCore uses)."""

    def __init__(self, spatial_dims, in_channels, embed_dims, kernel_size, stride=None, bias=True):
        super().__init__()
        stride = kernel_size if stride is None else stride
        pad = (0,) * spatial_dims
        self.embed_dims = embed_dims
        self.projection = get_dwconv_layer(spatial_dims, in_channels, embed_dims, kernel_size=kernel_size, stride=stride,
                                           padding=pad, bias=bias)
        self.norm = None

    def forward(self, x):
        x = self.projection(x)
        out_size = (x.shape[2], x.shape[3])                      # (:306) two extents only, as in the reference
        return x.flatten(2).transpose(1, 2), out_size


class MambaNDCore(nn.Module):
    """mamba_nd2net.py:725-1001: patch embedding, then ``num_layers`` Blocks.  Layer i walks the tokens in order
    orders[(i // 2) % len(orders)] -- 3-D: "t h w", "t w h", "w h t"; 2-D: "t h w", "t w h" -- and odd layers walk it
    backwards (:848).  forward(x) -> (tokens of the last layer, list of every layer's tokens).  (final_norm=True works
    here; in the reference ``self.ln1`` is then the (name, layer) tuple build_norm_layer returns and forward raises, :872,
    :998 -- MambaND2Net builds the core with final_norm=False, :1136.)"""

    ORDERS = {3: ("t h w", "t w h", "w h t"), 2: ("t h w", "t w h")}

    def __init__(self, spatial_dims, img_size=224, patch_size=16, in_channels=3, drop_rate=0.0, drop_path_rate=0.0,
                 norm_cfg=None, final_norm=True, pre_norm=False, expand=None, force_a2=False, fused_add_norm=True,
                 split_head=False, dt_scale=0.0, n_dim_pos=4, num_layers=None, single_dir=False, embed_dims=96,
                 d_state=16):
        super().__init__()
        norm_cfg = dict(type="LN", eps=1e-6) if norm_cfg is None else norm_cfg
        if norm_cfg.get("type", "LN") != "LN":
            raise NotImplementedError("only LayerNorm (norm_cfg type 'LN') is implemented")
        if force_a2 or expand is not None:
            raise NotImplementedError("force_a2 / ssm_ratio are not Mamba arguments nnUZoo's MambaND2Net sets (:1137-1147)")
        self.spatial_dims, self.embed_dims, self.n_dim_pos = spatial_dims, embed_dims, n_dim_pos
        self.img_size, self.num_layers = img_size, num_layers
        self.patch_embed = PatchEmbed(spatial_dims, in_channels, embed_dims, kernel_size=patch_size, stride=patch_size,
                                      bias=not pre_norm)
        self.drop_after_pos = nn.Dropout(p=drop_rate)
        dpr = np.linspace(0, drop_path_rate, num_layers)
        ssm_cfg = {"d_state": d_state}
        if dt_scale > 0:
            ssm_cfg["dt_scale"] = dt_scale
        self.layers = nn.ModuleList(
            create_block(spatial_dims=spatial_dims, d_model=embed_dims, ssm_cfg=ssm_cfg, fused_add_norm=fused_add_norm,
                         residual_in_fp32=True, drop_rate=drop_rate, drop_path_rate=float(dpr[i]),
                         reverse=(not split_head) and (not single_dir) and (i % 2) > 0, rms_norm=False)
            for i in range(num_layers))
        eps = norm_cfg.get("eps", 1e-5)
        self.pre_norm = nn.LayerNorm(embed_dims, eps=eps) if pre_norm else nn.Identity()
        self.final_norm = final_norm
        self.ln1 = nn.LayerNorm(embed_dims, eps=eps) if final_norm else nn.Identity()

    def token_shape(self, patch_resolution: Sequence[int]):
        if self.spatial_dims == 3:  # (:964-965, :985) the reference repeats the second extent for the third
            return (patch_resolution[0], patch_resolution[1], patch_resolution[1])
        return (patch_resolution[0], patch_resolution[1])

    def forward_tokens(self, x, shape):
        """The Block stack on tokens x: (n, prod(shape), c) in "t h w" order (:989-1000)."""
        orders = self.ORDERS[self.spatial_dims]
        outs = []
        for i, blk in enumerate(self.layers):
            x = blk(x, order=orders[(i // 2) % len(orders)], shape=shape, n_dim_pos=self.n_dim_pos)
            if i == len(self.layers) - 1 and self.final_norm:
                x = self.ln1(x)
            outs.append(x)
        return outs[-1], outs

    def forward(self, x):
        x, patch_resolution = self.patch_embed(x)
        shape = self.token_shape(patch_resolution)
        if int(np.prod(shape)) != x.shape[1]:
            raise ValueError(f"token grid {tuple(shape)} does not match {x.shape[1]} tokens: the reference's MambaNDCore "
                             "takes the third extent from the second (mamba_nd2net.py:964-965), so H must equal W")
        x = self.pre_norm(self.drop_after_pos(x))
        return self.forward_tokens(x, shape)
