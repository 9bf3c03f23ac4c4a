"""SS2D^2-Net (``M2Net``): the caller that drives 80 SS2D scans per training step (BASELINE config 2).

Mirror of the reference network ``nnunetv2/nets/m2net.py:805-971`` (nested U of Mamba-U blocks ``MU``
:713-766, VSS encoder :598-710, VSS decoder :359-483, patch merge / expand :228-319, RSU-4F :769-802),
with every ``SS2D`` being ``nnuzoo_b200.SS2D`` -- i.e. the sm_100a scan / CrossScan / CrossMerge kernels.
Module and parameter names are the reference's, so ``load_state_dict(reference_ckpt, strict=True)`` works
(tests/golden/module_m2net_*.npz pins forward outputs and gradients against the reference itself).

Only the module tree is new; the convolutions, LayerNorms and Linear layers around the scan stay the
library ops they are in the reference (cuDNN / cuBLAS through PyTorch) -- SURVEY.md section 8(f) rank 1
is where fusing them with the scan is scheduled.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .norm import LayerNorm
from .ss2d import SS2D

__all__ = ["M2Net", "MU", "get_m2net"]


class DropPath(nn.Module):
    """Per-sample stochastic depth (timm.layers.DropPath as used at m2net.py:526)."""

    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)

    def forward(self, x):
        if not self.training or self.drop_prob == 0.0:
            return x
        keep = 1.0 - self.drop_prob
        mask = torch.empty((x.shape[0],) + (1,) * (x.dim() - 1), dtype=x.dtype, device=x.device).bernoulli_(keep)
        return x * mask.div_(keep)


def _to_cl(x):   # (B, C, H, W) -> (B, H, W, C)
    return x.permute(0, 2, 3, 1)


def _to_cf(x):   # (B, H, W, C) -> (B, C, H, W)
    return x.permute(0, 3, 1, 2)


def _pixel_unshuffle_cl(x, s):
    """Channels-last space-to-depth in the reference's order (m2net.py:256-266): the s*s sub-grids are
    concatenated as (row offset 0, col 0), (1, 0), (0, 1), (1, 1) for s = 2."""
    B, H, W, C = x.shape
    h, w = H // s, W // s
    x = x[:, :h * s, :w * s].reshape(B, h, s, w, s, C)          # (B, h, i, w, j, C)
    return x.permute(0, 1, 3, 4, 2, 5).reshape(B, h, w, s * s * C)   # channel blocks ordered (j, i)


def _pixel_shuffle_cl(x, s):
    """'b h w (p1 p2 c) -> b (h p1) (w p2) c' (m2net.py:300, 311)."""
    B, H, W, C = x.shape
    c = C // (s * s)
    return x.reshape(B, H, W, s, s, c).permute(0, 1, 3, 2, 4, 5).reshape(B, H * s, W * s, c)


class REBNCONV(nn.Module):
    """conv3x3 (dilated) + BatchNorm + ReLU (m2net.py:18-30)."""

    def __init__(self, in_ch=3, out_ch=3, dirate=1):
        super().__init__()
        self.conv_s1 = nn.Conv2d(in_ch, out_ch, 3, padding=dirate, dilation=dirate)
        self.bn_s1 = nn.BatchNorm2d(out_ch)
        self.relu_s1 = nn.ReLU(inplace=True)

    def forward(self, x):
        return self.relu_s1(self.bn_s1(self.conv_s1(x)))


class PatchMerging2D(nn.Module):
    """Space-to-depth, LayerNorm, Linear (m2net.py:228-273)."""

    def __init__(self, input_dim, scale, output_features=None, norm_layer=LayerNorm):
        super().__init__()
        self.scale = scale
        self.input_feature_size = scale * scale * input_dim
        self.output_features = output_features or input_dim * scale
        self.reduction = nn.Linear(self.input_feature_size, self.output_features, bias=False)
        self.norm = norm_layer(self.input_feature_size)
        self.norm.feeds_linear = True      # only consumer: self.reduction

    def forward(self, x, permute=False):
        if permute:
            x = _to_cl(x)
        if self.scale == 2:
            x = _pixel_unshuffle_cl(x, 2)
        else:  # the reference gathers exactly four sub-grids whatever the scale (m2net.py:256-259)
            s = self.scale
            h, w = x.shape[1] // s, x.shape[2] // s
            x = torch.cat([x[:, i::s, j::s][:, :h, :w] for j in (0, 1) for i in (0, 1)], -1)
        x = self.reduction(self.norm(x))
        return _to_cf(x).contiguous() if permute else x


class PatchExpand(nn.Module):
    """Depth-to-space up-sampling (m2net.py:276-319); takes (B, C, H, W), returns channels-last unless
    ``permute``. With ``output_dim`` the shuffle comes first, then Linear(dim / scale^2 -> output_dim)."""

    def __init__(self, dim, scale, output_dim=None, norm_layer=LayerNorm):
        super().__init__()
        self.dim, self.scale, self.output_dim = dim, scale, output_dim
        if output_dim is None:
            self.expand = nn.Linear(dim, scale * dim, bias=False)
            self.norm = norm_layer(dim // scale)
        else:
            self.expand = nn.Linear(dim // (scale * scale), output_dim, bias=False)
            self.norm = norm_layer(output_dim)

    def forward(self, x, permute=False):
        x = _to_cl(x)
        if self.output_dim is None:
            x = self.norm(_pixel_shuffle_cl(self.expand(x), self.scale))
        else:
            x = self.norm(self.expand(_pixel_shuffle_cl(x, self.scale)))
        return _to_cf(x).contiguous() if permute else x


class VSSBlock(nn.Module):
    """x + DropPath(SS2D(LayerNorm(x)))  (m2net.py:513-530)."""

    def __init__(self, hidden_dim, drop_path=0.0, norm_layer=LayerNorm, attn_drop_rate=0.0, d_state=16, **kw):
        super().__init__()
        self.ln_1 = norm_layer(hidden_dim)
        self.ln_1.feeds_linear = True      # only consumer: SS2D.in_proj
        self.self_attention = SS2D(d_model=hidden_dim, dropout=attn_drop_rate, d_state=d_state, **kw)
        self.drop_path = DropPath(drop_path)

    def forward(self, x):
        return x + self.drop_path(self.self_attention(self.ln_1(x)))


class VSSLayer(nn.Module):
    """``depth`` VSSBlocks (+ optional downsample), m2net.py:533-595."""

    def __init__(self, dim, depth, attn_drop=0.0, drop_path=0.0, norm_layer=LayerNorm, downsample=None,
                 use_checkpoint=False, d_state=16):
        super().__init__()
        self.dim = dim
        self.use_checkpoint = use_checkpoint
        rates = drop_path if isinstance(drop_path, (list, tuple)) else [drop_path] * depth
        self.blocks = nn.ModuleList(VSSBlock(dim, rates[i], norm_layer, attn_drop, d_state) for i in range(depth))
        self.downsample = downsample(dim=dim, norm_layer=norm_layer) if downsample is not None else None

    def forward(self, x):
        for blk in self.blocks:
            x = torch.utils.checkpoint.checkpoint(blk, x) if self.use_checkpoint else blk(x)
        return x if self.downsample is None else self.downsample(x)


class PatchEmbed2D(nn.Module):
    """Strided conv to channels-last tokens (m2net.py:486-510)."""

    def __init__(self, patch_size=4, in_chans=3, embed_dim=96, norm_layer=None):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = norm_layer(embed_dim) if norm_layer is not None else None

    def forward(self, x):
        x = _to_cl(self.proj(x))
        return x if self.norm is None else self.norm(x)


def _encoder_init(m):
    """m2net.py:663-679: Linear ~ trunc_normal(0.02) / zero bias, LayerNorm = identity; convs untouched."""
    if isinstance(m, nn.Linear):
        nn.init.trunc_normal_(m.weight, std=0.02)
        if m.bias is not None:
            nn.init.zeros_(m.bias)
    elif isinstance(m, nn.LayerNorm):
        nn.init.ones_(m.weight)
        nn.init.zeros_(m.bias)


class VSSMEncoder(nn.Module):
    """m2net.py:598-710. Returns [stem output or None, stage outputs (B, C, H, W) ...]."""

    def __init__(self, patch_size=4, in_chans=3, depths=(2, 2, 9, 2), dims=(96, 192, 384, 768), d_state=16,
                 drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.1, norm_layer=LayerNorm, patch_norm=True,
                 use_checkpoint=False, skip_first_downsample=False, skip_last_downsample=False, add_last=False,
                 out_ch=None):
        super().__init__()
        n = len(depths)
        if isinstance(dims, int):
            dims = [dims * 2 ** i for i in range(n)]
        self.num_layers, self.dims, self.embed_dim = n, list(dims), dims[0]
        self.add_last = add_last
        self.skip_first_downsample, self.skip_last_downsample = skip_first_downsample, skip_last_downsample
        if add_last:
            self.rebnconvin = REBNCONV(in_chans, out_ch, dirate=1)
        self.patch_embed = PatchEmbed2D(patch_size, out_ch if add_last else in_chans, dims[0],
                                        norm_layer if patch_norm else None)
        self.pos_drop = nn.Dropout(p=drop_rate)
        rates = torch.linspace(0, drop_path_rate, sum(depths)).tolist()
        self.layers = nn.ModuleList()
        self.downsamples = nn.ModuleList()
        for i in range(n):
            lo = sum(depths[:i])
            self.layers.append(VSSLayer(dims[i], depths[i], attn_drop_rate, rates[lo:lo + depths[i]], norm_layer,
                                        None, use_checkpoint, d_state if d_state is not None else -(-dims[0] // 6)))
            skip = (i == 0 and skip_first_downsample) or (i == n - 2 and skip_last_downsample)
            if i < n - 1 and not skip:
                self.downsamples.append(PatchMerging2D(dims[i], 2, dims[i + 1], norm_layer))
        self.apply(_encoder_init)

    def forward(self, x):
        outs = []
        if self.add_last:
            x = self.rebnconvin(x)
            outs.append(x)
        else:
            outs.append(None)
        x = self.pos_drop(self.patch_embed(x))
        for s, layer in enumerate(self.layers):
            x = layer(x)
            outs.append(_to_cf(x))
            if s < len(self.downsamples) and not (s == 0 and self.skip_first_downsample):
                x = self.downsamples[s](x)
        return outs


class VSSMDecoder(nn.Module):
    """m2net.py:359-483: bottom-up expand -> concat skip -> Linear -> VSSLayer; 1x1 seg heads."""

    def __init__(self, num_classes, deep_supervision, features_per_stage=None, depths=None, drop_path_rate=0.2,
                 d_state=16, skip_first_expand=False, patch_size=4):
        super().__init__()
        f = list(features_per_stage)
        n = len(f)
        depths = list(depths or [2] * n)
        self.skip_first_expand, self.deep_supervision, self.num_classes = skip_first_expand, deep_supervision, num_classes
        rates = torch.linspace(drop_path_rate, 0, (n - 1) * 2).tolist()
        stages, expands, segs, backs = [], [], [], []
        skip_ch = f[0]
        for s in range(1, n):
            below, skip_ch = f[-s], f[-(s + 1)]
            expands.append(None if (s == 1 and skip_first_expand) else PatchExpand(below, 2, below))
            stages.append(VSSLayer(skip_ch, 1, 0.0, rates[sum(depths[:s - 1]):sum(depths[:s])], LayerNorm, None,
                                   False, d_state if d_state is not None else -(-2 * skip_ch // 6)))
            segs.append(nn.Conv2d(skip_ch, num_classes, 1, 1, 0, bias=True))
            backs.append(nn.Linear(2 * skip_ch, skip_ch))
        expands.append(PatchExpand(f[0], patch_size))
        stages.append(nn.Identity())
        segs.append(nn.Conv2d(skip_ch, num_classes, 1, 1, 0, bias=True))
        self.stages = nn.ModuleList(stages)
        self.expand_layers = nn.ModuleList(expands)
        self.seg_layers = nn.ModuleList(segs)
        self.concat_back_dim = nn.ModuleList(backs)

    def forward(self, skips):
        low = skips[-1]
        last = len(self.stages) - 1
        segs = []
        for s, stage in enumerate(self.stages):
            x = _to_cl(low) if (s == 0 and self.skip_first_expand) else self.expand_layers[s](low)
            if s < last:
                x = self.concat_back_dim[s](torch.cat((x, _to_cl(skips[-(s + 2)])), -1))
            x = _to_cf(stage(x))
            if self.deep_supervision:
                segs.append(self.seg_layers[s](x))
            elif s == last:
                segs.append(self.seg_layers[-1](x))
            low = x
        segs.reverse()
        return segs if self.deep_supervision else segs[0]


class MU(nn.Module):
    """Mamba-U block: VSS encoder + decoder of ``n_layers`` levels with a residual stem (m2net.py:713-766)."""

    def __init__(self, in_ch, mid_ch, out_ch, n_layers, skip_last_downsample=False, patch_size=4, add_last=False):
        super().__init__()
        self.add_last = add_last
        feats, depths = [mid_ch] * n_layers, [1] * n_layers
        self.vssm_encoder = VSSMEncoder(patch_size=patch_size, in_chans=in_ch, depths=depths, dims=feats,
                                        skip_first_downsample=False, skip_last_downsample=skip_last_downsample,
                                        add_last=add_last, out_ch=out_ch if add_last else None, drop_path_rate=0.2)
        self.vssm_decoder = VSSMDecoder(num_classes=out_ch, deep_supervision=False, features_per_stage=feats,
                                        drop_path_rate=0.2, d_state=16, depths=depths,
                                        skip_first_expand=skip_last_downsample, patch_size=patch_size)

    def forward(self, x):
        skips = self.vssm_encoder(x)
        out = self.vssm_decoder(skips)
        return out + skips[0] if self.add_last else out

    @torch.no_grad()
    def freeze_encoder(self):
        for name, p in self.vssm_encoder.named_parameters():
            if "patch_embed" not in name:
                p.requires_grad = False

    @torch.no_grad()
    def unfreeze_encoder(self):
        for p in self.vssm_encoder.parameters():
            p.requires_grad = True


class RSU4F(nn.Module):
    """Dilated residual U block without pooling (m2net.py:769-802)."""

    def __init__(self, in_ch=3, mid_ch=12, out_ch=3):
        super().__init__()
        self.rebnconvin = REBNCONV(in_ch, out_ch, 1)
        self.rebnconv1 = REBNCONV(out_ch, mid_ch, 1)
        self.rebnconv2 = REBNCONV(mid_ch, mid_ch, 2)
        self.rebnconv3 = REBNCONV(mid_ch, mid_ch, 4)
        self.rebnconv4 = REBNCONV(mid_ch, mid_ch, 8)
        self.rebnconv3d = REBNCONV(2 * mid_ch, mid_ch, 4)
        self.rebnconv2d = REBNCONV(2 * mid_ch, mid_ch, 2)
        self.rebnconv1d = REBNCONV(2 * mid_ch, out_ch, 1)

    def forward(self, x):
        xin = self.rebnconvin(x)
        e1 = self.rebnconv1(xin)
        e2 = self.rebnconv2(e1)
        e3 = self.rebnconv3(e2)
        d = self.rebnconv4(e3)
        for up, skip in ((self.rebnconv3d, e3), (self.rebnconv2d, e2), (self.rebnconv1d, e1)):
            d = up(torch.cat((d, skip), 1))
        return d + xin


def _resize(src, size):
    return F.interpolate(src, size=size, mode="bilinear")   # F.upsample(..., 'bilinear'), m2net.py:33-36


# (name suffix, in_ch of the MU, mid_ch, out_ch, n_layers): m2net.py:810-823 (encoder) / :841-868 (decoder)
_MU_SPECS = ((1, 16, 32, 7), (2, 32, 64, 6), (3, 64, 128, 5), (4, 128, 256, 4))


class M2Net(nn.Module):
    """m2net.py:805-971.  forward((B, in_ch, H, W)) -> d0 or (d0 .. d6) under deep supervision, d_i at
    H / 2^max(i-1, 0).  H and W must be multiples of 32 for the skip shapes to line up (512 in config 2)."""

    def __init__(self, in_ch: int, out_ch: int, deep_supervision: bool):
        super().__init__()
        self.deep_supervision = deep_supervision
        chan_in = in_ch
        for i, mid, out, nl in _MU_SPECS:
            setattr(self, f"stage{i}", MU(chan_in, mid, out, nl, skip_last_downsample=True, patch_size=1,
                                          add_last=True))
            setattr(self, f"patch_merging{i}", PatchMerging2D(out, scale=2))
            chan_in = 2 * out
        self.stage5 = RSU4F(512, 256, 512)
        self.pool56 = nn.MaxPool2d(2, stride=2, ceil_mode=True)
        self.stage6 = RSU4F(512, 256, 512)
        self.stage5d = RSU4F(1024, 256, 512)
        for i, mid, out, nl in reversed(_MU_SPECS):
            setattr(self, f"patch_expand{i}d", PatchExpand(dim=2 * out, scale=2))
            setattr(self, f"concat_back_dim{i}d", nn.Linear(2 * out, out))
            setattr(self, f"stage{i}d", MU(out, mid, out, nl, skip_last_downsample=True, patch_size=1,
                                           add_last=True))
        for i, ch in enumerate((32, 64, 128, 256, 512, 512), start=1):
            setattr(self, f"side{i}", nn.Conv2d(ch, out_ch, 3, padding=1))
        self.outconv = nn.Conv2d(6 * out_ch, out_ch, 1)

    def forward(self, x):
        enc = []
        h = x
        for i in (1, 2, 3, 4):
            e = getattr(self, f"stage{i}")(h)
            enc.append(e)
            h = getattr(self, f"patch_merging{i}")(e, permute=True)
        e5 = self.stage5(h)
        e6 = self.stage6(self.pool56(e5))
        dec = {6: e6, 5: self.stage5d(torch.cat((_resize(e6, e5.shape[2:]), e5), 1))}
        for i in (4, 3, 2, 1):
            up = getattr(self, f"patch_expand{i}d")(dec[i + 1])                       # channels-last
            up = getattr(self, f"concat_back_dim{i}d")(torch.cat((up, _to_cl(enc[i - 1])), -1))
            dec[i] = getattr(self, f"stage{i}d")(_to_cf(up))
        sides = [getattr(self, f"side{i}")(dec[i]) for i in range(1, 7)]
        full = sides[0].shape[2:]
        d0 = self.outconv(torch.cat([sides[0]] + [_resize(s, full) for s in sides[1:]], 1))
        return (d0, *sides) if self.deep_supervision else d0

    def _encoder_groups(self):
        return [getattr(self, n) for n in ("stage1", "stage2", "stage3", "stage4", "stage5", "stage6",
                                           "patch_merging1", "patch_merging2", "patch_merging3", "patch_merging4")]

    @torch.no_grad()
    def freeze_encoder(self):
        for g in self._encoder_groups():
            for p in g.parameters():
                p.requires_grad = False

    @torch.no_grad()
    def unfreeze_encoder(self):
        for g in self._encoder_groups():
            for p in g.parameters():
                p.requires_grad = True


def get_m2net(num_input_channels: int, num_classes: int, deep_supervision: bool = True) -> M2Net:
    """What ``get_m2net_from_plans`` (m2net.py:1187-1208) reduces to once the plans are read: it only takes the
    channel count, the number of segmentation heads and the deep-supervision flag from them."""
    model = M2Net(num_input_channels, num_classes, deep_supervision)
    for m in model.modules():   # InitWeights_He(1e-2), utilities/network_initialization.py:4-12
        if isinstance(m, (nn.Conv2d, nn.Conv3d, nn.ConvTranspose2d, nn.ConvTranspose3d)):
            nn.init.kaiming_normal_(m.weight, a=1e-2)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
    return model
