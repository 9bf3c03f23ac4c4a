"""Generate tests/golden/*.npz by running the REFERENCE ITSELF (read-only at /root/reference).

TEST INFRASTRUCTURE ONLY.  Run in the build container (the GPU box has no reference):

    python -m oracle.gen_golden

What is recorded (all seeds fixed, sizes tiny so the fixtures stay small):
  scan_*.npz  : inputs, ``selective_scan_ref`` output / last_state
                (nnunetv2/nets/seg_mamba/selective_scan_interface.py:86-152, verbatim), and the
                gradients torch.autograd produces through it for a random upstream gradient.
  cross_*.npz : what the reference's own ``SS2D.forward_core`` (m2net.py:170-206, sum :218) and
                ``SSND.forward_core`` (ssnd2net.py:239-302) feed to / make from the scan, captured by
                swapping ``self.selective_scan`` for a recorder that returns a fixed random out_y.
  module_*.npz: state_dict + input + output + input/parameter gradients of reference
                SS2D (m2net.py:39-225) and SSND 2-D / 3-D (ssnd2net.py:73-318) modules, with the
                reference's selective_scan_ref as the scan.
  module_mamba_*.npz: the vendored Mamba block (seg_mamba/mamba_simple.py:37-357), bimamba none / v2 / v3.
  module_m2net_64x32.npz: the whole reference ``M2Net`` (m2net.py:805-971; 80 SS2D scans) in eval mode on a
                (2, 1, 64, 32) input, parameters from oracle/fill.py (seeded, regenerated on both sides):
                input, the 7 deep-supervision outputs, input gradient, per-parameter gradient norms and a
                few whole parameter gradients.
  MANIFEST.json: case list plus the agreement of oracle/torch_port.py and oracle/scan_oracle.c with
                the verbatim reference at generation time.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from oracle import cross_oracle, ref_loader, scan_oracle, torch_port

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name: (batch, dim, groups (0 = 3-D B/C), dstate, L, has_D, has_z, has_bias, softplus, io dtype)
SCAN_CASES = {
    "scan_base_g2":       (2, 8, 2, 16, 64, True, False, True, True, "float32"),
    "scan_z_3d_L75":      (2, 6, 0, 16, 75, True, True, True, True, "float32"),
    "scan_plain_g4_L33":  (1, 4, 4, 16, 33, False, False, False, False, "float32"),
    "scan_long_z_L1100":  (1, 4, 1, 16, 1100, True, True, True, True, "float32"),
    "scan_bf16_z_L128":   (2, 8, 1, 16, 128, True, True, True, True, "bfloat16"),
    "scan_fp16_L96":      (1, 8, 2, 16, 96, True, False, True, True, "float16"),
    "scan_n8_L40":        (1, 4, 1, 8, 40, True, False, True, True, "float32"),
}


def make_scan_inputs(name, seed):
    bsz, dim, groups, n, L, has_D, has_z, has_bias, softplus, dt = SCAN_CASES[name]
    g = torch.Generator().manual_seed(seed)
    dtype = getattr(torch, dt)
    u = torch.randn(bsz, dim, L, generator=g).to(dtype)
    if softplus:
        delta = (0.5 * torch.randn(bsz, dim, L, generator=g)).to(dtype)
    else:
        delta = (0.001 + 0.1 * torch.rand(bsz, dim, L, generator=g)).to(dtype)
    # S4D-real init (m2net.py:144-149) with a trainable-style jitter so rows differ
    A = -(torch.arange(1, n + 1).float().repeat(dim, 1) * torch.exp(0.1 * torch.randn(dim, n, generator=g)))
    bshape = (bsz, n, L) if groups == 0 else (bsz, groups, n, L)
    B = torch.randn(*bshape, generator=g).to(dtype)
    C = torch.randn(*bshape, generator=g).to(dtype)
    D = (1.0 + 0.1 * torch.randn(dim, generator=g)) if has_D else None
    z = torch.randn(bsz, dim, L, generator=g).to(dtype) if has_z else None
    if has_bias:
        # inverse-softplus of log-uniform dt in [1e-3, 1e-1] (m2net.py:128-135)
        dtv = torch.exp(torch.rand(dim, generator=g) * (np.log(0.1) - np.log(0.001)) + np.log(0.001))
        bias = dtv + torch.log(-torch.expm1(-dtv))
    else:
        bias = None
    gout = torch.randn(bsz, dim, L, generator=g).to(dtype)
    return dict(u=u, delta=delta, A=A, B=B, C=C, D=D, z=z, delta_bias=bias, gout=gout), softplus


def _np(t):
    if t is None:
        return None
    t = t.detach()
    if t.dtype in (torch.bfloat16, torch.float16):
        # store 16-bit payloads losslessly as their fp32 images
        return t.float().numpy()
    return t.numpy()


def gen_scan(manifest):
    ref = ref_loader.selective_scan_ref()
    for seed, name in enumerate(SCAN_CASES):
        inp, softplus = make_scan_inputs(name, 1000 + seed)
        leaves = {k: (v.clone().requires_grad_(True) if v is not None and k != "gout" else v)
                  for k, v in inp.items()}
        out, last = ref(leaves["u"], leaves["delta"], leaves["A"], leaves["B"], leaves["C"],
                        leaves["D"], leaves["z"], leaves["delta_bias"], softplus, True)
        out.backward(inp["gout"])
        rec = {f"in_{k}": _np(v) for k, v in inp.items() if v is not None}
        rec["out"] = _np(out)
        rec["last_state"] = _np(last)
        for k in ("u", "delta", "A", "B", "C", "D", "z", "delta_bias"):
            if leaves[k] is not None:
                rec[f"grad_{k}"] = _np(leaves[k].grad)
        rec["meta_softplus"] = np.array(int(softplus))
        rec["meta_dtype"] = np.array(SCAN_CASES[name][-1])
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **rec)

        # agreement of the two restatements with the verbatim reference, recorded for the record
        with torch.no_grad():
            port = torch_port.selective_scan_port(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"],
                                                  inp["D"], inp["z"], inp["delta_bias"], softplus)
        c_out = scan_oracle.selective_scan_oracle(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"],
                                                  inp["D"], inp["z"], inp["delta_bias"], softplus)
        manifest["scan"][name] = dict(
            shape=SCAN_CASES[name][:5], dtype=SCAN_CASES[name][-1],
            torch_port_max_abs_diff=float((port.float() - out.float()).abs().max()),
            c_oracle_max_abs_diff=float(np.abs(c_out - _np(out)).max()),
            out_abs_max=float(out.float().abs().max()))


class _Recorder:
    """Stands in for ``self.selective_scan``: records xs, returns a fixed random out_y."""

    def __init__(self, seed):
        self.seed = seed
        self.xs = None
        self.out_y = None

    def __call__(self, xs, dts, As, Bs, Cs, Ds, z=None, delta_bias=None, delta_softplus=None,
                 return_last_state=None):
        self.xs = xs.detach().clone()
        g = torch.Generator().manual_seed(self.seed)
        self.out_y = torch.randn(xs.shape, generator=g)
        return self.out_y.clone()


def gen_cross(manifest):
    m2 = ref_loader.m2net()
    sn = ref_loader.ssnd2net()
    torch.manual_seed(7)
    # ---- 2-D through m2net.SS2D.forward_core (:170-206) and the :218 sum ----
    for name, (bsz, H, W, dm) in {"cross_2d_a": (2, 4, 6, 4), "cross_2d_sq": (1, 8, 8, 2),
                                  "cross_2d_odd": (1, 3, 5, 2)}.items():
        mod = m2.SS2D(d_model=dm).eval()
        rec = _Recorder(11)
        mod.selective_scan = rec
        x = torch.randn(bsz, mod.d_inner, H, W)
        with torch.no_grad():
            y1, y2, y3, y4 = mod.forward_core(x)
            y = y1 + y2 + y3 + y4                                    # m2net.py:218
        K = 4
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), x=_np(x),
                            xs=_np(rec.xs.view(bsz, K, -1, H * W)),
                            out_y=_np(rec.out_y.view(bsz, K, -1, H * W)), y=_np(y))
        ours_xs = cross_oracle.cross_scan_2d(_np(x))
        ours_y = cross_oracle.cross_merge_2d(_np(rec.out_y.view(bsz, K, -1, H * W)), H, W)
        manifest["cross"][name] = dict(shape=[bsz, mod.d_inner, H, W],
                                       oracle_scan_equal=bool(np.array_equal(ours_xs, _np(rec.xs.view(bsz, K, -1, H * W)))),
                                       oracle_merge_equal=bool(np.array_equal(ours_y, _np(y))))
    # ---- 2-D and 3-D through ssnd2net.SSND.forward_core (:239-302) ----
    for name, (sd, bsz, dims, dm) in {"cross_ssnd2d": (2, 1, (3, 4), 2),
                                      "cross_3d_a": (3, 1, (2, 3, 5), 2),
                                      "cross_3d_b": (3, 2, (4, 2, 3), 2)}.items():
        mod = sn.SSND(spatial_dims=sd, factorization_type="cross-scan", d_model=dm).eval()
        rec = _Recorder(13)
        mod.selective_scan = rec
        x = torch.randn(bsz, mod.d_inner, *dims)
        with torch.no_grad():
            y = mod.forward_core(x)          # (B, *dims, D)
        K = mod.k
        L = int(np.prod(dims))
        y_bdl = y.reshape(bsz, L, -1).transpose(1, 2).contiguous()   # undo :284 / :299 transpose
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), x=_np(x),
                            xs=_np(rec.xs.view(bsz, K, -1, L)), out_y=_np(rec.out_y.view(bsz, K, -1, L)),
                            y=_np(y_bdl), dims=np.array(dims))
        if sd == 2:
            ours_xs = cross_oracle.cross_scan_2d(_np(x))
            ours_y = cross_oracle.cross_merge_2d(_np(rec.out_y.view(bsz, K, -1, L)), *dims)
        else:
            ours_xs = cross_oracle.cross_scan_3d(_np(x))
            ours_y = cross_oracle.cross_merge_3d(_np(rec.out_y.view(bsz, K, -1, L)), *dims)
        manifest["cross"][name] = dict(shape=[bsz, mod.d_inner, *dims],
                                       oracle_scan_equal=bool(np.array_equal(ours_xs, _np(rec.xs.view(bsz, K, -1, L)))),
                                       oracle_merge_equal=bool(np.array_equal(ours_y, _np(y_bdl))))


def gen_module(manifest):
    m2 = ref_loader.m2net()
    sn = ref_loader.ssnd2net()
    specs = {
        "module_ss2d_m8": (lambda: m2.SS2D(d_model=8), (2, 6, 5, 8)),
        "module_ss2d_m32": (lambda: m2.SS2D(d_model=32), (1, 8, 8, 32)),
        "module_ssnd2d_m8": (lambda: sn.SSND(spatial_dims=2, factorization_type="cross-scan", d_model=8), (1, 4, 6, 8)),
        "module_ssnd3d_m8": (lambda: sn.SSND(spatial_dims=3, factorization_type="cross-scan", d_model=8), (1, 2, 3, 5, 8)),
    }
    for i, (name, (ctor, xshape)) in enumerate(specs.items()):
        torch.manual_seed(100 + i)
        mod = ctor().eval()
        # move the trainable scan parameters off their symmetric init so every direction differs
        with torch.no_grad():
            mod.A_logs.add_(0.1 * torch.randn_like(mod.A_logs))
            mod.Ds.add_(0.1 * torch.randn_like(mod.Ds))
        x = torch.randn(*xshape, requires_grad=True)
        y = mod(x)
        gy = torch.randn_like(y)
        y.backward(gy)
        rec = {"x": _np(x), "y": _np(y), "gy": _np(gy), "gx": _np(x.grad)}
        for k, v in mod.state_dict().items():
            rec["sd_" + k] = _np(v)
        for k, p in mod.named_parameters():
            rec["gp_" + k] = _np(p.grad) if p.grad is not None else np.zeros(tuple(p.shape), np.float32)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **rec)
        manifest["module"][name] = dict(x_shape=list(xshape), y_abs_mean=float(y.abs().mean()))


def gen_mamba(manifest):
    """module_mamba_*.npz: the reference's vendored Mamba block (mamba_simple.py:37-357), uni-, bi- (v2) and
    tri-directional (v3), fast path (fused call replaced by its op-for-op CPU stand-in, ref_loader.mamba_simple)
    and, for "none", also the reference's own slow path -- the two must agree."""
    ms = ref_loader.mamba_simple()
    specs = {"module_mamba_none": ("none", (2, 24, 16)), "module_mamba_v2": ("v2", (2, 20, 16)),
             "module_mamba_v3": ("v3", (1, 30, 32))}
    for i, (name, (kind, xshape)) in enumerate(specs.items()):
        torch.manual_seed(300 + i)
        mod = ms.Mamba(d_model=xshape[-1], bimamba_type=kind, nslices=5).eval()
        with torch.no_grad():  # move the scan parameters off their symmetric init
            for p in (mod.A_log, mod.A_b_log, mod.A_s_log, mod.D, mod.D_b, mod.D_s):
                p.add_(0.1 * torch.randn_like(p))
        x = torch.randn(*xshape, requires_grad=True)
        y = mod(x)
        if kind == "none":
            mod.use_fast_path = False
            y_slow = mod(x)
            mod.use_fast_path = True
            assert float((y - y_slow).abs().max()) < 1e-5 * float(y.abs().max() + 1e-9)
        gy = torch.randn_like(y)
        y.backward(gy)
        rec = {"x": _np(x), "y": _np(y), "gy": _np(gy), "gx": _np(x.grad)}
        for k, v in mod.state_dict().items():
            rec["sd_" + k] = _np(v)
        for k, p in mod.named_parameters():
            rec["gp_" + k] = _np(p.grad) if p.grad is not None else np.zeros(tuple(p.shape), np.float32)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **rec)
        manifest["module"][name] = dict(x_shape=list(xshape), y_abs_mean=float(y.abs().mean()), bimamba_type=kind)


def _record(mod, x, fwd, name, manifest, extra=None):
    y = fwd(mod, x)
    gy = torch.randn_like(y)
    y.backward(gy)
    rec = {"x": _np(x), "y": _np(y), "gy": _np(gy), "gx": _np(x.grad)}
    for k, v in mod.state_dict().items():
        rec["sd_" + k] = _np(v)
    for k, p in mod.named_parameters():
        rec["gp_" + k] = _np(p.grad) if p.grad is not None else np.zeros(tuple(p.shape), np.float32)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **rec)
    manifest["module"][name] = dict(x_shape=list(x.shape), y_abs_mean=float(y.abs().mean()), **(extra or {}))


def _perturb_mamba(mod):
    with torch.no_grad():  # move the scan parameters off their symmetric init
        for k, p in mod.named_parameters():
            if k.endswith("A_log") or k.endswith(".D") or k == "D":
                p.add_(0.1 * torch.randn_like(p))
            if k.endswith("skip_scale"):
                p.add_(0.3)


def gen_shells(manifest):
    """shell_*.npz: the callers of the 1-D path run by the reference itself -- MambaLayer / ResMambaBlock with its three
    axis orders (lm2net.py:64-176), MambaND's Block for every order x direction and a 7-layer MambaNDCore
    (mamba_nd2net.py:565-666, :725-1001; fused_add_norm=False, final_norm=False as MambaND2Net builds it, :1128-1147)."""
    lm = ref_loader.lm2net()
    nd = ref_loader.mamba_nd2net()
    torch.manual_seed(700)
    mod = lm.MambaLayer(input_dim=16, output_dim=24).eval()
    _perturb_mamba(mod)
    _record(mod, torch.randn(2, 16, 3, 4, 5, requires_grad=True), lambda m, x: m(x), "shell_mambalayer", manifest)
    specs = {"shell_resmamba_dhw": (3, "d h w", (1, 16, 3, 4, 5)), "shell_resmamba_dwh": (3, "d w h", (1, 16, 3, 4, 5)),
             "shell_resmamba_whd": (3, "w h d", (1, 16, 3, 4, 5)), "shell_resmamba_wh": (2, "w h", (2, 16, 5, 6))}
    for i, (name, (sd, order, xs)) in enumerate(specs.items()):
        torch.manual_seed(710 + i)
        mod = lm.ResMambaBlock(sd, 16, norm=("GROUP", {"num_groups": 8}), order=order).eval()
        _perturb_mamba(mod)
        _record(mod, torch.randn(*xs, requires_grad=True), lambda m, x: m(x), name, manifest,
                dict(order=order, spatial_dims=sd))
    shape = (2, 3, 4)
    for i, order in enumerate(("t h w", "t w h", "w h t")):
        for rev in (False, True):
            torch.manual_seed(730 + 2 * i + int(rev))
            blk = nd.create_block(spatial_dims=3, d_model=16, ssm_cfg={"d_state": 16}, fused_add_norm=False,
                                  residual_in_fp32=True, reverse=rev, drop_rate=0.0, drop_path_rate=0.0).eval()
            _perturb_mamba(blk)
            name = f"shell_ndblock_{order.replace(' ', '')}_{'rev' if rev else 'fwd'}"
            _record(blk, torch.randn(2, 24, 16, requires_grad=True),
                    lambda m, x, o=order: m(x, order=o, shape=shape, n_dim_pos=4), name, manifest,
                    dict(order=order, reverse=rev, shape=list(shape)))
    torch.manual_seed(750)
    core = nd.MambaNDCore(spatial_dims=3, img_size=(4, 6, 6), patch_size=(2, 2, 2), in_channels=2, embed_dims=16,
                          num_layers=7, fused_add_norm=False, final_norm=False, drop_rate=0.0, drop_path_rate=0.0).eval()
    _perturb_mamba(core)
    _record(core, torch.randn(2, 2, 4, 6, 6, requires_grad=True), lambda m, x: m(x)[0], "shell_ndcore", manifest,
            dict(num_layers=7, patch_size=[2, 2, 2]))
    torch.manual_seed(751)
    core2 = nd.MambaNDCore(spatial_dims=2, img_size=(8, 8), patch_size=(2, 2), in_channels=1, embed_dims=16,
                           num_layers=4, fused_add_norm=False, final_norm=False, drop_rate=0.0, drop_path_rate=0.0).eval()
    _perturb_mamba(core2)
    _record(core2, torch.randn(2, 1, 8, 8, requires_grad=True), lambda m, x: m(x)[0], "shell_ndcore2d", manifest,
            dict(num_layers=4, patch_size=[2, 2]))


M2NET_FULL_GRADS = ("stage1.vssm_encoder.layers.0.blocks.0.self_attention.A_logs",
                    "stage1.vssm_encoder.layers.0.blocks.0.self_attention.x_proj_weight",
                    "stage1d.vssm_decoder.stages.5.blocks.0.self_attention.dt_projs_bias",
                    "stage3.vssm_encoder.layers.2.blocks.0.self_attention.Ds",
                    "stage4d.vssm_decoder.concat_back_dim.0.weight", "side3.weight", "outconv.weight")


def gen_m2net(manifest):
    from oracle.fill import deterministic_fill
    m2 = ref_loader.m2net()
    torch.manual_seed(500)
    net = m2.M2Net(1, 4, True).eval()
    deterministic_fill(net, 7)
    x = torch.randn(2, 1, 64, 32, requires_grad=True)
    outs = net(x)
    gys = [torch.randn_like(o) for o in outs]
    torch.autograd.backward(list(outs), gys)
    rec = {"x": _np(x), "gx": _np(x.grad)}
    for i, (o, g) in enumerate(zip(outs, gys)):
        rec[f"d{i}"] = _np(o)
        rec[f"gd{i}"] = _np(g)
    names, norms = [], []
    for k, p in sorted(net.named_parameters()):
        names.append(k)
        norms.append(0.0 if p.grad is None else float(p.grad.double().norm()))
    rec["grad_norm_names"] = np.array(names)
    rec["grad_norms"] = np.array(norms, np.float64)
    params = dict(net.named_parameters())
    for k in M2NET_FULL_GRADS:
        rec["gp_" + k] = _np(params[k].grad)
    np.savez_compressed(os.path.join(GOLD, "module_m2net_64x32.npz"), **rec)
    manifest["module"]["module_m2net_64x32"] = dict(
        x_shape=[2, 1, 64, 32], fill_seed=7, n_params=int(sum(p.numel() for p in net.parameters())),
        unused_params=int(sum(1 for n in norms if n == 0.0)), d0_abs_mean=float(outs[0].abs().mean()))


def main():
    assert ref_loader.available(), "run in the build container: /root/reference is required"
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    manifest = {"generator": "python -m oracle.gen_golden", "torch": torch.__version__,
                "reference": "AI-in-Cardiovascular-Medicine/nnUZoo @ /root/reference (read-only mount)",
                "scan": {}, "cross": {}, "module": {}}
    gen_scan(manifest)
    gen_cross(manifest)
    gen_module(manifest)
    gen_mamba(manifest)
    gen_m2net(manifest)
    gen_shells(manifest)
    with open(os.path.join(GOLD, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print(json.dumps(manifest, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
