"""INTEGRATION.md, Option A: shadow ``mamba_ssm`` with this package BEFORE importing the reference's nets, and the
reference's own classes bind our operator by name (m2net.py:11, :107; lm2net.py:14).  Container-only (needs
/root/reference; it never exists on the GPU box) and CPU-only: it checks the binding, not the arithmetic -- that is what
the GPU parity suite does.  Runs in a subprocess so the shadow modules do not leak into other tests."""
import os
import subprocess
import sys

import pytest

REF = os.environ.get("NNUZOO_REFERENCE_ROOT", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CODE = r'''
import sys, types
sys.path.insert(0, %(root)r)
import nnuzoo_b200
from nnuzoo_b200 import selective_scan_interface as nz_ssi

# --- the shim of INTEGRATION.md, Option A ---
ops = types.ModuleType("mamba_ssm.ops.selective_scan_interface")
ops.selective_scan_fn = nz_ssi.selective_scan_fn
ops.SelectiveScanFn = nz_ssi.SelectiveScanFn
ops.mamba_inner_fn = nnuzoo_b200.mamba_inner_fn
pkg, sub = types.ModuleType("mamba_ssm"), types.ModuleType("mamba_ssm.ops")
pkg.ops, sub.selective_scan_interface = sub, ops
pkg.Mamba = lambda *a, **k: nnuzoo_b200.Mamba(*a, extra_directions=False, **k)
sys.modules.update({"mamba_ssm": pkg, "mamba_ssm.ops": sub, "mamba_ssm.ops.selective_scan_interface": ops})

# third-party packages the reference files import but this path never executes (stubs from the test infrastructure)
from oracle import ref_loader
ref_loader._STUB_ROOTS = tuple(r for r in ref_loader._STUB_ROOTS if r != "mamba_ssm")
import importlib.abc, importlib.machinery
sys.meta_path.insert(0, ref_loader._StubFinder())
import timm.layers as tl, torch.nn as nn
tl.DropPath = nn.Identity
tl.trunc_normal_ = lambda t, std=1.0, **kw: nn.init.trunc_normal_(t, std=std, **kw)
import dynamic_network_architectures.initialization.weight_init as wi
wi.init_last_bn_before_add_to_0 = lambda m: None

import importlib.util
spec = importlib.util.spec_from_file_location("ref_m2net_shim", %(ref)r + "/nnunetv2/nets/m2net.py")
m2 = importlib.util.module_from_spec(spec)
spec.loader.exec_module(m2)
assert m2.selective_scan_fn is nz_ssi.selective_scan_fn, "the reference module must have bound OUR operator"
blk = m2.SS2D(d_model=8)
assert blk.selective_scan is nz_ssi.selective_scan_fn            # m2net.py:107
ours = nnuzoo_b200.SS2D(d_model=8)
ours.load_state_dict(blk.state_dict(), strict=True)              # and the block itself swaps one for one
print("OK")
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "nnunetv2", "nets")), reason="reference not mounted")
def test_option_a_shim_binds_our_operator_into_the_reference_module():
    r = subprocess.run([sys.executable, "-c", CODE % {"root": ROOT, "ref": REF}], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stderr[-2000:]
