"""Host-side logic of bench.py: workload definition, byte accounting, and the N > 1 timing reduction
(world_size 2 over gloo on CPU -- the GPU run uses the same code over NCCL)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_scan_list_matches_the_survey_totals():
    import bench
    scans = bench.m2net_scan_list(512)
    assert len(scans) == 80                                            # 80 calls per forward
    assert sum(L for _, L in scans) == 1_857_536                       # 8a: sum L per sample
    assert sum(kd * L for kd, L in scans) == 335_872_000               # 8a: sum K*D*L per sample
    assert sum(4 * 16 * L for _, L in scans) == 118_882_304            # 8a: sum K*N*L per sample
    total = sum(bench.scan_bytes(12, kd, L)["total"] for kd, L in scans)
    assert abs(total / 1e9 - 163.2) < 0.1                              # 163.2 GB per step
    b = bench.scan_bytes(2, 768, 4096)                                 # BASELINE config 1
    assert abs(b["fwd"] / 1e6 - 79.7) < 0.1 and abs(b["total"] / 1e6 - 213.9) < 0.1


_WORKER = r"""
import os, sys, json
sys.path.insert(0, %r)
import torch, torch.distributed as dist
import bench
dist.init_process_group("gloo")
rank = dist.get_rank()
ms = bench.max_over_ranks(10.0 + 5.0 * rank, torch.device("cpu"), dist.get_world_size())
if rank == 0:
    print(json.dumps({"ms": ms, "world": dist.get_world_size()}))
dist.destroy_process_group()
"""


def test_timing_is_the_max_over_ranks_world_size_2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29571")
    out = subprocess.check_output([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                                   "--master-addr", "127.0.0.1", "--master-port", "29571", str(script)],
                                  env=env, text=True, stderr=subprocess.DEVNULL, timeout=180)
    line = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
    assert line == {"ms": 15.0, "world": 2}


def test_reference_arm_other_ranks_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                                  env=env, text=True, timeout=60)
    assert out.strip() == ""


def test_exposed_nccl_interval_arithmetic():
    """tools/prof_ddp.py: EXPOSED NCCL time = length of (union of NCCL kernel intervals) minus (union of compute kernel
    intervals)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "prof_ddp", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "prof_ddp.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    nccl = m.union([(0, 10), (5, 12), (20, 30), (40, 41)])
    comp = m.union([(2, 6), (8, 9), (11, 25), (50, 60)])
    assert nccl == [[0, 12], [20, 30], [40, 41]]
    assert m.length(nccl) == 23
    # exposed: [0,2) + [6,8) + [9,11) + [25,30) + [40,41) = 2 + 2 + 2 + 5 + 1
    assert m.subtract(nccl, comp) == 12
    assert m.subtract(nccl, []) == 23
    assert m.subtract([], comp) == 0
