"""CPU: host-side logic, the drop-in boundary, and that the C-ABI library loads and exports every symbol the
header declares (no kernel is launched here - there is no GPU)."""
import ctypes
import os
import re
import subprocess
import sys
import warnings

import numpy as np
import pytest
import torch

import cova_b200.synth as synth
from conftest import ROOT, load_golden

warnings.filterwarnings("ignore")


@pytest.fixture(scope="module")
def built_lib():
    from cova_b200 import _lib
    _lib.build()
    return _lib


def test_header_symbols_all_exported_and_bound(built_lib):
    hdr = open(os.path.join(ROOT, "include", "cova_b200.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*)\s+(cova_\w+)\(", hdr, flags=re.M))
    assert len(declared) >= 11
    assert declared == set(built_lib.SIGNATURES), "ctypes table and header disagree"
    h = ctypes.CDLL(built_lib.LIB_PATH)
    for name in declared:
        assert hasattr(h, name), name
    assert built_lib.lib().cova_abi_version() == 1


def test_library_is_sm100a_tcgen05_tma(built_lib):
    sass = subprocess.run(["cuobjdump", "-sass", built_lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "UBLKCP"):   # tcgen05.mma, TMA tensor load, tcgen05.ld, TMA bulk
        assert mnemonic in sass, mnemonic


def test_argument_validation_returns_error_not_crash(built_lib):
    L = built_lib.lib()
    rc = L.cova_roi_fwd(None, 1, 8, 8, 64, None, 1, 3, 3, 0.25, 0, 2, None, 576, None, None)
    assert rc == 1 and b"null" in L.cova_last_error()
    assert L.cova_roi_fwd(None, 1, 8, 8, 64, None, 0, 3, 3, 0.25, 0, 2, None, 576, None, None) == 0   # no boxes: no-op
    rc = L.cova_gat_fwd(1, 384, 1, 1, 1, 0.0, 0.2, 1, 4, 500, 384, 1, 384, None, None)
    assert rc == 1 and b"K=500" in L.cova_last_error()


def test_missing_library_fails_loudly(monkeypatch):
    from cova_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libcova_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()


def _model(**kw):
    from cova_b200.models import CoVA
    return CoVA((3, 3), 1280, 4, True, 384, 32, 0, 0.2, None, pretrained=False, **kw)


def test_state_dict_keys_match_reference():
    for bk in ("resnet18", "resnet50"):
        g = load_golden("state_dict_keys_" + bk)
        m = _model(backbone=bk)
        sd = m.state_dict()
        assert list(sd.keys()) == list(g["keys"])
        assert [str(tuple(v.shape)) for v in sd.values()] == list(g["shapes"])
        m.load_state_dict(synth.make_state_dict(123, backbone=bk), strict=True)
        assert all(p.requires_grad for p in m.parameters())   # nothing frozen


def test_constructor_contract():
    from cova_b200.models import CoVA, GraphAttentionLayer
    m = CoVA((3, 3), 1280, 4, True, 384, 32, 0, 0.2, ["BG", "Price", "Title", "Image"], pretrained=False)
    assert m.n_classes == 4 and list(m.class_names) == ["BG", "Price", "Title", "Image"]
    assert (m.n_visual_feat, m.n_feat, m.n_total_feat) == (576, 608, 992) and m.spatial_scale == 0.25
    assert isinstance(m.gat, GraphAttentionLayer) and m.bn_additional_feat(3) == 3
    m = CoVA((3, 3), 1280, 4, False, 0, 32, 0, 0.2, None, pretrained=False)     # main.py:58-59 (context_size == 0)
    assert not hasattr(m, "gat") and m.n_total_feat == 608
    m = CoVA((3, 3), 1280, 4, pretrained=False, backbone="resnet50")
    assert (m.n_visual_feat, m.n_feat, m.n_total_feat) == (2304, 2336, 2720)


def test_cpu_tensors_raise():
    m = _model().eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA tensors"):
        m(torch.zeros(1, 3, 1280, 1280), torch.zeros(2, 5), torch.zeros(2, 0), torch.zeros(2, 24, dtype=torch.long))


def test_synth_layout_matches_collate():
    images, bboxes, add, ci = synth.gen(3, 0, 8, seed=1, img=64, counts=[5, 1, 9])
    assert images.shape == (3, 3, 64, 64) and bboxes.shape == (15, 5) and add.shape == (15, 0) and ci.dtype == torch.int64
    assert bboxes[:, 0].tolist() == [0] * 5 + [1] + [2] * 9
    assert (bboxes[:, 3] > bboxes[:, 1]).all() and (bboxes[:, 4] > bboxes[:, 2]).all()
    assert ci[5].tolist() == [-1] * 8                       # single-box page: all padding
    assert ci[6].tolist() == [7, 8, 9, 10, -1, -1, -1, -1]  # first box of page 2: ids are batch-global
    assert ci[0].tolist() == [1, 2, 3, 4, -1, -1, -1, -1]


def test_shard_batch_rebases_indices():
    from cova_b200.dist import page_range, shard_batch
    assert [page_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    batch = synth.gen(4, 0, 8, seed=2, img=32, counts=[3, 6, 2, 5], with_labels=True)
    parts = [shard_batch(*batch, rank=r, world=2) for r in range(2)]
    assert sum(p[1].shape[0] for p in parts) == 16
    im, bb, add, ci, lab = parts[1]
    assert im.shape[0] == 2 and bb[:, 0].tolist() == [0, 0, 1, 1, 1, 1, 1]
    assert ci.max() < 7 and (ci[ci >= 0] >= 0).all()
    ref = synth.gen(2, 0, 8, seed=0, img=32, counts=[2, 5])[3]
    assert torch.equal(ci, ref)                             # identical to a batch built from those pages alone
    assert torch.equal(lab, batch[4][9:])


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import warnings; warnings.filterwarnings("ignore")
from cova_b200.dist import FlatGradBucket, shard_batch
import cova_b200.synth as synth
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
torch.manual_seed(0)
lin = torch.nn.Linear(6, 3)                      # same init on every rank (same seed)
bucket = FlatGradBucket(lin)
batch = synth.gen(4, 0, 4, seed=5, img=16, counts=[3, 2, 4, 1], with_labels=True)
def loss_of(bb, lab):
    x = torch.cat([bb, bb[:, 1:2] * bb[:, 2:3] / 256.0], 1)
    return torch.nn.functional.cross_entropy(lin(x / 16.0), lab.clamp_max(2), reduction="sum")
bucket.zero()
_, bb, _, _, lab = shard_batch(*batch, rank=rank, world=world)
loss_of(bb, lab).backward()
bucket.allreduce_sum()
got = bucket.flat.clone()
torch.save(got, sys.argv[2] + f".{rank}")
dist.destroy_process_group()
'''


def test_grad_allreduce_sum_world2_gloo(tmp_path):
    """world_size-2 gloo run: the all-reduced flat gradient bucket is identical on both ranks and equals the
    sum of the per-shard gradients (SUM semantics for a sum-reduced loss)."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    out = tmp_path / "res"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", str(script), ROOT, str(out)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    g0 = torch.load(str(out) + ".0")
    g1 = torch.load(str(out) + ".1")
    assert torch.equal(g0, g1) and g0.numel() == 6 * 3 + 3 and float(g0.abs().sum()) > 0
    # single-process recomputation of the per-shard gradients, summed
    from cova_b200.dist import shard_batch
    torch.manual_seed(0)
    lin = torch.nn.Linear(6, 3)
    batch = synth.gen(4, 0, 4, seed=5, img=16, counts=[3, 2, 4, 1], with_labels=True)
    tot = torch.zeros(21)
    for rank in range(2):
        _, bb, _, _, lab = shard_batch(*batch, rank=rank, world=2)
        x = torch.cat([bb, bb[:, 1:2] * bb[:, 2:3] / 256.0], 1)
        loss = torch.nn.functional.cross_entropy(lin(x / 16.0), lab.clamp_max(2), reduction="sum")
        gs = torch.autograd.grad(loss, list(lin.parameters()))
        tot += torch.cat([g.flatten() for g in gs])
    assert torch.allclose(g0, tot, rtol=1e-5, atol=1e-6)


def test_native_tail_fails_loudly_on_cpu_tensors():
    """The callers either side of the forward have no CPU path either: criterion and optimizer raise on CPU inputs."""
    import pytest as _pytest
    import torch
    from cova_b200.train_ops import CrossEntropyLossSum, FlatAdam
    with _pytest.raises(RuntimeError, match="CUDA"):
        CrossEntropyLossSum()(torch.zeros(3, 4), torch.zeros(3, dtype=torch.long))
    with _pytest.raises(RuntimeError, match="CUDA"):
        FlatAdam([torch.nn.Parameter(torch.zeros(5))], lr=1e-3)


def test_page_offsets_and_numa_binding_host_logic():
    """Host-side helpers of the native callers: page offsets of a collated batch (pages are contiguous row ranges,
    `datasets.py:170-181`) and the NUMA binding helper (returns None instead of raising when NVML has no device)."""
    import torch
    from cova_b200.pipeline import bind_to_gpu_numa
    from cova_b200.train_ops import page_offsets_of
    bb = torch.zeros(9, 5)
    bb[:, 0] = torch.tensor([0, 0, 0, 1, 2, 2, 2, 2, 2]).float()
    assert page_offsets_of(bb).tolist() == [0, 3, 4, 9]
    assert page_offsets_of(torch.zeros(0, 5)).tolist() == [0]
    cores = bind_to_gpu_numa(0)
    assert cores is None or (isinstance(cores, list) and len(cores) > 0)


def test_bench_alg_table_matches_abi_signatures():
    """bench.py derives the algorithmic work of a launch from the ABI call's own arguments by POSITION: every entry must
    name a declared symbol and index inside its argument list (a signature change must not silently zero a roofline)."""
    import bench
    from cova_b200 import _lib
    table = bench._alg_table(1.0e6)
    for name, fn in table.items():
        assert name in _lib.SIGNATURES, name
        nargs = len(_lib.SIGNATURES[name][1])
        kind, work = fn([8] * nargs)
        assert kind in ("tensor", "hbm") and work > 0, name
    for cid, cfg in bench.CONFIGS.items():
        assert cfg["mode"] in ("infer", "train") and cfg["backbone"] in ("resnet18", "resnet50")
    assert bench.get_config(3)["precision"] == "bf16" and bench.get_config(2).get("precision", "fp32") == "fp32"
