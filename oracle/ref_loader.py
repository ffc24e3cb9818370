"""ORACLE - TEST / BASELINE INFRASTRUCTURE ONLY.  Imports the reference's own modules from `oracle/_ref/`
(staged by `oracle/build_ref.py`; never from `/root/reference`, which does not exist on the GPU box).

The only patch is the one SURVEY.md section 8(c) documents: `torchvision.models.resnet18(pretrained=True)`
(`models.py:49`) would download a checkpoint; here it returns the same architecture with `weights=None`
(or resnet50).  The modules are imported under their own top-level names (`models`, `train`, `utils` ...,
as the reference's files import each other) and removed from `sys.modules` again, so nothing else in the process sees
them (`datasets` would shadow the HuggingFace package of the same name)."""
import importlib
import os
import sys
import types

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_NAMES = ["utils", "constants", "datasets", "models", "train", "evaluate"]
_backbone = {"name": "resnet18"}
_cache = {}


def available():
    return all(os.path.exists(os.path.join(REF_DIR, n + ".py")) for n in _NAMES)


def set_backbone(name):
    assert name in ("resnet18", "resnet50")
    _backbone["name"] = name


def load(models_module=None):
    """Returns a namespace with the reference's modules.  `models_module`: a module to install as `models` while
    `evaluate.py` (`from models import CoVA`) is imported - the INTEGRATION.md stub that re-exports this repo's CoVA;
    default = the reference's own models.py."""
    key = id(models_module)
    if key in _cache:
        return _cache[key]
    if not available():
        raise RuntimeError("oracle/_ref is missing - run `python oracle/build_ref.py` where /root/reference exists")
    import torchvision
    if not getattr(torchvision.models.resnet18, "_cova_patched", False):
        r18, r50 = torchvision.models.resnet18, torchvision.models.resnet50

        def factory(pretrained=True, **kw):
            return r18(weights=None) if _backbone["name"] == "resnet18" else r50(weights=None)
        factory._cova_patched = True
        torchvision.models.resnet18 = factory
    saved = {n: sys.modules.get(n) for n in _NAMES}
    sys.path.insert(0, REF_DIR)
    try:
        for n in _NAMES:
            sys.modules.pop(n, None)
        mods = {}
        for n in _NAMES:
            if n == "models" and models_module is not None:
                sys.modules["models"] = models_module
                mods[n] = models_module
            else:
                mods[n] = importlib.import_module(n)
        for n, m in mods.items():
            if m is not models_module and not os.path.abspath(m.__file__).startswith(REF_DIR):
                raise RuntimeError("reference module %s resolved to %s" % (n, m.__file__))
    finally:
        sys.path.remove(REF_DIR)
        for n in _NAMES:
            sys.modules.pop(n, None)
            if saved[n] is not None:
                sys.modules[n] = saved[n]
    ns = types.SimpleNamespace(**mods)
    _cache[key] = ns
    return ns


def build_model(backbone="resnet18", img=1280, n_heads=1, drop=0.2, state_dict=None):
    """The reference's `models.CoVA` built the way `main.py:122-132` builds it (9 positional arguments).  SURVEY D2: the
    backbone factory returns resnet50 when asked; D3: `n_heads` > 1 composes H of the reference's own
    `GraphAttentionLayer(n_feat, hidden // H)` on the same inputs, outputs concatenated (keys `gat.heads.{i}.*`)."""
    import torch
    ref = load()
    set_backbone(backbone)
    try:
        m = ref.models.CoVA((3, 3), img, 4, True, 384, 32, 0, drop, None)
    finally:
        set_backbone("resnet18")
    if n_heads > 1:
        layer_cls = ref.models.GraphAttentionLayer

        class RefMultiHead(torch.nn.Module):
            def __init__(self, n_feat, hidden, H):
                super().__init__()
                self.heads = torch.nn.ModuleList(layer_cls(n_feat, hidden // H) for _ in range(H))

            def forward(self, h, ci, return_attn_wts=False):
                return torch.cat([hd(h, ci) for hd in self.heads], 1)
        m.gat = RefMultiHead(m.n_feat, 384, n_heads)
    if state_dict is not None:
        m.load_state_dict(state_dict, strict=True)
    return m
