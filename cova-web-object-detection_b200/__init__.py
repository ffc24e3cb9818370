"""cova_b200 - B200-native implementation of CoVA's per-webpage forward hot path
(`/root/reference/models.py` `CoVA.forward`), behind the reference's own `models.py` interface.

Import as ``import cova_b200`` (shim `cova_b200.py` at the repo root) - the directory name
``cova-web-object-detection_b200`` is not a Python identifier."""
__version__ = "0.1.0"
