"""mmdet `BACKBONES` registry hook (reference: mmdet/models/builder.py:6, used at
simple_panoswin_transformer.py:36-40, :779).  When mmdet/mmcv are importable the class registers into the
real registry, so a config with `type='SimplePanoSwinTransformer'` plus
`custom_imports=dict(imports=['panoswintransformerobjectdetection_b200'])` builds this implementation;
otherwise a minimal local registry with the same `register_module` / `build` surface is used."""
from __future__ import annotations


class _LocalRegistry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            key = name or cls.__name__
            if key in self.module_dict and not force:
                raise KeyError(f"{key} is already registered in {self.name}")
            self.module_dict[key] = cls
            return cls
        return deco(module) if module is not None else deco

    def get(self, key):
        return self.module_dict.get(key)

    def build(self, cfg: dict):
        """build_from_cfg semantics: `type` selects the class, the other keys are constructor kwargs
        (unknown keys raise TypeError, as plain Python does for the reference)."""
        cfg = dict(cfg)
        cls = self.get(cfg.pop("type"))
        if cls is None:
            raise KeyError(f"type not found in registry {self.name}")
        return cls(**cfg)


try:  # pragma: no cover - mmdet is not installed in the build image
    from mmdet.models.builder import BACKBONES  # type: ignore
    USING_MMDET_REGISTRY = True
except Exception:  # ImportError, or mmcv version asserts
    BACKBONES = _LocalRegistry("backbone")
    USING_MMDET_REGISTRY = False


def build_backbone(cfg: dict):
    """Mirror of mmdet.models.builder.build_backbone (mmdet/models/builder.py:38-40)."""
    if USING_MMDET_REGISTRY:  # pragma: no cover
        from mmdet.models.builder import build_backbone as _bb  # type: ignore
        return _bb(cfg)
    return BACKBONES.build(cfg)
