"""Registries with the detectron2 surface (`META_ARCH_REGISTRY.get("HIPIE_IMG")(cfg)`, `build_model(cfg)`;
/root/reference/detectron2/modeling/meta_arch/build.py:7-25) and a `DetectionCheckpointer` that loads the reference's
`.pth` files ({"model": state_dict, ...}; detectron2/checkpoint/detection_checkpoint.py:15-125)."""
import torch


class Registry(dict):
    def __init__(self, name):
        super().__init__()
        self._name = name

    def register(self, obj=None):
        if obj is None:
            def deco(o):
                self[o.__name__] = o
                return o
            return deco
        self[obj.__name__] = obj
        return obj

    def get(self, name):
        if name not in self:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return self[name]


META_ARCH_REGISTRY = Registry("META_ARCH")
BACKBONE_REGISTRY = Registry("BACKBONE")
SEM_SEG_HEADS_REGISTRY = Registry("SEM_SEG_HEADS")
TRANSFORMER_DECODER_REGISTRY = Registry("TRANSFORMER_MODULE")


def _register_defaults():
    from .modeling.components import D2ViT, MaskDINODecoder, MaskDINOEncoder, MaskDINOHead
    from .modeling.hipie_img import HIPIE_IMG
    META_ARCH_REGISTRY.setdefault("HIPIE_IMG", HIPIE_IMG)
    BACKBONE_REGISTRY.setdefault("D2ViT", D2ViT)
    SEM_SEG_HEADS_REGISTRY.setdefault("MaskDINOHead", MaskDINOHead)
    SEM_SEG_HEADS_REGISTRY.setdefault("MaskDINOEncoder", MaskDINOEncoder)
    TRANSFORMER_DECODER_REGISTRY.setdefault("MaskDINODecoder", MaskDINODecoder)


def build_backbone(cfg, input_shape=None):
    """detectron2.modeling.build_backbone: BACKBONE_REGISTRY.get(cfg.MODEL.BACKBONE.NAME)(cfg, input_shape)"""
    _register_defaults()
    return BACKBONE_REGISTRY.get(cfg.MODEL.BACKBONE.NAME)(cfg, input_shape)


def install_detectron2_aliases():
    """Make `import detectron2.projects.hipie` resolve to this package (the reference maps projects/HIPIE/hipie there through
    detectron2/projects/__init__.py:5-33) and publish the registries under the detectron2 module paths the reference imports them
    from.  With a real detectron2 installed the classes are registered into ITS registries instead."""
    import sys
    import types
    _register_defaults()
    import hipie_b200
    try:
        import detectron2.modeling as d2m            # a real detectron2: add our entries to its registries
        for reg, mine in ((d2m.META_ARCH_REGISTRY, META_ARCH_REGISTRY), (d2m.BACKBONE_REGISTRY, BACKBONE_REGISTRY),
                          (d2m.SEM_SEG_HEADS_REGISTRY, SEM_SEG_HEADS_REGISTRY)):
            for name, obj in mine.items():
                if name not in reg:
                    reg.register(obj)
    except Exception:
        for name in ("detectron2", "detectron2.projects", "detectron2.modeling", "detectron2.checkpoint", "detectron2.config"):
            if name not in sys.modules:
                m = types.ModuleType(name)
                m.__path__ = []
                sys.modules[name] = m
        d2m = sys.modules["detectron2.modeling"]
        d2m.META_ARCH_REGISTRY, d2m.BACKBONE_REGISTRY, d2m.SEM_SEG_HEADS_REGISTRY = META_ARCH_REGISTRY, BACKBONE_REGISTRY, SEM_SEG_HEADS_REGISTRY
        d2m.build_model, d2m.build_backbone = build_model, build_backbone
        sys.modules["detectron2.checkpoint"].DetectionCheckpointer = DetectionCheckpointer
        from . import config as _cfg
        sys.modules["detectron2.config"].get_cfg = _cfg.get_cfg
    sys.modules["detectron2.projects.hipie"] = hipie_b200
    return hipie_b200


def build_model(cfg):
    _register_defaults()
    model = META_ARCH_REGISTRY.get(cfg.MODEL.META_ARCHITECTURE)(cfg)
    return model


class DetectionCheckpointer:
    def __init__(self, model, save_dir=""):
        self.model = model

    def load(self, path, checkpointables=None):
        if not path:
            return {}
        ckpt = torch.load(path, map_location="cpu")
        sd = ckpt["model"] if isinstance(ckpt, dict) and "model" in ckpt else ckpt
        missing = self.model.load_state_dict(sd, strict=False)
        unexpected = list(getattr(self.model, "unexpected_keys", []))
        if missing or unexpected:          # detection_checkpoint.py:105-125 logs both lists; so do we
            import logging
            log = logging.getLogger("hipie_b200.checkpoint")
            if missing:
                log.warning("checkpoint %s: %d keys of the model are missing (e.g. %s)", path, len(missing), missing[:5])
            if unexpected:
                log.warning("checkpoint %s: %d keys are not used by the model (e.g. %s)", path, len(unexpected), unexpected[:5])
        return {"missing_keys": missing, "unexpected_keys": unexpected}

    def save(self, path):
        torch.save({"model": self.model.state_dict()}, path)
