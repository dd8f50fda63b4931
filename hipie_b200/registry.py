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


def _register_defaults():
    from .modeling.hipie_img import HIPIE_IMG
    META_ARCH_REGISTRY.setdefault("HIPIE_IMG", HIPIE_IMG)


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
        return {"missing_keys": missing}
