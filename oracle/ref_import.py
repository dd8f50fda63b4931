"""Import machinery that lets the UNMODIFIED reference modules under /root/reference execute in this container
(test infrastructure; used by oracle/gen_golden_modules.py only, never by product code).

The reference cannot be imported as a package: its third-party dependencies (fvcore, yacs, iopath, timm, fairscale,
open_clip, pycocotools, skimage ...) are absent, detectron2's own __init__ chain needs them, and its MSDeformAttn CUDA
extension has no CPU implementation.  What this module does instead:

  * the package tree `hipie` (= projects/HIPIE/hipie) is mounted through stub packages whose __path__ points at the
    real directories, so `import hipie.backbone.vit` executes the real file without running the package __init__ chain;
  * detectron2's arithmetic-carrying leaf files (layers/wrappers.py, layers/batch_norm.py, layers/blocks.py,
    layers/shape_spec.py) are loaded from /root/reference/detectron2 by path — the real code;
  * everything else that is absent resolves to a permissive stub module (decorators become identities, classes become
    plain bases, functions raise when CALLED), which is enough because only eval-forward code is executed;
  * dependencies whose arithmetic IS on the path but whose package is absent are restated next to their stub and listed in
    tests/golden/MANIFEST.md: timm.models.layers.Mlp (fc1 -> GELU(erf) -> fc2) and DropPath (identity at eval);
  * the pybind module `MultiScaleDeformableAttention` is bound to the reference's own pure-PyTorch statement of the op
    (`ms_deform_attn_core_pytorch`, ops/functions/ms_deform_attn_func.py:43-63 — "for debug and test only", same function
    as the CUDA kernel per the reference's ops/test.py).
"""
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = "/root/reference"
HIPIE = f"{REF_ROOT}/projects/HIPIE/hipie"
STUB_PREFIXES = ("fvcore", "detectron2", "timm", "fairscale", "iopath", "yacs", "open_clip", "pycocotools", "skimage", "panopticapi",
                 "lvis", "cv2", "shapely", "termcolor", "tabulate", "cloudpickle", "omegaconf", "hydra", "PIL", "scipy.optimize",
                 "segment_anything", "kornia", "ftfy", "clip")


class _Anything:
    """Stand-in for any absent class / function / decorator."""

    def __init__(self, name="stub"):
        self._name = name

    def __call__(self, *a, **k):
        # decorator use: @stub or @stub(...)  -> identity
        if len(a) == 1 and not k and (isinstance(a[0], type) or callable(a[0])):
            return a[0]
        return _Anything(self._name + "()")

    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Anything(f"{self._name}.{n}")

    def __mro_entries__(self, bases):
        return (object,)

    def __iter__(self):
        return iter(())


class _StubModule(types.ModuleType):
    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Anything(f"{self.__name__}.{n}")


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in {p.split(".")[0] for p in STUB_PREFIXES} and any(
                fullname == p or fullname.startswith(p + ".") or p.startswith(fullname + ".") for p in STUB_PREFIXES):
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def _mount(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    m.__package__ = name
    sys.modules[name] = m
    return m


def _load_file(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


class Mlp(nn.Module):
    """timm.models.layers.Mlp (timm is absent): fc1 -> act_layer() -> drop -> fc2 -> drop, attribute names as in timm."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0, **kw):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop)
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop2 = nn.Dropout(drop)

    def forward(self, x):
        return self.drop2(self.fc2(self.drop1(self.act(self.fc1(x)))))


class DropPath(nn.Module):
    """timm DropPath: identity in eval mode (and for drop_prob 0)."""

    def __init__(self, drop_prob=0.0, *a, **k):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        assert not self.training or self.drop_prob == 0.0
        return x


_installed = False


def install():
    """Idempotent.  After this, `import hipie.<module path>` runs the reference file."""
    global _installed
    if _installed:
        return
    _installed = True
    import transformers  # noqa: F401  real package; must be imported before timm is stubbed
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu
    for n in ("apply_chunking_to_forward", "prune_linear_layer", "find_pruneable_heads_and_indices"):
        if not hasattr(mu, n) and hasattr(pu, n):
            setattr(mu, n, getattr(pu, n))
    if not hasattr(mu, "find_pruneable_heads_and_indices"):
        mu.find_pruneable_heads_and_indices = _Anything("find_pruneable_heads_and_indices")
    sys.meta_path.append(_StubFinder())

    # ---- real detectron2 leaf files
    import detectron2  # stub
    import detectron2.layers as d2l
    import detectron2.utils  # noqa: F401
    import fvcore.nn.weight_init as wi
    wi.c2_xavier_fill = lambda m: (nn.init.kaiming_uniform_(m.weight, a=1), m.bias is not None and nn.init.constant_(m.bias, 0))
    wi.c2_msra_fill = lambda m: (nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu"),
                                 m.bias is not None and nn.init.constant_(m.bias, 0))
    wr = _load_file("detectron2.layers.wrappers", f"{REF_ROOT}/detectron2/layers/wrappers.py")
    import detectron2.utils.env as d2env
    d2env.TORCH_VERSION = tuple(int(x) for x in torch.__version__.split(".")[:2])
    bn = _load_file("detectron2.layers.batch_norm", f"{REF_ROOT}/detectron2/layers/batch_norm.py")
    bl = _load_file("detectron2.layers.blocks", f"{REF_ROOT}/detectron2/layers/blocks.py")
    ss = _load_file("detectron2.layers.shape_spec", f"{REF_ROOT}/detectron2/layers/shape_spec.py")
    for k in ("Conv2d", "ConvTranspose2d", "cat", "interpolate", "Linear", "nonzero_tuple", "cross_entropy", "shapes_to_tensor"):
        if hasattr(wr, k):
            setattr(d2l, k, getattr(wr, k))
    d2l.get_norm, d2l.FrozenBatchNorm2d, d2l.NaiveSyncBatchNorm = bn.get_norm, bn.FrozenBatchNorm2d, getattr(bn, "NaiveSyncBatchNorm", None)
    d2l.CNNBlockBase, d2l.ShapeSpec = bl.CNNBlockBase, ss.ShapeSpec
    import detectron2.modeling as d2m
    d2m.ShapeSpec = ss.ShapeSpec

    class Backbone(nn.Module):      # detectron2/modeling/backbone/backbone.py:10-53 (abstract base; no arithmetic)
        @property
        def size_divisibility(self):
            return 0

        def output_shape(self):
            return {name: ss.ShapeSpec(channels=self._out_feature_channels[name], stride=self._out_feature_strides[name])
                    for name in self._out_features}
    d2m.Backbone = Backbone

    # ---- restated absent arithmetic
    import timm.models.layers as tl
    tl.Mlp, tl.DropPath = Mlp, DropPath
    tl.trunc_normal_ = nn.init.trunc_normal_

    # ---- the pybind op -> the reference's own pure-PyTorch statement of it
    msda = types.ModuleType("MultiScaleDeformableAttention")
    sys.modules["MultiScaleDeformableAttention"] = msda

    # ---- mount the package tree without running the __init__ chain
    _mount("hipie", HIPIE)
    for sub in ("backbone", "models", "util", "data", "open_vocab", "models/deformable_detr", "models/deformable_detr/ops",
                "models/deformable_detr/ops/functions", "models/deformable_detr/ops/modules", "models/maskdino",
                "models/maskdino/utils", "models/maskdino/pixel_decoder", "models/maskdino/pixel_decoder/ops",
                "models/maskdino/pixel_decoder/ops/functions", "models/maskdino/pixel_decoder/ops/modules",
                "models/maskdino/transformer_decoder", "models/maskdino/meta_arch", "models/maskdino/backbone", "data/datasets", "models/sam",
                "open_vocab"):
        if os.path.isdir(os.path.join(HIPIE, sub)):
            _mount("hipie." + sub.replace("/", "."), os.path.join(HIPIE, sub))
    for pkg in ("hipie.models.deformable_detr.ops", "hipie.models.maskdino.pixel_decoder.ops"):
        f = importlib.import_module(pkg + ".functions.ms_deform_attn_func")
        sys.modules[pkg + ".functions"].MSDeformAttnFunction = f.MSDeformAttnFunction
        sys.modules[pkg + ".functions"].ms_deform_attn_core_pytorch = f.ms_deform_attn_core_pytorch
        core = f.ms_deform_attn_core_pytorch
    msda.ms_deform_attn_forward = lambda value, shapes, lsi, loc, w, step: core(value, shapes, loc, w)
    for pkg in ("hipie.models.deformable_detr.ops", "hipie.models.maskdino.pixel_decoder.ops"):
        m = importlib.import_module(pkg + ".modules.ms_deform_attn")
        sys.modules[pkg + ".modules"].MSDeformAttn = m.MSDeformAttn


def ref(module_path):
    """import a reference module, e.g. ref('backbone.vit')."""
    install()
    return importlib.import_module("hipie." + module_path)
